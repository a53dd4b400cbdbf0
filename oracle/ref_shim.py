"""Import shims that let the UNMODIFIED reference (pyDcop @ /root/reference) import under
Python 3.12 / numpy 2 in the build container.  TEST INFRASTRUCTURE ONLY.

Only `oracle/make_golden.py` (fixture generation, build container only), the plugin tests and
bench.py's `reference-threadmode` baseline use this; nothing in the product path imports it.
/root/reference does not exist on the GPU box: there the reference is the unmodified copy that
`python -m pip install --target baseline/_ref` made in the build container (git-ignored, shipped
with the snapshot) — used by the drop-in test that runs the REAL engine under the reference's own
orchestrator and by the timed thread-mode baseline, never by the product path.

Shims:
  1. collections.Iterable/Mapping/... aliases (reference: pydcop/dcop/yamldcop.py:32, dcop.py:238)
  2. stub `websocket_server` package (reference: pydcop/infrastructure/ui.py:36)
  3. stub `pulp` package (reference: ILP distributions, commands/generators/iot.py:47-49)
"""
import collections
import collections.abc
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
#: the unmodified reference: the source tree in the build container, else the copy `pip install --target
#: baseline/_ref` made of it (git-ignored; it travels to the GPU box with the snapshot, __graft_entry__.build)
INSTALLED_ROOT = os.path.join(_REPO, "baseline", "_ref")
REFERENCE_ROOT = os.environ.get("PYDCOP_REFERENCE") or (
    "/root/reference" if os.path.isdir("/root/reference/pydcop") else INSTALLED_ROOT)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pydcop"))


def install() -> None:
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for n in ("Iterable", "Mapping", "Callable", "Sized", "MutableMapping", "Sequence",
              "Hashable", "Container", "Set", "MutableSet", "MutableSequence"):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    if "websocket_server" not in sys.modules:
        ws = types.ModuleType("websocket_server")
        ws2 = types.ModuleType("websocket_server.websocket_server")
        ws2.WebsocketServer = type(
            "WebsocketServer", (), {"__init__": lambda self, *a, **k: None})
        ws.websocket_server = ws2
        sys.modules["websocket_server"] = ws
        sys.modules["websocket_server.websocket_server"] = ws2
    if "pulp" not in sys.modules:
        class _Stub(types.ModuleType):
            def __getattr__(self, n):
                if n.startswith("__"):
                    raise AttributeError(n)
                return object
        for name in ("pulp", "pulp.constants", "pulp.pulp", "pulp.solvers"):
            sys.modules[name] = _Stub(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
