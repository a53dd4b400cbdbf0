"""GPU: the opt-in chunked DSA kernel (PYDCOP_B200_DSA_V2=2|4, csrc/dsa_v2.cu) against the default
fast kernel and the oracle: identical assignments every cycle."""
import numpy as np
import pytest

import oracle as orc
from pydcop_b200.generators import random_factor_graph
from pydcop_b200.layout import build_layout, default_var_csr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,variant,mode,precision,chunk", [
    (20, "B", "min", "f32", 4), (20, "B", "min", "f32", 2), (20, "C", "max", "f64", 4),
    (10, "A", "min", "f32", 4), (16, "B", "min", "f64", 2), (8, "C", "min", "f32", 4), (4, "B", "max", "f64", 2),
])
def test_chunked_kernel_equals_default_and_oracle(monkeypatch, d, variant, mode, precision, chunk):
    from pydcop_b200.engine import DsaEngine
    n = 20000
    inst = random_factor_graph(n, d, n * 3, 2, seed=d + chunk, noise=0.0, int_tables=False)
    inst["tables"] = np.round(inst["tables"] / 3.0).astype(np.float32)
    L = build_layout(**inst)
    kw = dict(precision=precision, mode=mode, variant=variant, probability=0.6, seed=77)
    monkeypatch.delenv("PYDCOP_B200_DSA_V2", raising=False)
    base = DsaEngine(L, **kw).init()
    monkeypatch.setenv("PYDCOP_B200_DSA_V2", str(chunk))
    v2 = DsaEngine(L, **kw).init()
    assert v2._v2_chunk == chunk and base._v2_chunk == 0
    vp, ve = default_var_csr(n, inst["edge_var"])
    o = orc.DsaOracle(dict(inst, var_ptr=vp, var_edge=ve), np.float64 if precision == "f64" else np.float32,
                      **{k: v for k, v in kw.items() if k != "precision"}).init()
    assert np.array_equal(v2.values(), o.val)
    for k in range(8):
        o.step()
        base.step()
        v2.step()
        assert np.array_equal(base.values(), o.val), k
        assert np.array_equal(v2.values(), o.val), k
    assert v2.launch_count > base.launch_count - 8 and v2._v2_launches == 8


def test_unsupported_shapes_fall_back_to_the_default_path(monkeypatch):
    from pydcop_b200.engine import DsaEngine
    monkeypatch.setenv("PYDCOP_B200_DSA_V2", "4")
    inst = random_factor_graph(500, 3, 900, 2, seed=1, noise=0.0)     # d=3: not compiled for the chunked kernel
    e = DsaEngine(build_layout(**inst), precision="f32", seed=3).init().step(3)
    assert e._v2_chunk == 0 and e.cycle == 3
