"""`python -m pydcop_b200.launcher <pydcop args>` == `pydcop <args>` with the GPU algorithm
modules registered, e.g.

    python -m pydcop_b200.launcher -t 10 solve --algo maxsum_gpu -d adhoc graph_coloring.yaml

`install()` appends pydcop_b200/algorithms to `pydcop.algorithms.__path__`, which is all the
unmodified reference needs to list (`list_available_algorithms`,
pydcop/algorithms/__init__.py:508-525) and load (`load_algorithm_module`, :527-566) them.
"""
import os
import sys


def install():
    import pydcop.algorithms as A
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "algorithms")
    if here not in list(A.__path__):
        A.__path__.append(here)
    return here


def main(argv=None):
    install()
    from pydcop.dcop_cli import main as pydcop_main
    if argv is not None:
        sys.argv = [sys.argv[0]] + list(argv)
    return pydcop_main()


if __name__ == "__main__":
    main()
