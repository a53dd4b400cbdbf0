"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/pydcop_b200.h declares; without a GPU the product path fails loudly (no CPU fallback)."""
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pydcop_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    from pydcop_b200 import _cabi, build
    build.build()
    lib = _cabi.load()
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    assert sorted(_cabi.SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fg_abi_version() == _cabi.FG_ABI_VERSION


def test_struct_sizes_match_header():
    """sizeof of the ctypes mirrors == what the C compiler lays out for include/pydcop_b200.h."""
    import ctypes as C
    import subprocess
    import tempfile
    from pydcop_b200 import _cabi
    prog = r'''
#include <stdio.h>
#include "pydcop_b200.h"
int main(void){printf("%zu %zu %zu\n", sizeof(fg_class_t), sizeof(fg_maxsum_desc_t), sizeof(fg_dsa_desc_t));return 0;}
'''
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "s.c"), os.path.join(td, "s")
        open(src, "w").write(prog)
        subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-I",
                        os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [C.sizeof(_cabi.FgClass), C.sizeof(_cabi.FgMaxSumDesc),
                                     C.sizeof(_cabi.FgDsaDesc)]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from pydcop_b200 import build_layout
    from pydcop_b200.engine import EngineError, MaxSumEngine, DsaEngine
    L = build_layout([2, 2], [0, 2], [0, 1], np.zeros(4))
    with pytest.raises(EngineError):
        MaxSumEngine(L)
    with pytest.raises(EngineError):
        DsaEngine(L)
