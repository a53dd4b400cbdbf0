"""The chunked DSA kernel SOURCE (pydcop_b200/csrc/dsa_v2_kernels.cuh, an opt-in experiment) run on
the CPU through tests/hostshim/dsa_v2_host.cpp with exactly the fast-path arrays DsaEngine builds
(`dsa_fast_arrays`), against the oracle and the reference trajectory.  Checks the kernel logic and
the oriented-table layout without a GPU; the device run is tests/test_gpu_zz_dsa_v2.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

import oracle as orc
from conftest import GOLDEN_DIR, ROOT
from pydcop_b200 import _cabi
from pydcop_b200.engine import dsa_fast_arrays
from pydcop_b200.generators import random_factor_graph
from pydcop_b200.layout import default_var_csr, layout_from_instance

SRC = os.path.join(ROOT, "tests", "hostshim", "dsa_v2_host.cpp")
SO = os.path.join(ROOT, "tests", "hostshim", "_build", "dsa_v2_host.so")
P = C.c_void_p


class _Arrays(C.Structure):
    _fields_ = [("var_ptr", P), ("slot_nbr", P), ("slot_tab", P), ("slot_opt", P), ("tables_or", P),
                ("has_nbr", P), ("prob", P), ("var_id", P), ("val", P), ("val_next", P), ("val_cost", P),
                ("n_vars", C.c_int32), ("precision", C.c_int32), ("dom", C.c_int32), ("chunk", C.c_int32),
                ("mode_max", C.c_int32), ("variant", C.c_int32), ("seed", C.c_uint64)]


@pytest.fixture(scope="module")
def shim():
    deps = [SRC, os.path.join(ROOT, "pydcop_b200", "csrc", "dsa_v2_kernels.cuh"),
            os.path.join(ROOT, "pydcop_b200", "csrc", "philox.cuh"), os.path.join(ROOT, "include", "pydcop_b200.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([gxx, "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO, SRC],
                       check=True, capture_output=True)
    return C.CDLL(SO)


class HostDsaV2:
    """The state DsaEngine keeps for the fast path, stepped by the host-shimmed chunked kernel."""

    def __init__(self, lib, inst, precision, chunk, mode="min", probability=0.7, variant="B", seed=0, **_):
        self.lib = lib
        self.L = L = layout_from_instance(inst)
        dt = np.float64 if precision == "f64" else np.float32
        fast = dsa_fast_arrays(L, torch.from_numpy(np.ascontiguousarray(L.tables, dtype=dt)), mode)
        assert fast is not None
        tables_or, slot_tab, slot_nbr, slot_opt, D = fast
        arity = np.array([c.arity for c in L.classes], dtype=np.int64)
        n_count = np.zeros(L.n_vars, dtype=np.int64)
        np.add.at(n_count, L.slot_var, arity[L.edge_class[L.slot_edge]] - 1)
        c = np.ascontiguousarray
        self.keep = dict(var_ptr=c(L.var_ptr, np.int32), slot_nbr=c(slot_nbr.numpy(), np.int32),
                         slot_tab=c(slot_tab.numpy(), np.int64), slot_opt=c(slot_opt.numpy(), dt),
                         tables_or=c(tables_or.numpy(), dt), has_nbr=c(n_count > 0, np.uint8),
                         prob=np.full(L.n_vars, float(probability)), var_id=c(L.var_order, np.int32),
                         val=np.zeros(L.n_vars, np.int32), val_next=np.zeros(L.n_vars, np.int32),
                         val_cost=np.zeros(L.n_vars, dt))
        self.meta = dict(n_vars=L.n_vars, precision=_cabi.FG_F64 if precision == "f64" else _cabi.FG_F32,
                         dom=D, chunk=chunk, mode_max=int(mode == "max"), variant=_cabi.DSA_VARIANTS[variant],
                         seed=int(seed))
        self.cycle = 0

    def set_values(self, canonical):
        self.keep["val"][:] = np.asarray(canonical, dtype=np.int32)[self.L.var_order]

    def step(self):
        a = _Arrays()
        for k, v in self.keep.items():
            setattr(a, k, P(v.ctypes.data))
        for k, v in self.meta.items():
            setattr(a, k, v)
        assert self.lib.dsa_v2_host_step(C.byref(a), C.c_uint32(self.cycle)) == 0
        self.keep["val"], self.keep["val_next"] = self.keep["val_next"], self.keep["val"]
        self.cycle += 1

    def values(self):
        return self.L.vars_to_canonical(self.keep["val"])


@pytest.mark.parametrize("chunk", [2, 4])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_chunked_kernel_source_matches_reference_trajectory(shim, precision, chunk):
    inst, meta = orc.load_golden(os.path.join(GOLDEN_DIR, "dsa_rand_d20.npz"))
    h = HostDsaV2(shim, inst, precision, chunk, mode=meta["mode"], seed=meta["seed"], **meta["params"])
    h.set_values(inst["value"][0])
    for k in range(1, meta["n_cycles"] + 1):
        h.step()
        assert np.array_equal(h.values(), inst["value"][k]), k


@pytest.mark.parametrize("d,variant,mode,precision,chunk", [
    (4, "A", "min", "f64", 4), (8, "B", "max", "f32", 2), (10, "C", "min", "f32", 4),
    (16, "B", "min", "f64", 2), (20, "B", "min", "f32", 4), (20, "C", "max", "f64", 4),
])
def test_chunked_kernel_source_matches_oracle(shim, d, variant, mode, precision, chunk):
    n = 2500
    inst = random_factor_graph(n, d, int(n * 2.6), 2, seed=d, noise=0.0, int_tables=False)
    inst["tables"] = np.round(inst["tables"] / 3.0).astype(np.float32)       # few levels: many ties
    inst["var_ptr"], inst["var_edge"] = default_var_csr(n, inst["edge_var"])
    dt = np.float64 if precision == "f64" else np.float32
    o = orc.DsaOracle(inst, dt, mode=mode, variant=variant, probability=0.6, seed=31).init()
    h = HostDsaV2(shim, inst, precision, chunk, mode=mode, variant=variant, probability=0.6, seed=31)
    h.set_values(o.val)
    moved = 0
    for k in range(8):
        prev = o.val.copy()
        o.step()
        h.step()
        assert np.array_equal(h.values(), o.val), k
        moved += int((prev != o.val).sum())
    assert moved > 0
    assert np.diff(inst["var_ptr"]).max() > 2 * chunk    # several trips and a ragged tail
