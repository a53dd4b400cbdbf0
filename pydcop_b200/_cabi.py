"""ctypes binding of include/pydcop_b200.h (the C-ABI of libpydcop_b200.so).

Fails loudly when the library is missing: there is no Python/CPU fallback for the hot path.
"""
import ctypes as C
import os

FG_ABI_VERSION = 1
FG_MAX_ARITY = 8
FG_MAX_DOM = 256
FG_OK, FG_ERR_ARG, FG_ERR_CUDA, FG_ERR_UNSUPPORTED = 0, 1, 2, 3
FG_F32, FG_F64 = 0, 1
START_MESSAGES = {"leafs": 0, "leafs_vars": 1, "all": 2}
DSA_VARIANTS = {"A": 0, "B": 1, "C": 2}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libpydcop_b200.so")


class FgClass(C.Structure):
    _fields_ = [("arity", C.c_int32), ("dom", C.c_int32 * FG_MAX_ARITY),
                ("row_off", C.c_int32 * FG_MAX_ARITY), ("row_total", C.c_int32),
                ("n_factors", C.c_int32), ("flags", C.c_int32), ("first_factor", C.c_int32),
                ("first_edge", C.c_int32),
                ("table_size", C.c_int64), ("table_base", C.c_int64), ("msg_base", C.c_int64)]


P = C.c_void_p


class FgVarClass(C.Structure):
    _fields_ = [("dom", C.c_int32), ("degree", C.c_int32), ("n_vars", C.c_int32),
                ("first_var", C.c_int32), ("first_slot", C.c_int32), ("n_slots", C.c_int32),
                ("flags", C.c_int32), ("reserved", C.c_int32),
                ("unary_base", C.c_int64), ("q_base", C.c_int64)]


class FgMaxSumDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("precision", C.c_int32),
                ("n_vars", C.c_int32), ("n_factors", C.c_int32), ("n_edges", C.c_int32),
                ("n_classes", C.c_int32), ("n_msg_r", C.c_int64), ("n_msg_q", C.c_int64),
                ("uniform_dom", C.c_int32), ("max_degree", C.c_int32),
                ("classes", C.POINTER(FgClass)),
                ("n_varclasses", C.c_int32), ("reserved0", C.c_int32),
                ("varclasses", C.POINTER(FgVarClass)),
                ("dev_tables", P), ("dev_unary", P), ("dev_dom_size", P), ("dev_unary_off", P),
                ("dev_var_ptr", P), ("dev_var_qbase", P), ("dev_slot_roff", P), ("dev_edge_qoff", P),
                ("dev_slot_roff32", P), ("dev_edge_qoff32", P),
                ("dev_slot_edge", P), ("dev_slot_var", P),
                ("dev_init_value", P),
                ("dev_q", P * 2), ("dev_r", P * 2), ("dev_q_valid", P), ("dev_r_valid", P),
                ("dev_q_cnt", P), ("dev_r_cnt", P), ("dev_q_sent", P), ("dev_r_sent", P),
                ("dev_value", P), ("dev_value_cost", P),
                ("mode_max", C.c_int32), ("damp_vars", C.c_int32), ("damp_factors", C.c_int32),
                ("start_messages", C.c_int32), ("damping", C.c_double), ("stability", C.c_double)]


class FgDsaDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("precision", C.c_int32),
                ("n_vars", C.c_int32), ("n_factors", C.c_int32), ("n_edges", C.c_int32),
                ("n_classes", C.c_int32), ("classes", C.POINTER(FgClass)),
                ("dev_tables", P), ("dev_dom_size", P), ("dev_var_id", P), ("dev_edge_var", P),
                ("dev_edge_class", P),
                ("dev_var_ptr", P), ("dev_slot_edge", P), ("dev_has_nbr", P), ("dev_prob", P),
                ("dev_con_opt", P), ("dev_tables_or", P), ("dev_slot_nbr", P), ("dev_slot_tab", P),
                ("dev_slot_opt", P), ("fast_dom", C.c_int32), ("reserved1", C.c_int32),
                ("dev_value", P * 2), ("dev_value_cost", P),
                ("mode_max", C.c_int32), ("variant", C.c_int32), ("stop_cycle", C.c_int32),
                ("seed", C.c_uint64), ("dev_var_cost", P), ("dev_unary_off", P),
                ("dev_row_cache", P), ("dev_slot_last", P)]


class FgMgmDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("precision", C.c_int32),
                ("n_vars", C.c_int32), ("n_factors", C.c_int32), ("n_edges", C.c_int32),
                ("n_classes", C.c_int32), ("classes", C.POINTER(FgClass)),
                ("dev_tables", P), ("dev_unary", P), ("dev_unary_off", P), ("dev_dom_size", P),
                ("dev_var_id", P), ("dev_var_rank", P), ("dev_edge_var", P), ("dev_edge_class", P),
                ("dev_var_ptr", P), ("dev_slot_edge", P), ("dev_nbr_ptr", P), ("dev_nbr_idx", P),
                ("dev_init_value", P), ("dev_value", P), ("dev_cost", P), ("dev_has_cost", P),
                ("dev_gain", P), ("dev_new_value", P),
                ("mode_max", C.c_int32), ("stop_cycle", C.c_int32), ("seed", C.c_uint64),
                ("dev_tables_or", P), ("dev_slot_nbr", P), ("dev_slot_tab", P),
                ("fast_dom", C.c_int32), ("fast_chunk", C.c_int32),
                ("dev_row_cache", P), ("dev_slot_last", P)]


FG_MAX_PEERS = 16


class FgPeerSync(C.Structure):
    _fields_ = [("n_peers", C.c_int32), ("my_rank", C.c_int32), ("dev_flags", P),
                ("peer_rank", C.c_int32 * FG_MAX_PEERS), ("peer_slot", P * FG_MAX_PEERS),
                ("dev_error", P), ("timeout_ns", C.c_uint64)]


class FgHaloPlan(C.Structure):
    _fields_ = [("elem_bytes", C.c_int32), ("dom", C.c_int32), ("n_r", C.c_int64), ("n_q", C.c_int64),
                ("dev_src_r_off", P), ("dev_src_q_off", P), ("dev_dst_r", P * 2), ("dev_dst_q", P * 2),
                ("dev_counter", P), ("sync", FgPeerSync),
                ("dev_runs_r", P * 2), ("dev_runs_q", P * 2), ("n_runs_r", C.c_int32 * 2), ("n_runs_q", C.c_int32 * 2),
                ("units_r", C.c_int64 * 2), ("units_q", C.c_int64 * 2),
                ("dev_edge_dst_r", P * 2), ("dev_slot_dst_q", P * 2)]


# every symbol include/pydcop_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "fg_abi_version": (C.c_int, []),
    "fg_device_count": (C.c_int, []),
    "fg_maxsum_create": (C.c_int, [C.POINTER(FgMaxSumDesc), C.POINTER(P)]),
    "fg_maxsum_destroy": (C.c_int, [P]),
    "fg_maxsum_last_error": (C.c_char_p, [P]),
    "fg_maxsum_init": (C.c_int, [P, P]),
    "fg_maxsum_step": (C.c_int, [P, C.c_int32, P]),
    "fg_maxsum_cycle_compute": (C.c_int, [P, P]),
    "fg_maxsum_cycle_commit": (C.c_int, [P]),
    "fg_maxsum_current": (C.c_int, [P, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "fg_maxsum_launch_count": (C.c_int64, [P]),
    "fg_maxsum_kernel_plan": (C.c_int, [P, P, C.c_int32]),
    "fg_maxsum_shard_fused": (C.c_int, [P]),
    "fg_halo_pack": (C.c_int, [C.c_int32, P, P, P, P, P, C.c_int64, P]),
    "fg_halo_unpack": (C.c_int, [C.c_int32, P, P, P, P, P, C.c_int64, P]),
    "fg_halo_rows_uniform": (C.c_int, [C.c_int32, C.c_int32, P, P, P, P, P, C.c_int64, C.c_int64,
                                       C.c_int32, C.c_int32, P, P, P, P]),
    "fg_enable_peer_access": (C.c_int, [C.c_int32]),
    "fg_ipc_export": (C.c_int, [P, P, C.POINTER(C.c_int64)]),
    "fg_ipc_import": (C.c_int, [P, C.POINTER(P)]),
    "fg_ipc_close": (C.c_int, [P]),
    "fg_halo_push": (C.c_int, [C.c_int32, P, P, P, P, P, P, C.c_int64, C.c_int64, C.c_int32, P]),
    "fg_peer_signal": (C.c_int, [C.POINTER(FgPeerSync), C.c_uint64, P]),
    "fg_peer_wait": (C.c_int, [C.POINTER(FgPeerSync), C.c_uint64, P]),
    "fg_maxsum_shard_attach": (C.c_int, [P, C.POINTER(FgHaloPlan)]),
    "fg_maxsum_shard_step": (C.c_int, [P, C.c_int32, P]),
    "fg_maxsum_shard_phase": (C.c_int, [P, C.c_int32, P]),
    "fg_maxsum_shard_profile": (C.c_int, [P, C.c_int32, P, C.POINTER(C.c_double)]),
    "fg_dsa_shard_attach": (C.c_int, [P, C.POINTER(FgHaloPlan)]),
    "fg_dsa_shard_step": (C.c_int, [P, C.c_int32, P]),
    "fg_dsa_create": (C.c_int, [C.POINTER(FgDsaDesc), C.POINTER(P)]),
    "fg_dsa_destroy": (C.c_int, [P]),
    "fg_dsa_last_error": (C.c_char_p, [P]),
    "fg_dsa_init": (C.c_int, [P, P]),
    "fg_dsa_step": (C.c_int, [P, C.c_int32, P]),
    "fg_dsa_cycle_compute": (C.c_int, [P, P]),
    "fg_dsa_cycle_commit": (C.c_int, [P]),
    "fg_dsa_current": (C.c_int, [P, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "fg_dsa_launch_count": (C.c_int64, [P]),
    "fg_selftest_approx_match": (C.c_int, [C.c_int32, C.c_int64, P, P, C.c_double, P, P, P]),
    "fg_selftest_gather": (C.c_int, [P, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_int32, P, P]),
    "fg_mgm_create": (C.c_int, [C.POINTER(FgMgmDesc), C.POINTER(P)]),
    "fg_mgm_destroy": (C.c_int, [P]),
    "fg_mgm_last_error": (C.c_char_p, [P]),
    "fg_mgm_init": (C.c_int, [P, P]),
    "fg_mgm_step": (C.c_int, [P, C.c_int32, P]),
    "fg_mgm_current": (C.c_int, [P, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "fg_mgm_launch_count": (C.c_int64, [P]),
    "fg_solution_cost": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(FgClass), P, P, P, P, P,
                                   C.c_int32, P, P, C.c_double, P, P]),
}

_lib = None


class EngineError(RuntimeError):
    pass


def load():
    """Load libpydcop_b200.so and bind every declared symbol; raise if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `python -m pydcop_b200.build` "
            "(pydcop_b200 has no CPU fallback for the MaxSum/DSA hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.fg_abi_version() != FG_ABI_VERSION:
        raise EngineError("libpydcop_b200.so ABI version mismatch: rebuild it")
    _lib = lib
    return lib
