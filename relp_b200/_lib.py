"""ctypes binding of librelp_gpu.so (include/relp_gpu.h, include/relp_host.h).

The product path has no CPU fallback: if the CUDA library is missing, importing fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librelp_gpu.so")


class rg_options(C.Structure):
    _fields_ = [("device", C.c_int32), ("initial_limbs", C.c_int32), ("rank", C.c_int32),
                ("world", C.c_int32), ("dense_carry", C.c_int32), ("reserved", C.c_int32),
                ("nccl_unique_id", C.c_void_p)]


RG_NWIDTHS = 8     # include/relp_gpu.h: the limb-width ladder 1, 2, 4, 8, 10, 12, 14, 16


class rg_stats(C.Structure):
    _fields_ = [("pivots", C.c_int64), ("promotions", C.c_int64), ("limbs", C.c_int32),
                ("max_bits", C.c_int32), ("denominator_bits", C.c_int32), ("reserved", C.c_int32),
                ("kernel_launches", C.c_int64), ("demotions", C.c_int64),
                ("limb_widths", C.c_int32 * RG_NWIDTHS), ("pivots_at_limbs", C.c_int64 * RG_NWIDTHS),
                ("k1_launches_at_limbs", C.c_int64 * RG_NWIDTHS), ("k1_ms_at_limbs", C.c_double * RG_NWIDTHS),
                ("timer_ms", C.c_double), ("phase_ms", C.c_double * 8),
                ("k1_bytes_at_limbs", C.c_double * RG_NWIDTHS), ("k1_imads_at_limbs", C.c_double * RG_NWIDTHS)]


class rg_pivot_info(C.Structure):
    _fields_ = [("status", C.c_int32), ("entering", C.c_int32), ("row", C.c_int32),
                ("leaving", C.c_int32)]


class rh_problem(C.Structure):
    _fields_ = [("m", C.c_int32), ("n", C.c_int32),
                ("colptr", C.POINTER(C.c_int64)), ("rowidx", C.POINTER(C.c_int32)),
                ("vals", C.POINTER(C.c_int64)), ("cost", C.POINTER(C.c_int64)),
                ("rhs", C.POINTER(C.c_int64)), ("n_pivots", C.c_int32),
                ("pivot_rows", C.POINTER(C.c_int32)), ("pivot_cols", C.POINTER(C.c_int32)),
                ("full_initial_basis", C.c_int32),
                ("colfac", C.POINTER(C.c_int64)), ("artfac", C.POINTER(C.c_int64)),
                ("colw", C.POINTER(C.c_int64)), ("artcost", C.POINTER(C.c_int64)),
                ("n_dense", C.c_int32), ("reserved0", C.c_int32), ("dense", C.POINTER(C.c_int8))]


class rh_trace_entry(C.Structure):
    _fields_ = [("phase", C.c_int32), ("entering", C.c_int32), ("row", C.c_int32),
                ("leaving", C.c_int32)]


class rh_options(C.Structure):
    _fields_ = [("device", C.c_int32), ("initial_limbs", C.c_int32), ("rule", C.c_int32),
                ("fused", C.c_int32), ("max_pivots", C.c_int64), ("profile", C.c_int32),
                ("rank", C.c_int32), ("world", C.c_int32), ("dense_carry", C.c_int32),
                ("nccl_unique_id", C.c_void_p)]


# every symbol include/*.h declares: name -> (restype, argtypes)
P = C.c_void_p
SYMBOLS = {
    "rg_create": (C.c_int, [C.POINTER(rg_options), C.POINTER(P)]),
    "rg_destroy": (C.c_int, [P]),
    "rg_last_error": (C.c_char_p, [P]),
    "rg_nccl_unique_id": (C.c_int, [C.c_void_p, C.c_int32]),
    "rg_shard_block": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rg_release_cached_memory": (C.c_int64, [C.c_int32]),
    "rg_load_csc": (C.c_int, [P, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                              C.POINTER(C.c_int64)]),
    "rg_load_dense_i8": (C.c_int, [P, C.c_int32, C.POINTER(C.c_int8)]),
    "rg_set_rhs": (C.c_int, [P, C.POINTER(C.c_int64)]),
    "rg_set_weights": (C.c_int, [P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int64)]),
    "rg_init_identity_basis": (C.c_int, [P, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "rg_init_basis": (C.c_int, [P, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "rg_phase_switch": (C.c_int, [P, C.POINTER(C.c_int64)]),
    "rg_rule_new": (C.c_int, [P, C.c_int32]),
    "rg_select_primal_pivot_column": (C.c_int, [P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rg_generate_column": (C.c_int, [P, C.c_int32]),
    "rg_select_primal_pivot_row": (C.c_int, [P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rg_bring_into_basis": (C.c_int, [P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(rg_pivot_info)]),
    "rg_iterate": (C.c_int, [P, C.c_int64, C.POINTER(rg_pivot_info), C.POINTER(C.c_int64),
                             C.POINTER(C.c_int32)]),
    "rg_remove_artificial_row": (C.c_int, [P, C.c_int32, C.POINTER(rg_pivot_info)]),
    "rg_get_limbs": (C.c_int, [P, C.POINTER(C.c_int32)]),
    "rg_get_denominator": (C.c_int, [P, C.POINTER(C.c_uint64)]),
    "rg_get_basis": (C.c_int, [P, C.POINTER(C.c_int32)]),
    "rg_get_b": (C.c_int, [P, C.POINTER(C.c_uint64)]),
    "rg_get_minus_objective": (C.c_int, [P, C.POINTER(C.c_uint64)]),
    "rg_get_minus_pi": (C.c_int, [P, C.POINTER(C.c_uint64)]),
    "rg_get_basis_inverse_row": (C.c_int, [P, C.c_int32, C.POINTER(C.c_uint64)]),
    "rg_get_pivot_column": (C.c_int, [P, C.POINTER(C.c_uint64)]),
    "rg_get_relative_costs": (C.c_int, [P, C.POINTER(C.c_uint64)]),
    "rg_get_gamma": (C.c_int, [P, C.POINTER(C.c_uint64)]),
    "rg_get_element": (C.c_int, [P, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]),
    "rg_get_basis_change_info": (C.c_int, [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_uint64)]),
    "rg_get_stats": (C.c_int, [P, C.POINTER(rg_stats)]),
    "rg_set_profile": (C.c_int, [P, C.c_int32]),
    "rg_timer_start": (C.c_int, [P]),
    "rg_timer_stop": (C.c_int, [P]),
    "rg_debug_scalars": (C.c_int, [P, C.c_void_p, C.c_int64]),
    "rg_debug_vector": (C.c_int, [P, C.c_int32, C.POINTER(C.c_uint64)]),
    "rg_selftest": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                              C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int64,
                              C.POINTER(C.c_uint64)]),
    "rg_measure_imad_peak": (C.c_int, [C.c_int32, C.c_double, C.POINTER(C.c_double)]),
    "rh_solve_relaxation": (C.c_int, [C.POINTER(rh_problem), C.POINTER(rh_options), C.POINTER(P)]),
    "rh_result_free": (None, [P]),
    "rh_result_error": (C.c_char_p, [P]),
    "rh_result_status": (C.c_int32, [P]),
    "rh_result_pivots": (C.c_int64, [P]),
    "rh_result_trace_len": (C.c_int64, [P]),
    "rh_result_trace": (C.POINTER(rh_trace_entry), [P]),
    "rh_result_limbs": (C.c_int32, [P]),
    "rh_result_minus_objective": (C.POINTER(C.c_uint64), [P]),
    "rh_result_denominator": (C.POINTER(C.c_uint64), [P]),
    "rh_result_basis": (C.POINTER(C.c_int32), [P]),
    "rh_result_b": (C.POINTER(C.c_uint64), [P]),
    "rh_result_nr_artificial": (C.c_int32, [P]),
    "rh_result_rows_removed_len": (C.c_int32, [P]),
    "rh_result_rows_removed": (C.POINTER(C.c_int32), [P]),
    "rh_result_stats": (None, [P, C.POINTER(rg_stats)]),
    "rh_result_seconds": (C.c_double, [P]),
    "rh_result_device_ms": (C.c_double, [P]),
    "rh_result_seconds_total": (C.c_double, [P]),
}

_lib = None


def load():
    """Loads the shared library; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m relp_b200.build` "
            "(the engine has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
