"""ctypes binding of include/tal_b200.h (libtal_b200.so).

There is no CPU fallback: importing this module without the built library, or
calling into it on a machine whose GPU is not a B200-class (CC 10.x) device,
fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libtal_b200.so")

# mirrors of the header's constants
TAL_OK = 0
TAL_MAX_Q = 16
TAL_PEER_MAX = 16
TAL_MAX_PENDING = 32
TAL_PEER_HANDLE_BYTES = 64
F64, F32 = 0, 1
KERNEL_GAUSSIAN, KERNEL_LAPLACE, KERNEL_MATERN32, KERNEL_MATERN52, KERNEL_LINEAR = range(5)
RULE_VTL, RULE_CTL, RULE_MMITL, RULE_TRUVAR = range(4)
ARGMAX_OK, ARGMAX_EMPTY_SAFE_SET, ARGMAX_ALL_NEG_INF, ARGMAX_ALL_NAN, ARGMAX_PEER_ERROR = 0, 2, 3, 4, 7
ARITH_FMA = 16


class ArgmaxResult(C.Structure):
    _fields_ = [
        ("idx", C.c_longlong),
        ("val", C.c_double),
        ("n_ties", C.c_longlong),
        ("status", C.c_int),
        ("flags", C.c_int),
    ]


TAL_CTX_REBASE = 6
TAL_CTX_MAX_EVENTS = 64


class StepCtx(C.Structure):
    """tal_step_ctx of include/tal_b200.h (every field 8 bytes wide)."""
    _fields_ = (
        [("cov", C.c_void_p)] + [(n, C.c_longlong) for n in ("cov_stride_q", "ld", "row0", "n_loc", "N", "q", "dtype", "lower")]
        + [(n, C.c_void_p) for n in ("mean", "var", "l", "u", "rho", "scal", "kbuf", "wbuf")]
        + [("beta", C.c_double * 16)]
        + [("Kp", C.c_void_p), ("Wp", C.c_void_p)]
        + [(n, C.c_longlong) for n in ("pend_stride", "pend_slots", "update_block", "pending", "passes")]
        + [("Wb", C.c_void_p)]
        + [(n, C.c_longlong) for n in ("ldw", "cap_rows", "m_pad", "rows", "ncols", "cand0", "n_cand", "n_split", "cond_valid")]
        + [(n, C.c_void_p) for n in ("cs", "app_scratch", "pcol")]
        + [("F", C.c_void_p), ("safe", C.c_void_p), ("sample", C.c_void_p), ("first_constraint", C.c_longlong),
           ("argmax_scratch", C.c_void_p), ("rec", C.c_void_p), ("combined", C.c_void_p)]
        + [("y_table", C.c_void_p), ("y_stride", C.c_longlong)]
        + [("mailboxes", C.c_void_p), ("mailbox", C.c_void_p), ("G", C.c_longlong), ("rank", C.c_longlong)]
        + [("bounds", C.c_longlong * 17)]
        + [("kbuf_off", C.c_longlong), ("pcol_off", C.c_longlong), ("pcol_peer", C.c_void_p), ("pcol_stride", C.c_longlong)]
        + [("ev_start", C.c_void_p * 64), ("ev_stop", C.c_void_p * 64), ("n_ev", C.c_longlong), ("timing", C.c_longlong)]
        + [("ev_updates", C.c_longlong * 64)]
        + [(n, C.c_longlong) for n in ("arith", "auto_select", "sel_state", "scal_ok")]
        + [("sel_safe", C.c_void_p), ("sel_sample", C.c_void_p), ("sel_fc", C.c_longlong), ("fz_scratch", C.c_void_p)]
        + [("side_stream", C.c_void_p), ("ev_fork", C.c_void_p), ("ev_join", C.c_void_p), ("overlap", C.c_longlong)]
    )


class TalError(RuntimeError):
    pass


_p = C.c_void_p
_ll = C.c_longlong
_i = C.c_int
_d = C.c_double

# name -> (restype, argtypes); every symbol include/tal_b200.h declares
SIGNATURES = {
    "tal_abi_version": (_i, []),
    "tal_last_error_string": (C.c_char_p, []),
    "tal_launch_count": (_ll, []),
    "tal_reset_launch_count": (None, []),
    "tal_device_ok": (_i, []),
    "tal_gram_stationary": (_i, [_i, _p, _ll, _p, _ll, _i, _d, _d, _p, _ll, _ll, _ll, _i, _p]),
    "tal_gram_stationary_lower": (_i, [_i, _p, _ll, _i, _d, _d, _p, _ll, _ll, _ll, _i, _p]),
    "tal_nearest_index": (_i, [_p, _ll, _p, _ll, _i, _p, _p]),
    "tal_extract_row": (_i, [_p, _ll, _ll, _ll, _ll, _ll, _i, _p, _ll, _p, _p, _p, _i, _p]),
    "tal_step_vectors": (_i, [_p, _p, _ll, _i, _p, _ll, _p, _ll, _i, _p, C.POINTER(_d), _p,
                              _p, _p, _p, _p, _i, _p]),
    "tal_rank1_update": (_i, [_p, _ll, _ll, _ll, _ll, _ll, _i, _p, _p, _i, _i, _p]),
    "tal_gather_rows": (_i, [_p, _ll, _ll, _ll, _ll, _p, _ll, _p, _ll, _i, _p]),
    "tal_small_posterior_weights": (_i, [_p, _ll, _ll, _p, _i, _p, _d, _p, _p, _p, _p, _p, _p]),
    "tal_small_scratch_bytes": (_ll, []),
    "tal_rankm_update": (_i, [_p, _ll, _ll, _ll, _ll, _p, _p, _ll, _i, _i, _p]),
    "tal_bounds_masks": (_i, [_p, _p, _i, _i, _ll, _p, _p, _p, _p, _p, _p, _p]),
    "tal_mask_to_indices": (_i, [_p, _ll, _p, _p, _p]),
    "tal_scalar_subset": (_i, [_p, _i, _ll, _ll, _p, _p, _p, _ll, _p, _p, _p]),
    "tal_score_itl_undirected": (_i, [_p, _p, _i, _ll, _p, _p]),
    "tal_colsum_rows": (_i, [_i, _p, _ll, _ll, _ll, _ll, _p, _ll, _p, _p, _d, _d, _p, _i, _p,
                             _i, _p]),
    "tal_colsum_lower": (_i, [_i, _p, _ll, _ll, _ll, _ll, _p, _ll, _p, _p, _d, _d, _p, _i, _p,
                              _i, _p]),
    "tal_colsum_lower_dense": (_i, [_i, _p, _ll, _ll, _ll, _ll, _p, _ll, _p, _p, _p, _p, _d, _d, _p, _i, _p,
                                    _i, _p]),
    "tal_update_colsum_scratch_bytes": (_ll, [_ll, _ll, _i]),
    "tal_update_colsum_lower_dense": (_i, [_i, _p, _ll, _ll, _ll, _ll, _p, _p, _p, _p, _p, _d, _d, _p, _i, _p, _p,
                                           _i, _p]),
    "tal_colsum_cols": (_i, [_i, _p, _ll, _ll, _p, _ll, _p, _p, _d, _d, _p, _i, _p]),
    "tal_roi_trace": (_i, [_p, _p, _ll, _p, _p]),
    "tal_roi_weights": (_i, [_p, _p, _ll, _i, _p, _p]),
    "tal_pred_var_inv": (_i, [_p, _p, _ll, _p, _p]),
    "tal_score_rows_finish": (_i, [_i, _p, _p, _p, _p, _ll, _i, _p, _p]),
    "tal_gather_roi": (_i, [_p, _ll, _ll, _p, _ll, _ll, _d, _p, _ll, _p, _ll, _i, _p]),
    "tal_gather_roi_lower": (_i, [_p, _ll, _ll, _p, _ll, _ll, _d, _p, _ll, _p, _ll, _i, _p]),
    "tal_gather_roi_sharded": (_i, [_p, _ll, _ll, _ll, _p, _ll, _ll, _d, _p, _ll, _p, _ll, _ll, _i, _p]),
    "tal_potrf_work_bytes": (_ll, [_ll]),
    "tal_potrf_lower": (_i, [_p, _ll, _ll, _p, _p]),
    "tal_trsm_lower": (_i, [_p, _ll, _ll, _p, _p, _ll, _ll, _p]),
    "tal_colsumsq_dense": (_i, [_p, _ll, _ll, _ll, _p, _i, _p, _p]),
    "tal_score_itl_directed": (_i, [_p, _p, _p, _ll, _i, _p, _p]),
    "tal_cond_extract_col": (_i, [_p, _ll, _ll, _ll, _ll, _p, _ll, _p, _p]),
    "tal_cond_update": (_i, [_p, _ll, _ll, _ll, _ll, _ll, _p, _p, _p, _p, _p, _ll, _p, _p, _i, _p,
                             _i, _p]),
    "tal_cond_append_update": (_i, [_p, _ll, _ll, _ll, _ll, _ll, _ll, _ll, _p, _p, _p, _p, _p, _ll, _p, _p, _i,
                                    _i, _p]),
    "tal_cond_append_scratch_bytes": (_ll, [_ll, _i]),
    "tal_cond_coef_bytes": (_ll, [_ll]),
    "tal_cond_chunk_scratch_bytes": (_ll, [_ll]),
    "tal_argmax_scratch_bytes": (_ll, []),
    "tal_masked_argmax": (_i, [_p, _p, _p, _i, _i, _ll, _p, _ll, _ll, _p, _p, _p]),
    "tal_argmax_combine": (_i, [_p, _i, _p, _p]),
    "tal_peer_header_bytes": (_ll, []),
    "tal_peer_alloc": (_i, [_ll, C.POINTER(_p)]),
    "tal_peer_free": (_i, [_p]),
    "tal_peer_export": (_i, [_p, C.c_char_p]),
    "tal_peer_import": (_i, [C.c_char_p, C.POINTER(_p)]),
    "tal_peer_unimport": (_i, [_p]),
    "tal_peer_enable_access": (_i, [_i]),
    "tal_peer_publish_record": (_i, [_p, _p, _i, _i, _p]),
    "tal_peer_combine_records": (_i, [_p, _i, _p, _p]),
    "tal_peer_push_row": (_i, [_p, _i, _i, C.POINTER(_ll), _p, _ll, _ll, _ll, _i, _i, _p, _ll, _ll, _p, _ll,
                               _ll, _ll, _ll, _ll, _ll, _ll, _i, _p]),
    "tal_peer_wait_row": (_i, [_p, _i, _p, _ll, _p, _ll, _i, _p, _p]),
    "tal_sym_block": (_ll, [_i]),
    "tal_rankb_update": (_i, [_p, _ll, _ll, _ll, _ll, _ll, _i, _p, _p, _ll, _i, _i, _i, _p]),
    "tal_pending_correct_row": (_i, [_p, _ll, _ll, _i, _p, _p, _ll, _i, _p, _ll, _i, _i, _p]),
    "tal_rank1_update_lower": (_i, [_p, _ll, _ll, _ll, _ll, _ll, _i, _p, _p, _i, _p]),
    "tal_extract_row_lower": (_i, [_p, _ll, _ll, _ll, _ll, _ll, _i, _p, _ll, _p, _p, _p, _i, _p]),
    "tal_gather_rows_lower": (_i, [_p, _ll, _ll, _ll, _ll, _p, _ll, _p, _ll, _i, _p]),
    "tal_rankm_update_lower": (_i, [_p, _ll, _ll, _ll, _ll, _p, _p, _ll, _i, _i, _p]),
    "tal_sym_mirror": (_i, [_p, _ll, _ll, _ll, _i, _i, _p]),
    "tal_peer_error": (_i, [_p, C.POINTER(_i), _p]),
    "tal_step_ctx_bytes": (_ll, []),
    "tal_ctx_breakdown": (_i, [C.POINTER(_d), C.POINTER(_ll)]),
    "tal_ctx_select": (_i, [_p, _p]),
    "tal_ctx_observe": (_i, [_p, _p, _ll, _p, _ll, _i, _p]),
    "tal_ctx_step": (_i, [_p, _p]),
    "tal_ctx_flush": (_i, [_p, _p]),
    "tal_ctx_timing_begin": (_i, [_p]),
    "tal_ctx_timing_read": (_i, [_p, C.POINTER(_d), C.POINTER(_ll), C.POINTER(_ll), _p]),
    "tal_ctx_observe_y": (_i, [_p, _p, _ll, _p, _ll, _i, _p]),
    "tal_fused_scratch_bytes": (_ll, [_i, _ll, _i]),
    "tal_probe_fp64": (_i, [_i, _i, _i, _p, _p]),
    "tal_gather_sym_block": (_i, [_p, _ll, _ll, _p, _ll, _ll, _p, _d, _p, _ll, _i, _i, _p]),
    "tal_chol_logdet": (_i, [_p, _ll, _ll, _p, _p]),
    "tal_tri_sample": (_i, [_p, _ll, _ll, _p, _ll, _i, _p, _p, _ll, _p]),
    "tal_gather_cols_t": (_i, [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _p]),
    "tal_syrk_lower_sub": (_i, [_p, _ll, _ll, _p, _ll, _ll, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libtal_b200.so and declare every signature.  Raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TalError(
            f"{LIB_PATH} is missing: build it with "
            "`python transductive-active-learning_b200/build.py` (no CPU fallback exists)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().tal_last_error_string().decode()


def check(rc: int, what: str = "") -> None:
    if rc != TAL_OK:
        raise TalError(f"{what} failed with status {rc}: {last_error()}")


def require_device() -> None:
    """Fail loudly unless a B200-class GPU is usable."""
    import torch

    if not torch.cuda.is_available():
        raise TalError("CUDA device required: the discrete-GP path has no CPU fallback")
    if not load().tal_device_ok():
        raise TalError(last_error())
