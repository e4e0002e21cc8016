"""ctypes binding of libhoic_b200.so (include/hoic.h).  There is no CPU fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HOIC_LIB", os.path.join(_HERE, "libhoic_b200.so"))   # HOIC_LIB: alternate build for A/B timing
_lib = None

OBS_DIM, ACT_DIM, FRAME_DIM, INFO_DIM, NQ, NV, MAX_CONTACTS = 617, 32, 212, 9, 33, 32, 32
NHG = 19
REC_MAXCON, REC_CON, REC_STRIDE, SURF_MAXPT, SURF_PT, SURF_STRIDE = 32, 116, 116 + 19 * 32, 95, 8, 8 + 15 * 95 + 3

EXPORTS = [
    "hoic_last_error", "hoic_version", "hoic_model_create", "hoic_model_destroy", "hoic_batch_create",
    "hoic_batch_destroy", "hoic_batch_size", "hoic_set_motions", "hoic_set_reward_weights", "hoic_set_mode",
    "hoic_reset", "hoic_step", "hoic_get_state", "hoic_set_state", "hoic_set_cur_t", "hoic_get_warmstart",
    "hoic_set_warmstart", "hoic_get_contacts", "hoic_substep_debug", "hoic_get_derived", "hoic_get_rfc",
    "hoic_get_counters", "hoic_launch_count", "hoic_measure_fp32_peak", "hoic_get_work", "hoic_set_env_order", "hoic_set_balance", "hoic_set_sampler", "hoic_reset_sampled", "hoic_get_episode", "hoic_set_timing", "hoic_get_timing", "hoic_zfilter", "hoic_zfilter16", "hoic_sample_action", "hoic_logger_step", "hoic_set_episode", "hoic_push_frame", "hoic_glue_last_error", "hoic_set_record", "hoic_preprocess_seq", "hoic_solve_wrench_qp",
]


class Counters(C.Structure):
    _fields_ = [(n, C.c_ulonglong) for n in (
        "env_steps", "substeps", "ncon", "nefc", "newton_iters", "ls_evals", "mpr_calls", "mpr_iters", "qp_iters",
        "fail_physics", "fail_overflow", "chol_calls", "max_ncon", "max_nbasis", "max_nedge")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class HoicError(RuntimeError):
    pass


def lib():
    """Load the CUDA library; raises if it has not been built (python -m hoic_b200.build / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HoicError(f"{LIB_PATH} not found: build it with `python -m hoic_b200.build` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        vp, ip, sz = C.c_void_p, C.c_int, C.c_size_t
        L.hoic_last_error.restype = C.c_char_p
        L.hoic_version.restype = ip
        L.hoic_model_create.restype = vp
        L.hoic_model_create.argtypes = [C.c_char_p, sz, ip]
        L.hoic_model_destroy.argtypes = [vp]
        L.hoic_batch_create.restype = vp
        L.hoic_batch_create.argtypes = [C.POINTER(vp), ip, vp, ip, ip]
        L.hoic_batch_destroy.argtypes = [vp]
        L.hoic_batch_size.argtypes = [vp]
        L.hoic_set_motions.argtypes = [vp, vp, vp, vp, ip]
        L.hoic_set_reward_weights.argtypes = [vp, vp]
        L.hoic_set_mode.argtypes = [vp, ip]
        L.hoic_set_record.argtypes = [vp, vp, vp]
        L.hoic_solve_wrench_qp.argtypes = [ip, ip, ip, ip, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, vp, vp, vp]
        L.hoic_preprocess_seq.argtypes = [vp, vp, vp, ip, C.c_float, ip, vp, vp, vp, vp]
        L.hoic_reset.argtypes = [vp, vp, vp, vp, vp, vp]
        L.hoic_step.argtypes = [vp] + [vp] * 8 + [C.c_int, vp]
        L.hoic_set_sampler.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_ulonglong]
        L.hoic_reset_sampled.argtypes = [vp, vp, vp, vp]
        L.hoic_get_episode.argtypes = [vp, vp, vp, vp, vp]
        L.hoic_set_balance.argtypes = [vp, C.c_int]
        L.hoic_set_timing.argtypes = [vp, C.c_int]
        L.hoic_set_episode.argtypes = [vp, vp, vp, vp, vp]
        L.hoic_push_frame.argtypes = [vp, vp, C.c_int, vp]
        L.hoic_zfilter.argtypes = [vp, ip, ip] + [vp] * 8 + [C.c_float, ip, vp, vp]
        L.hoic_logger_step.argtypes = [vp, vp, vp, ip, vp, vp, vp, vp]
        L.hoic_zfilter16.argtypes = [vp, ip, ip] + [vp] * 8 + [C.c_float, ip, vp, vp, ip, vp]
        L.hoic_sample_action.argtypes = [vp, ip, vp, vp, vp, vp, ip, ip, vp, vp, vp]
        L.hoic_glue_last_error.restype = C.c_char_p
        L.hoic_get_timing.argtypes = [vp, vp, vp]
        L.hoic_get_state.argtypes = [vp, vp, vp, vp]
        L.hoic_set_state.argtypes = [vp, vp, vp, vp]
        L.hoic_set_cur_t.argtypes = [vp, vp, vp]
        L.hoic_get_warmstart.argtypes = [vp, vp, vp]
        L.hoic_set_warmstart.argtypes = [vp, vp, vp]
        L.hoic_get_contacts.argtypes = [vp, vp, vp, vp, vp]
        L.hoic_substep_debug.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.hoic_get_derived.argtypes = [vp, vp, vp, vp, vp, vp]
        L.hoic_get_rfc.argtypes = [vp, vp, vp]
        L.hoic_get_counters.argtypes = [vp, C.POINTER(Counters)]
        L.hoic_launch_count.restype = C.c_ulonglong
        L.hoic_get_work.argtypes = [vp, vp, vp]
        L.hoic_set_env_order.argtypes = [vp, vp]
        L.hoic_measure_fp32_peak.argtypes = [ip, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise HoicError(f"{what}: {lib().hoic_last_error().decode()}")


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)
