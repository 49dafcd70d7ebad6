"""ctypes binding of libklara_b200.so (include/klara_b200.h).

There is deliberately no fallback: if the CUDA library has not been built, or no GPU is
present, every compute call raises.  Build with `python klara.jl_b200/build.py`
(or `__graft_entry__.build()`).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# KLB_LIB_PATH selects an experimental build variant (klara.jl_b200/build.py KLB_VARIANT=...)
LIB_PATH = os.environ.get("KLB_LIB_PATH") or os.path.join(HERE, "lib", "libklara_b200.so")

# error codes / enums (mirror include/klara_b200.h)
KLB_OK, KLB_EINVAL, KLB_ECUDA, KLB_ENOTFINITE, KLB_ESTATE, KLB_EUNSUPPORTED, KLB_ENOMEM = 0, -1, -2, -3, -4, -5, -6
SAMPLER_MH, SAMPLER_MALA, SAMPLER_HMC, SAMPLER_NUTS = 0, 1, 2, 3
TARGET_ISO, TARGET_SHIFTED_ISO, TARGET_DENSE, TARGET_ROSENBROCK, TARGET_LOGIT = 0, 1, 2, 3, 4
TUNER_VANILLA, TUNER_ACCEPTANCE_RATE, TUNER_DUAL_AVERAGING = 0, 1, 2
SCORE_LOGISTIC, SCORE_ERF = 0, 1
ARITH_REFERENCE, ARITH_FMA = 0, 1
MONITOR_VALUE, MONITOR_LOGTARGET, MONITOR_GRADLOGTARGET = 1, 2, 4
DIAG_ACCEPT, DIAG_NDOUBLINGS, DIAG_NUTS_A, DIAG_NUTS_NA = 1, 2, 4, 8
DEST_NSTATE, DEST_NONE = 0, 1
PARAM_MU, PARAM_C, PARAM_SIGMA, PARAM_ROSEN, PARAM_LOGIT_X, PARAM_LOGIT_Y, PARAM_LOGIT_LAMBDA = 0, 1, 2, 3, 4, 5, 6
(OUT_VALUE, OUT_LOGTARGET, OUT_GRADLOGTARGET, OUT_ACCEPT, OUT_STATE, OUT_STATE_LOGTARGET,
 OUT_TUNE_STEP, OUT_TUNE_COUNTERS, OUT_TUNE_RATE, OUT_ESS, OUT_TUNE_DA, OUT_TUNE_RATES, OUT_NDOUBLINGS, OUT_NUTS_A,
 OUT_NUTS_NA) = range(15)
PEAK_FP64, PEAK_DMMA = 0, 1
GATHER_HANDLE_BYTES = 128
(STAT_MEAN, STAT_MCVAR_IID, STAT_MCVAR_IMSE, STAT_ESS, STAT_IACT, STAT_ACCEPTANCE, STAT_ACCEPTANCE_VALUE) = range(7)


class KlbConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("sampler", C.c_int32), ("target", C.c_int32), ("tuner", C.c_int32),
        ("arith", C.c_int32),
        ("nchains", C.c_int64), ("dim", C.c_int64), ("nsteps", C.c_int64), ("burnin", C.c_int64),
        ("thinning", C.c_int64),
        ("step", C.c_double), ("nleaps", C.c_int32),
        ("target_rate", C.c_double), ("score_k", C.c_double), ("period", C.c_int64),
        ("verbose", C.c_int32), ("monitor", C.c_uint32), ("diagnostics", C.c_uint32), ("destination", C.c_int32),
        ("seed", C.c_uint64), ("chain_offset", C.c_int64), ("device", C.c_int32), ("score", C.c_int32),
        ("da_nadapt", C.c_int64), ("da_t0", C.c_int64), ("da_eps0bar", C.c_double), ("da_h0bar", C.c_double),
        ("da_gamma", C.c_double), ("da_kappa", C.c_double),
        ("nuts_maxdelta", C.c_int32), ("nuts_maxndoublings", C.c_int32),
    ]


class KlbPlan(C.Structure):
    _fields_ = [("nv", C.c_int32), ("warps_per_block", C.c_int32), ("warps_per_chain", C.c_int32),
                ("regs_per_thread", C.c_int32),
                ("blocks_per_sm", C.c_int32), ("ld", C.c_int64), ("npoststeps", C.c_int64), ("transitions_done", C.c_int64),
                ("saved", C.c_int64)]


class KlbHostField(C.Structure):
    _fields_ = [("field", C.c_int32), ("reserved", C.c_int32), ("host_dst", C.c_void_p), ("nbytes", C.c_int64)]


class KlaraError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("klara_b200 error %d: %s" % (code, msg))
        self.code = code


# every symbol include/klara_b200.h declares: (name, restype, argtypes)
_vp, _i64, _u64, _int, _dbl = C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.c_double
SYMBOLS = [
    ("klb_version", _int, []),
    ("klb_last_error", C.c_char_p, []),
    ("klb_device_count", _int, []),
    ("klb_job_create", _int, [C.POINTER(KlbConfig), C.POINTER(_vp)]),
    ("klb_job_set_target_f64", _int, [_vp, _int, _vp, _i64]),
    ("klb_job_set_state", _int, [_vp, _vp]),
    ("klb_job_set_state_device", _int, [_vp, _vp]),
    ("klb_job_set_state_synthetic", _int, [_vp]),
    ("klb_job_seek", _int, [_vp, _u64]),
    ("klb_job_run", _int, [_vp]),
    ("klb_job_run_async", _int, [_vp]),
    ("klb_job_sync", _int, [_vp]),
    ("klb_job_run_host", _int, [_vp, _vp, C.POINTER(KlbHostField), C.c_int32, C.c_int32]),
    ("klb_job_run_host_async", _int, [_vp, _vp, C.POINTER(KlbHostField), C.c_int32, C.c_int32]),
    ("klb_job_run_host_finish", _int, [_vp]),
    ("klb_job_run_host_wait", _int, [_vp]),
    ("klb_job_run_host_abort", _int, [_vp]),
    ("klb_job_set_chunk", _int, [_vp, _i64]),
    ("klb_job_reset", _int, [_vp]),
    ("klb_job_output", _int, [_vp, _int, _vp, _i64]),
    ("klb_job_device_ptr", _int, [_vp, _int, C.POINTER(_vp), C.POINTER(_i64)]),
    ("klb_job_ess", _int, [_vp, _vp]),
    ("klb_job_stat", _int, [_vp, _int, _vp]),
    ("klb_job_plan", _int, [_vp, C.POINTER(KlbPlan)]),
    ("klb_job_config", _int, [_vp, C.POINTER(KlbConfig)]),
    ("klb_job_launches", _i64, [_vp]),
    ("klb_job_last_run_ms", _dbl, [_vp]),
    ("klb_job_stream", _vp, [_vp]),
    ("klb_job_destroy", None, [_vp]),
    ("klb_multi_create", _int, [C.POINTER(KlbConfig), C.c_int32, _vp, C.POINTER(_vp)]),
    ("klb_multi_ngpus", _int, [_vp]),
    ("klb_multi_job", _int, [_vp, C.c_int32, C.POINTER(_vp)]),
    ("klb_multi_set_target_f64", _int, [_vp, _int, _vp, _i64]),
    ("klb_multi_set_state", _int, [_vp, _vp]),
    ("klb_multi_set_state_synthetic", _int, [_vp]),
    ("klb_multi_reset", _int, [_vp]),
    ("klb_multi_seek", _int, [_vp, _u64]),
    ("klb_multi_run", _int, [_vp]),
    ("klb_multi_run_async", _int, [_vp]),
    ("klb_multi_sync", _int, [_vp]),
    ("klb_multi_run_host", _int, [_vp, _vp, C.POINTER(KlbHostField), C.c_int32, C.c_int32]),
    ("klb_multi_output", _int, [_vp, _int, _vp, _i64]),
    ("klb_multi_gathered", _int, [_vp, C.c_int32, _int, C.POINTER(_vp), C.POINTER(_i64)]),
    ("klb_multi_gathered_output", _int, [_vp, C.c_int32, _int, _vp, _i64]),
    ("klb_multi_last_run_ms", _dbl, [_vp]),
    ("klb_multi_destroy", None, [_vp]),
    ("klb_gather_create", _int, [_vp, C.c_int32, C.c_int32, _i64, _i64, C.POINTER(_vp)]),
    ("klb_gather_handle", _int, [_vp, _vp]),
    ("klb_gather_connect", _int, [_vp, _vp]),
    ("klb_gather_push_async", _int, [_vp]),
    ("klb_gather_sync", _int, [_vp]),
    ("klb_gather_join", _int, [_vp]),
    ("klb_gather_device_ptr", _int, [_vp, _int, C.POINTER(_vp), C.POINTER(_i64)]),
    ("klb_gather_output", _int, [_vp, _int, _vp, _i64]),
    ("klb_gather_disconnect", _int, [_vp]),
    ("klb_gather_destroy", None, [_vp]),
    ("klb_device_peak", _int, [_int, _int, C.POINTER(_dbl)]),
    ("klb_host_alloc", _int, [C.POINTER(_vp), _i64]),
    ("klb_host_free", _int, [_vp]),
    ("klb_debug_normals", _int, [_int, _u64, _u64, _u64, _i64, _vp]),
    ("klb_debug_math", _int, [_int, _int, _i64, _vp, _vp]),
    ("klb_debug_uniform", _int, [_int, _u64, _u64, _u64, _vp]),
]

_lib = None


def lib():
    """Load the shared library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libklara_b200.so is not built (%s): run `python klara.jl_b200/build.py`; "
                              "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise KlaraError(rc, lib().klb_last_error().decode())
    return rc
