"""ctypes binding of the CPU oracle (oracle/klb_oracle.c).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libklb_oracle.so")

MH, MALA, HMC, NUTS = 0, 1, 2, 3
ISO, SHIFTED, DENSE, ROSEN, LOGIT = 0, 1, 2, 3, 4
VANILLA, ACCRATE, DUALAVG = 0, 1, 2


class OrcConfig(C.Structure):
    _fields_ = [
        ("sampler", C.c_int32), ("target", C.c_int32), ("tuner", C.c_int32), ("arith", C.c_int32),
        ("nchains", C.c_int64), ("dim", C.c_int64), ("nsteps", C.c_int64), ("burnin", C.c_int64),
        ("thinning", C.c_int64),
        ("step", C.c_double), ("nleaps", C.c_int32),
        ("target_rate", C.c_double), ("score_k", C.c_double), ("period", C.c_int64),
        ("verbose", C.c_int32), ("monitor", C.c_uint32), ("diagnostics", C.c_uint32),
        ("seed", C.c_uint64), ("chain_offset", C.c_uint64), ("t0", C.c_uint64),
        ("score", C.c_int32), ("nv", C.c_int32), ("nthreads", C.c_int32),
        ("da_nadapt", C.c_int64), ("da_t0", C.c_int64),
        ("da_eps0bar", C.c_double), ("da_h0bar", C.c_double), ("da_gamma", C.c_double), ("da_kappa", C.c_double),
        ("da", C.c_void_p),
        ("nuts_maxdelta", C.c_int32), ("nuts_maxndoublings", C.c_int32), ("nuts_ndoublings", C.c_void_p),
        ("nuts_a", C.c_void_p), ("nuts_na", C.c_void_p),
    ]


class OrcTune(C.Structure):
    _fields_ = [("step", C.c_double), ("accepted", C.c_int64), ("proposed", C.c_int64),
                ("totproposed", C.c_int64), ("rate", C.c_double)]


DA_DTYPE = np.dtype([(n, "<f8") for n in ("lambda", "mu", "epsbar", "hbar", "hweight", "epsweight", "nleaps", "count")])
TUNE_DTYPE = np.dtype([("step", "<f8"), ("accepted", "<i8"), ("proposed", "<i8"),
                       ("totproposed", "<i8"), ("rate", "<f8")])


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc, no GPU needed)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, p)) > os.path.getmtime(_LIB_PATH)
            for p in ("klb_oracle.c", "../klara.jl_b200/csrc/klb_math.h", "../klara.jl_b200/csrc/klb_tables.h")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dbl = C.c_double
        L.orc_logistic.restype = dbl
        L.orc_logistic.argtypes = [dbl] * 5
        L.orc_logistic_rate_score.restype = dbl
        L.orc_logistic_rate_score.argtypes = [dbl, dbl]
        L.orc_erf_rate_score.restype = dbl
        L.orc_erf_rate_score.argtypes = [dbl, dbl]
        L.orc_exp.restype = dbl
        L.orc_exp.argtypes = [dbl]
        L.orc_log.restype = dbl
        L.orc_log.argtypes = [dbl]
        L.orc_uniform.restype = dbl
        L.orc_uniform.argtypes = [C.c_uint64] * 3
        L.orc_normals.restype = None
        L.orc_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p]
        L.orc_erf.restype = dbl
        L.orc_erf.argtypes = [dbl]
        L.orc_philox.restype = None
        L.orc_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_philox_rounds.restype = C.c_int
        L.orc_plan_nv.restype = C.c_int
        L.orc_plan_nv.argtypes = [C.c_int64]
        L.orc_dot.restype = dbl
        L.orc_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int]
        L.orc_npoststeps.restype = C.c_int64
        L.orc_npoststeps.argtypes = [C.c_int64] * 3
        L.orc_tuner_state.restype = None
        L.orc_tuner_state.argtypes = [C.POINTER(OrcConfig), C.c_void_p]
        L.orc_da_state.restype = None
        L.orc_da_state.argtypes = [C.POINTER(OrcConfig), C.c_void_p, C.c_void_p, C.c_int]
        L.orc_run.restype = C.c_int
        L.orc_run.argtypes = [C.POINTER(OrcConfig)] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 4
        L.orc_eval_target.restype = C.c_int
        L.orc_eval_target.argtypes = [C.POINTER(OrcConfig)] + [C.c_void_p] * 4
        L.orc_max_threads.restype = C.c_int
        L.orc_ma.restype = dbl
        L.orc_ma.argtypes = [dbl, dbl, dbl, C.c_int]
        L.orc_ess.restype = None
        L.orc_ess.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        L.orc_stats.restype = None
        L.orc_stats.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        L.orc_acceptance.restype = None
        L.orc_acceptance.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
        L.orc_ess_series.restype = dbl
        L.orc_ess_series.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox(ctr, key, rounds=10):
    """Philox4x32-`rounds`: 10 = the published generator (known-answer vectors), 7 = the contract's"""
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(_ptr(c), _ptr(k), rounds, _ptr(out))
    return out


def philox_rounds():
    return lib().orc_philox_rounds()


def normals(seed, chain, t, n):
    out = np.empty(n, dtype=np.float64)
    lib().orc_normals(seed, chain, t, n, _ptr(out))
    return out


def uniform(seed, chain, t):
    return lib().orc_uniform(seed, chain, t)


def exp(x):
    return lib().orc_exp(float(x))


def log(x):
    return lib().orc_log(float(x))


def erf(x):
    return lib().orc_erf(float(x))


def plan_nv(dim):
    return lib().orc_plan_nv(dim)


def dot(a, b, nv=None, arith=0):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    nv = plan_nv(a.size) if nv is None else nv
    return lib().orc_dot(_ptr(a), _ptr(b), a.size, nv, arith)


def logit_params(X, y, lam):
    """tparams of the LOGIT target: [lambda, ndata, X (ndata x d, row-major), y]"""
    X = np.ascontiguousarray(X, dtype=np.float64)
    return np.concatenate([[float(lam), float(X.shape[0])], X.reshape(-1), np.asarray(y, dtype=np.float64).reshape(-1)])


def npoststeps(burnin, thinning, nsteps):
    return lib().orc_npoststeps(burnin, thinning, nsteps)


def make_config(sampler, target, nchains, dim, nsteps, burnin=0, thinning=1, step=0.1, nleaps=10,
                tuner=VANILLA, target_rate=0.574, score_k=7.0, period=100, verbose=0,
                monitor=1, diagnostics=0, seed=0, chain_offset=0, t0=0, arith=0, nv=None, nthreads=1,
                nadapt=1000, eps0bar=1.0, h0bar=0.0, gamma=0.05, da_t0=10, kappa=0.75, score=0, maxdelta=1000,
                maxndoublings=5):
    cfg = OrcConfig()
    cfg.nuts_maxdelta, cfg.nuts_maxndoublings, cfg.nuts_ndoublings = maxdelta, maxndoublings, None
    cfg.score = score                      # AcceptanceRateMCTuner: 0 logistic_rate_score, 1 erf_rate_score
    cfg.sampler, cfg.target, cfg.tuner, cfg.arith = sampler, target, tuner, arith
    cfg.nchains, cfg.dim, cfg.nsteps, cfg.burnin, cfg.thinning = nchains, dim, nsteps, burnin, thinning
    cfg.step, cfg.nleaps = step, nleaps
    cfg.target_rate, cfg.score_k, cfg.period, cfg.verbose = target_rate, score_k, period, verbose
    cfg.monitor, cfg.diagnostics = monitor, diagnostics
    cfg.seed, cfg.chain_offset, cfg.t0 = seed, chain_offset, t0
    cfg.nv = (0 if target == LOGIT else plan_nv(dim)) if nv is None else nv   # LOGIT: sequential order
    cfg.nthreads = nthreads
    cfg.da_nadapt, cfg.da_t0 = nadapt, da_t0
    cfg.da_eps0bar, cfg.da_h0bar, cfg.da_gamma, cfg.da_kappa = eps0bar, h0bar, gamma, kappa
    cfg.da = None
    return cfg


def tuner_state(cfg):
    t = np.zeros(cfg.nchains, dtype=TUNE_DTYPE)
    one = OrcTune()
    lib().orc_tuner_state(C.byref(cfg), C.byref(one))
    t["step"], t["accepted"], t["proposed"] = one.step, one.accepted, one.proposed
    t["totproposed"], t["rate"] = one.totproposed, one.rate
    return t


def da_state(cfg, first=True):
    """(tune, da) records of tuner_state + sampler_state for HMC with a DualAveragingMCTuner; first=False gives
    what reset!(tune, sampler, tuner) leaves (step = 1)"""
    t = np.zeros(cfg.nchains, dtype=TUNE_DTYPE)
    d = np.zeros(cfg.nchains, dtype=DA_DTYPE)
    one, oned = OrcTune(), (C.c_double * 8)()
    lib().orc_da_state(C.byref(cfg), C.byref(one), oned, int(first))
    t["step"], t["accepted"], t["proposed"] = one.step, one.accepted, one.proposed
    t["totproposed"], t["rate"] = one.totproposed, one.rate
    for i, n in enumerate(DA_DTYPE.names):
        d[n] = oned[i]
    return t, d


def eval_target(cfg, x, tparams=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    tp = None if tparams is None else np.ascontiguousarray(tparams, dtype=np.float64)
    lt = C.c_double()
    g = np.empty_like(x)
    rc = lib().orc_eval_target(C.byref(cfg), _ptr(tp), _ptr(x), C.byref(lt), _ptr(g))
    if rc:
        raise ValueError("orc_eval_target failed: %d" % rc)
    return lt.value, g


def run(cfg, x0, tparams=None, sigma=None, tune=None, logtarget=None, da=None):
    """Run all chains.  x0: (nchains, dim) array (row c = chain c, i.e. Julia's d x nchains
    column-major matrix).  Returns dict with final state, tune records and the monitored output
    in the NState layout: value (nchains, npost, dim), logtarget / accept (nchains, npost)."""
    N, d = cfg.nchains, cfg.dim
    x = np.array(x0, dtype=np.float64, order="C").reshape(N, d).copy()
    npost = npoststeps(cfg.burnin, cfg.thinning, cfg.nsteps)
    tp = None if tparams is None else np.ascontiguousarray(tparams, dtype=np.float64)
    sg = None if sigma is None else np.ascontiguousarray(sigma, dtype=np.float64)
    if cfg.tuner == DUALAVG:
        t0_, d0_ = da_state(cfg)
        tune = t0_ if tune is None else tune.copy()
        da = d0_ if da is None else da.copy()
        cfg.da = da.ctypes.data
    else:
        tune = tuner_state(cfg) if tune is None else tune.copy()
    initialized = logtarget is not None
    lt = np.zeros(N) if logtarget is None else np.array(logtarget, dtype=np.float64)
    ov = np.zeros((N, npost, d)) if cfg.monitor & 1 else None
    ol = np.zeros((N, npost)) if cfg.monitor & 2 else None
    og = np.zeros((N, npost, d)) if cfg.monitor & 4 else None
    oa = np.zeros((N, npost), dtype=np.uint8) if cfg.diagnostics & 1 else None
    ond = np.zeros((N, npost), dtype=np.uint8) if (cfg.diagnostics & 2 and cfg.sampler == NUTS) else None
    cfg.nuts_ndoublings = None if ond is None else ond.ctypes.data
    nuts_da = cfg.sampler == NUTS and cfg.tuner == DUALAVG          # :a, :na exist for that pair only (NUTS.jl:317,344)
    ona = np.zeros((N, npost)) if (cfg.diagnostics & 4 and nuts_da) else None
    onn = np.zeros((N, npost), dtype=np.int32) if (cfg.diagnostics & 8 and nuts_da) else None
    cfg.nuts_a = None if ona is None else ona.ctypes.data
    cfg.nuts_na = None if onn is None else onn.ctypes.data
    rc = lib().orc_run(C.byref(cfg), _ptr(tp), _ptr(sg), _ptr(x), _ptr(lt), _ptr(tune), int(initialized),
                       _ptr(ov), _ptr(ol), _ptr(og), _ptr(oa))
    if rc:
        raise ValueError("oracle: initial log-target/gradient not finite in chain %d" % (-rc - 1))
    cfg.da = None
    cfg.nuts_ndoublings = None
    cfg.nuts_a = cfg.nuts_na = None
    return {"a": ona, "na": onn, "x": x, "logtarget_state": lt, "tune": tune, "da": da, "value": ov, "logtarget": ol,
            "gradlogtarget": og, "accept": oa, "ndoublings": ond, "npost": npost}


def uniform_seq(seed, chain, t, q):
    """uniform number q of (seed, chain, transition t): what NUTS consumes in order (q = 0: the accept uniform)"""
    f = lib().orc_uniform_seq
    f.restype = C.c_double
    f.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    return f(seed, chain, t, q)


def max_threads():
    return lib().orc_max_threads()


def ess(value, nthreads=None):
    """ess(chain, :imse) per (chain, coordinate) of a (nchains, npost, dim) value array"""
    v = np.ascontiguousarray(value, dtype=np.float64)
    if v.ndim == 2:
        v = v[None]
    N, P, d = v.shape
    out = np.empty((N, d))
    lib().orc_ess(_ptr(v), N, P, d, _ptr(out), nthreads or max_threads())
    return out


STAT_NAMES = ("mean", "mcvar_iid", "mcvar_imse", "ess", "iact")


def stats(value, nthreads=None):
    """mean / mcvar(:iid) / mcvar(:imse) / ess / iact per (chain, coordinate) of a (nchains, npost, dim) value
    array: dict of (nchains, dim) arrays"""
    v = np.ascontiguousarray(value, dtype=np.float64)
    if v.ndim == 2:
        v = v[None]
    N, P, d = v.shape
    out = np.empty((5, N, d))
    lib().orc_stats(_ptr(v), N, P, d, _ptr(out), nthreads or max_threads())
    return dict(zip(STAT_NAMES, out))


def acceptance(accept=None, value=None):
    """acceptance(s) from the :accept diagnostic (nchains, npost), or acceptance(s, diagnostics=false) from the
    values (nchains, npost, dim)"""
    if accept is not None:
        a = np.ascontiguousarray(accept, dtype=np.uint8)
        N, P = a.shape
        out = np.empty(N)
        lib().orc_acceptance(_ptr(a), None, N, P, 0, _ptr(out))
        return out
    v = np.ascontiguousarray(value, dtype=np.float64)
    N, P, d = v.shape
    out = np.empty(N)
    lib().orc_acceptance(None, _ptr(v), N, P, d, _ptr(out))
    return out
