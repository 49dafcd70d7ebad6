"""Python mirror of the Klara.jl user surface for the MCMC hot path, on top of the C ABI.

Same names, argument meaning and error behaviour as the reference (paths relative to the Klara.jl
checkout, commit ffa4f6d0):

    BasicContMuvParameter   src/variables/parameters/BasicContMuvParameter.jl:383-411
    likelihood_model        src/models/generators.jl:5-18
    MH / MALA / HMC         src/samplers/MH.jl:47-66, MALA.jl:61-70, HMC.jl:89-100
    BasicMCRange            src/ranges/BasicMCRange.jl:7-33
    VanillaMCTuner          src/tuners/VanillaMCTuner.jl:6-16
    AcceptanceRateMCTuner   src/tuners/AcceptanceRateMCTuner.jl:25-46
    BasicMCJob / run / reset / output    src/jobs/BasicMCJob.jl:24-244, 279

Batch extension: the initial value may be a `(nchains, dim)` array (Julia: `dim x nchains` matrix);
every row is an independent chain with its own tuner record, i.e. `map(run, jobs)`
(src/jobs/jobs.jl:212) executed in lockstep on the GPU.

Julia symbols (`:p`, `:value`) are plain strings here; `Dict(:p => v)` is `{"p": v}`.
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L
from .targets import Target

__all__ = ["SyntheticNormal", "device_peak", "BasicContMuvParameter", "Hyperparameter", "Data", "GenericModel", "likelihood_model", "MH", "MALA", "HMC", "NUTS", "BasicMCRange",
           "VanillaMCTuner", "AcceptanceRateMCTuner", "DualAveragingMCTuner", "BasicMCTune", "DualAveragingMCTune", "BasicMCJob", "run", "reset", "output",
           "BasicContMuvParameterNState", "logistic", "logistic_rate_score", "erf_rate_score", "ess", "mean", "mcvar", "mcse", "iact",
           "acceptance", "diagnostics"]


# ----------------------------------------------------------------------------- scalars
def logistic(x, l=1., k=1., x0=0., y0=0.):
    """l/(1+exp(-k*(x-x0)))+y0        src/stats/logistic.jl:11"""
    return l / (1 + math.exp(-k * (x - x0))) + y0


def logistic_rate_score(x, k=7.):
    """src/tuners/AcceptanceRateMCTuner.jl:9"""
    return logistic(x, 2., k, 0., 0.)


def erf_rate_score(x, k=3.):
    """erf(k*x)+1        src/tuners/AcceptanceRateMCTuner.jl:17"""
    return math.erf(k * x) + 1


# ----------------------------------------------------------------------------- parameter / model
class BasicContMuvParameter:
    """BasicContMuvParameter(key; logtarget, gradlogtarget) with a target descriptor.

    `gradlogtarget` may be omitted or must be the same descriptor: the device targets carry their
    analytic gradient (the reference's autodiff path, src/autodiff/, is out of scope)."""

    def __init__(self, key, logtarget=None, gradlogtarget=None, loglikelihood=None, logprior=None, nkeys=0, index=0):
        # loglikelihood= + logprior= (logtarget = their sum, BasicContMuvParameter.jl:185-190): the bound methods of
        # one descriptor, as in doc/examples/swiss/*/analytical.jl
        if logtarget is None and loglikelihood is not None:
            owner = getattr(loglikelihood, "__self__", None)
            if not isinstance(owner, Target) or getattr(logprior, "__self__", None) is not owner:
                raise TypeError("loglikelihood and logprior must be the .loglikelihood / .logprior of one target descriptor")
            logtarget = owner
        self.nkeys = nkeys
        if not isinstance(logtarget, Target):
            raise TypeError("logtarget must be a klara_b200 target descriptor (IsoGaussian(), ...); arbitrary "
                            "host closures cannot run inside the CUDA kernels")
        if gradlogtarget is not None and gradlogtarget is not logtarget and gradlogtarget != logtarget.gradient:  # noqa: E501
            raise TypeError("gradlogtarget must be omitted, the same descriptor, or descriptor.gradient")
        self.key = key
        self.index = index
        self.logtarget = logtarget
        self.gradlogtarget = logtarget.gradient
        self.target = logtarget


class Hyperparameter:
    """Hyperparameter(key) = Constant: a vertex whose value is fixed by v0 (src/variables/variables.jl:41-60).
    Its value reaches the parameter's closures through the states vector (BasicContMuvParameter.jl:497-501)."""

    def __init__(self, key, index=0):
        self.key, self.index = key, index


class Data(Hyperparameter):
    """Data(key)        src/variables/variables.jl:64-95"""


class GenericModel:
    """GenericModel(vs, ds=[]; isdirected=true, isindexed=true)        src/models/GenericModel.jl:94-119

    isindexed=true: the vertices already carry their positions (`index`, 1-based) and are stored sorted by it;
    isindexed=false: they are stored in the given order and numbered 1..n.  The job only uses the graph to locate
    the parameter and to collect the states of the other vertices in vertex order (src/jobs/BasicMCJob.jl:50-51);
    dependences are kept as given (graph algorithms are outside the hot path)."""

    def __init__(self, vertices, dependences=None, isdirected=True, isindexed=True):
        vs = list(vertices)
        if isindexed:
            idx = [v.index for v in vs]
            if sorted(idx) != list(range(1, len(vs) + 1)):
                raise AssertionError("isindexed=true needs vertices indexed 1..%d, got %s (pass isindexed=false to "
                                     "number them in the given order)" % (len(vs), idx))
            vs.sort(key=lambda v: v.index)
        else:
            for i, v in enumerate(vs):
                v.index = i + 1
        self.isdirected = bool(isdirected)
        self.vertices = vs
        self.edges = list(dependences) if dependences is not None else []
        self.ofkey = {v.key: i for i, v in enumerate(self.vertices)}


def likelihood_model(vertices, isindexed=True, isdirected=True):
    """likelihood_model(vs; isdirected, isindexed) / likelihood_model(v, isindexed)        src/models/generators.jl:5-18:
    every non-parameter vertex points at every parameter"""
    if not isinstance(vertices, (list, tuple)):
        return GenericModel([vertices], isindexed=isindexed)
    m = GenericModel(vertices, isdirected=isdirected, isindexed=isindexed)
    params = [v for v in m.vertices if isinstance(v, BasicContMuvParameter)]
    m.edges = [(u.key, p.key) for p in params for u in m.vertices if not isinstance(u, BasicContMuvParameter)]
    return m


class SyntheticNormal:
    """Initial value generated on the device: x0[c, i] = N(0,1) of the job's Philox stream (seed, global chain
    chain_offset + c, transition 0, element i) -- the input of the benchmark configurations.
    `{"p": SyntheticNormal(nchains, dim)}` in place of a `(nchains, dim)` array."""

    def __init__(self, nchains, dim):
        self.shape = (int(nchains), int(dim))


def device_peak(kind="fp64", device=0):
    """measured fp64 results/s ("fp64") or fp64 tensor-pipe flop/s ("dmma") of one device (klb_device_peak)"""
    v = C.c_double()
    L.check(L.lib().klb_device_peak(device, {"fp64": L.PEAK_FP64, "dmma": L.PEAK_DMMA}[kind], C.byref(v)))
    return v.value


# ----------------------------------------------------------------------------- samplers
class MH:
    """MH(sigma::Vector): Metropolis with the symmetric normal random walk MvNormal(x, sigma),
    sigma = per-coordinate standard deviations (src/samplers/MH.jl:64)."""
    code = L.SAMPLER_MH

    def __init__(self, sigma):
        self.sigma = np.ascontiguousarray(np.atleast_1d(sigma), dtype=np.float64)
        if self.sigma.ndim != 1:
            raise TypeError("only the diagonal MH(sigma::Vector) proposal is supported on the device")
        assert np.all(self.sigma > 0), "Proposal standard deviations should be positive"
        self.symmetric = True
        self.normalised = True


class MALA:
    """MALA(driftstep=1.)        src/samplers/MALA.jl:61-70"""
    code = L.SAMPLER_MALA

    def __init__(self, driftstep=1.):
        assert driftstep > 0, "Drift step is not positive"
        self.driftstep = float(driftstep)


class HMC:
    """HMC(leapstep=0.1, nleaps=10)        src/samplers/HMC.jl:89-100"""
    code = L.SAMPLER_HMC

    def __init__(self, leapstep=0.1, nleaps=10):
        assert leapstep > 0, "Leapfrog step is not positive"
        assert nleaps > 0, "Number of leapfrog steps is not positive"
        self.leapstep = float(leapstep)
        self.nleaps = int(nleaps)


class NUTS:
    """NUTS(leapstep=0.1; maxδ=1000, maxndoublings=5)        src/samplers/NUTS.jl:228-241
    The multivariate transition as the reference computes it (its four tree states are one object, NUTS.jl:198-225;
    DESIGN.md section 6b), with VanillaMCTuner or DualAveragingMCTuner, on the elementwise targets and the logistic-regression
    target; diagnostics :accept and :ndoublings, and :a / :na with dual averaging (NUTS.jl:317)."""
    code = L.SAMPLER_NUTS

    def __init__(self, leapstep=0.1, maxdelta=1000, maxndoublings=5):
        assert leapstep > 0, "Leapfrog step is not positive"
        assert maxdelta > 0, "maxδ is not positive"
        assert maxndoublings > 0, "Maximum number of doublings is not positive"
        self.leapstep = float(leapstep)
        self.maxdelta = int(maxdelta)
        self.maxndoublings = int(maxndoublings)


# ----------------------------------------------------------------------------- range / tuners
class BasicMCRange:
    """BasicMCRange(; burnin=0, thinning=1, nsteps=100): postrange = (burnin+1):thinning:nsteps
    (src/ranges/BasicMCRange.jl:7-33)."""

    def __init__(self, burnin=0, thinning=1, nsteps=100):
        assert burnin >= 0, "Number of burn-in iterations should be non-negative"
        assert thinning >= 1, "Thinning should be >= 1"
        assert nsteps > burnin, "Total number of MCMC iterations should be greater than number of burn-in iterations"
        self.burnin, self.thinning, self.nsteps = int(burnin), int(thinning), int(nsteps)
        self.postrange = range(self.burnin + 1, self.nsteps + 1, self.thinning)
        self.npoststeps = len(self.postrange)


class VanillaMCTuner:
    """VanillaMCTuner(; period=100, verbose=false)        src/tuners/VanillaMCTuner.jl:6-16
    `verbose` switches the acceptance counters on (iterate/HMC.jl:129-133); the kernels record the acceptance rate
    of every burn-in period (BasicMCJob.burnin_rates) and run() prints the reference's per-period lines when the
    launch returns."""
    code = L.TUNER_VANILLA

    def __init__(self, period=100, verbose=False):
        assert period > 0, "Adaptation period should be positive"
        self.period, self.verbose = int(period), bool(verbose)


class AcceptanceRateMCTuner:
    """AcceptanceRateMCTuner(targetrate; score=logistic_rate_score, period=100, verbose=false)
    (src/tuners/AcceptanceRateMCTuner.jl:25-44).  `score` is one of the reference's two score functions,
    logistic_rate_score (2/(1+exp(-k*x)), k = 7) or erf_rate_score (erf(k*x)+1, k = 3), optionally with another
    steepness `k`; tune! multiplies the step by score(rate - targetrate) (:46).  Arbitrary host closures cannot run
    inside the kernels."""
    code = L.TUNER_ACCEPTANCE_RATE

    def __init__(self, targetrate, score=logistic_rate_score, period=100, verbose=False, k=None):
        assert 0 < targetrate < 1, "Target acceptance rate should be between 0 and 1"
        assert period > 0, "Tuning period should be positive"
        if score is not logistic_rate_score and score is not erf_rate_score:
            raise TypeError("score must be logistic_rate_score or erf_rate_score (the two the device evaluates)")
        self.score_code = L.SCORE_ERF if score is erf_rate_score else L.SCORE_LOGISTIC
        k = (3. if score is erf_rate_score else 7.) if k is None else k
        self.targetrate, self.score, self.k = float(targetrate), score, float(k)
        self.period, self.verbose = int(period), bool(verbose)


class DualAveragingMCTuner:
    """DualAveragingMCTuner(targetrate, nadapt; ε0bar=1., h0bar=0., γ=0.05, t0=10, κ=0.75, period=100, verbose=false)
    (src/tuners/DualAveragingMCTuner.jl:53-93): Nesterov dual averaging of the HMC leapfrog step (Hoffman & Gelman),
    with nleaps = max(1, round(λ/step)) per chain, λ = nleaps*leapstep of the sampler.  Greek keywords are spelled
    out: eps0bar, h0bar, gamma, t0, kappa."""
    code = L.TUNER_DUAL_AVERAGING

    def __init__(self, targetrate, nadapt, eps0bar=1., h0bar=0., gamma=0.05, t0=10, kappa=0.75, period=100,
                 verbose=False):
        assert 0 < targetrate < 1, "Target acceptance rate should be between 0 and 1"
        assert nadapt > 0, "Number of adaptation steps should be positive"
        assert eps0bar > 0, "ε0bar should be positive"
        assert period > 0, "Period over which acceptance rate is reported in verbose mode should be positive"
        assert t0 > 0, "t0 should be positive"
        self.targetrate, self.nadapt = float(targetrate), int(nadapt)
        self.eps0bar, self.h0bar, self.gamma, self.t0, self.kappa = float(eps0bar), float(h0bar), float(gamma), int(t0), float(kappa)
        self.period, self.verbose = int(period), bool(verbose)


class DualAveragingMCTune:
    """Per-chain DualAveragingMCTune records (src/tuners/DualAveragingMCTuner.jl:1-13) as arrays over chains;
    `count` is sstate.count."""

    def __init__(self, base, da):
        self.step, self.accepted, self.proposed, self.totproposed, self.rate = \
            base.step, base.accepted, base.proposed, base.totproposed, base.rate
        (self.lam, self.mu, self.epsbar, self.hbar, self.hweight, self.epsweight) = (da[:, i].copy() for i in range(6))
        self.nleaps, self.count = da[:, 6].astype(np.int64), da[:, 7].astype(np.int64)


class BasicMCTune:
    """Per-chain tuner records (src/tuners/tuners.jl:5-25) as arrays over chains."""

    def __init__(self, step, accepted, proposed, totproposed, rate):
        self.step, self.accepted, self.proposed, self.totproposed, self.rate = step, accepted, proposed, totproposed, rate


# ----------------------------------------------------------------------------- output container
class BasicContMuvParameterNState:
    """Monitored output (src/nstates/ParameterNStates/BasicContMuvParameterNState.jl:1-61).

    numpy C-order arrays; memory layout identical to the Julia column-major NState fields:
      value          (nchains, npost, dim)   <->  Julia  dim x npost x nchains   (value[:, i] = sample i)
      logtarget      (nchains, npost)
      gradlogtarget  (nchains, npost, dim)
      diagnosticvalues (nchains, npost) uint8 accept flags
    For a job built from a single vector the leading chain axis is dropped."""

    def __init__(self, size, n):
        self.size, self.n = size, n
        self.value = None
        self.logtarget = None
        self.gradlogtarget = None
        self.diagnosticvalues = None
        self.diagnostickeys = []
        self._job = None


_MONITOR_BITS = {"value": L.MONITOR_VALUE, "logtarget": L.MONITOR_LOGTARGET, "gradlogtarget": L.MONITOR_GRADLOGTARGET}
_DIAG_BITS = {"accept": L.DIAG_ACCEPT, "ndoublings": L.DIAG_NDOUBLINGS, "a": L.DIAG_NUTS_A, "na": L.DIAG_NUTS_NA}
_DIAG_FIELDS = {"accept": (L.OUT_ACCEPT, np.uint8), "ndoublings": (L.OUT_NDOUBLINGS, np.uint8), "a": (L.OUT_NUTS_A, np.float64),
                "na": (L.OUT_NUTS_NA, np.int32)}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------- the job
class BasicMCJob:
    """BasicMCJob(model, sampler, mcrange, v0; tuner=VanillaMCTuner(), outopts=...)
    (src/jobs/BasicMCJob.jl:107-185).

    Extra keyword arguments of the batched GPU job:
      seed          Philox key (the reference draws from Julia's unseeded global RNG)
      arith         "reference" (un-fused, reference evaluation order) or "fma"
      device        CUDA device ordinal
      chain_offset  global index of the first chain when the chains of one logical job are sharded
                    over several processes / GPUs
      ngpus         shard the chains over this many devices of THIS process (klb_multi_*: contiguous chain blocks,
                    global-chain RNG, one closing all-gather of the final states and tuner records by the copy
                    engines); 0 = every visible device.  `devices` lists the ordinals (default 0 .. ngpus-1).
                    Results do not depend on ngpus.
    """

    def __init__(self, model, sampler, mcrange, v0, tuner=None, outopts=None, pindex=None,
                 seed=0, arith="reference", device=0, chain_offset=0, verbose=False, ngpus=1, devices=None,
                 resetpstate=True, check=False):
        # resetpstate, check: accepted for the reference's signature (src/jobs/BasicMCJob.jl:107-119).  resetpstate only
        # matters to BasicGibbsJob's reset (a draw from the parameter's prior, host-side: not on this path); the
        # constructor's consistency checks always run here.
        self.resetpstate, self.check = bool(resetpstate), bool(check)
        tuner = VanillaMCTuner() if tuner is None else tuner
        self.model, self.sampler, self.range, self.tuner = model, sampler, mcrange, tuner
        self.pindex = next(i for i, v in enumerate(model.vertices) if isinstance(v, BasicContMuvParameter)) \
            if pindex is None else pindex
        self.parameter = model.vertices[self.pindex]
        # outopts defaults: augment_parameter_outopts!  (src/jobs/jobs.jl:9-43)
        oo = dict(outopts) if outopts else {}
        oo.setdefault("destination", "nstate")
        if oo["destination"] != "none":
            oo.setdefault("monitor", ["value"])
            oo.setdefault("diagnostics", [])
        else:
            oo.setdefault("monitor", [])
            oo.setdefault("diagnostics", [])
        if oo["destination"] not in ("nstate", "iostream", "none"):
            raise ValueError(":destination must be set to :nstate or :iostream or :none, got %r" % (oo["destination"],))
        if oo["destination"] == "iostream":         # augment_variable_outopts!  (src/jobs/jobs.jl:17-29)
            oo.setdefault("filepath", "")
            oo.setdefault("filesuffix", "csv")
            oo.setdefault("flush", False)
        for m in oo["monitor"]:
            if m not in _MONITOR_BITS:
                raise KeyError("cannot monitor %r on the device path" % (m,))
        for dg in oo["diagnostics"]:
            # :accept everywhere; :ndoublings for NUTS (src/samplers/NUTS.jl:285); :a and :na for NUTS with
            # DualAveragingMCTuner (NUTS.jl:317,344)
            nuts_da = isinstance(sampler, NUTS) and isinstance(tuner, DualAveragingMCTuner)
            if dg != "accept" and not (dg == "ndoublings" and isinstance(sampler, NUTS)) and not (dg in ("a", "na") and nuts_da):
                raise KeyError("unknown diagnostic %r" % (dg,))
        self.outopts = oo

        # hyper-parameters / data: the states of the model's other vertices, in vertex order, are what the
        # reference passes to the closures as `v` (BasicContMuvParameter.jl:497-501; `nkeys` only tested > 0)
        others = [v for i, v in enumerate(model.vertices) if i != self.pindex]
        if others:
            if isinstance(v0, (list, tuple)) and len(v0) == len(model.vertices):
                v0 = {v.key: val for v, val in zip(model.vertices, v0)}        # v0::Vector: values in vertex order (BasicMCJob.jl:139-152)
            if not isinstance(v0, dict):
                raise TypeError("a model with Hyperparameter / Data vertices needs v0 as a Dict of initial values (or a "
                                "list with one entry per vertex)")
            if not hasattr(self.parameter.target, "bind"):
                raise TypeError("target %s takes no hyper-parameters" % type(self.parameter.target).__name__)
            self.parameter.target.bind([v0[v.key] for v in others])
        x0 = v0[self.parameter.key] if isinstance(v0, dict) else v0
        synthetic = isinstance(x0, SyntheticNormal)
        if synthetic:
            self.single = False
            self.nchains, self.dim = x0.shape
        else:
            x0 = np.asarray(x0, dtype=np.float64)
            self.single = x0.ndim == 1
            x0 = np.ascontiguousarray(np.atleast_2d(x0))
            self.nchains, self.dim = x0.shape

        cfg = L.KlbConfig()
        cfg.struct_size = C.sizeof(L.KlbConfig)
        cfg.sampler, cfg.target, cfg.tuner = sampler.code, self.parameter.target.code, tuner.code
        cfg.arith = {"reference": L.ARITH_REFERENCE, "fma": L.ARITH_FMA}[arith]
        cfg.nchains, cfg.dim = self.nchains, self.dim
        cfg.nsteps, cfg.burnin, cfg.thinning = mcrange.nsteps, mcrange.burnin, mcrange.thinning
        cfg.step = getattr(sampler, "leapstep", getattr(sampler, "driftstep", 1.0))
        cfg.nleaps = getattr(sampler, "nleaps", 1)
        cfg.target_rate = getattr(tuner, "targetrate", 0.5)
        cfg.score_k = getattr(tuner, "k", 7.0)
        cfg.score = getattr(tuner, "score_code", L.SCORE_LOGISTIC)
        cfg.period, cfg.verbose = tuner.period, int(tuner.verbose)
        cfg.monitor = sum(_MONITOR_BITS[m] for m in set(oo["monitor"]))
        cfg.diagnostics = sum(_DIAG_BITS[k] for k in set(oo["diagnostics"]))
        cfg.nuts_maxdelta, cfg.nuts_maxndoublings = getattr(sampler, "maxdelta", 0), getattr(sampler, "maxndoublings", 0)
        cfg.destination = L.DEST_NONE if oo["destination"] == "none" else L.DEST_NSTATE
        cfg.seed, cfg.chain_offset, cfg.device = seed, chain_offset, device
        cfg.da_nadapt, cfg.da_t0 = getattr(tuner, "nadapt", 0), getattr(tuner, "t0", 10)
        cfg.da_eps0bar, cfg.da_h0bar = getattr(tuner, "eps0bar", 1.0), getattr(tuner, "h0bar", 0.0)
        cfg.da_gamma, cfg.da_kappa = getattr(tuner, "gamma", 0.05), getattr(tuner, "kappa", 0.75)
        self.cfg = cfg
        self._h = C.c_void_p()
        self._m = None                          # klb_multi handle when the chains are sharded over devices
        lib = L.lib()
        if ngpus != 1 or devices is not None:
            devs = None
            if devices is not None:
                ngpus = len(devices)
                devs = (C.c_int32 * ngpus)(*devices)
            self._m = C.c_void_p()
            L.check(lib.klb_multi_create(C.byref(cfg), ngpus, devs, C.byref(self._m)))
            self.ngpus = lib.klb_multi_ngpus(self._m)
            self._shards = []                   # (job handle, first chain within the logical job, chains)
            for g in range(self.ngpus):
                h, c = C.c_void_p(), L.KlbConfig()
                L.check(lib.klb_multi_job(self._m, g, C.byref(h)))
                L.check(lib.klb_job_config(h, C.byref(c)))
                self._shards.append((h, c.chain_offset - chain_offset, c.nchains))
            self._h = self._shards[0][0]
        else:
            self.ngpus = 1
            L.check(lib.klb_job_create(C.byref(cfg), C.byref(self._h)))
            self._shards = [(self._h, 0, self.nchains)]
        set_target = lib.klb_multi_set_target_f64 if self._m else lib.klb_job_set_target_f64
        top = self._m if self._m else self._h
        try:
            for which, arr in self.parameter.target.params(self.dim):
                L.check(set_target(top, which, _ptr(arr), arr.size))
            if isinstance(sampler, MH):
                if sampler.sigma.size != self.dim:
                    raise AssertionError("MH sigma has %d entries, parameter has %d" % (sampler.sigma.size, self.dim))
                L.check(set_target(top, L.PARAM_SIGMA, _ptr(sampler.sigma), self.dim))
            # initialize!: first target (+gradient) evaluation; finiteness asserts (HMC.jl:113-114)
            if synthetic:
                L.check((lib.klb_multi_set_state_synthetic if self._m else lib.klb_job_set_state_synthetic)(top))
            else:
                L.check((lib.klb_multi_set_state if self._m else lib.klb_job_set_state)(top, _ptr(x0)))
        except Exception:
            self.close()
            raise
        self.count = 0

    # -- lifecycle
    def close(self):
        if getattr(self, "_m", None) is not None and self._m.value:
            L.lib().klb_multi_destroy(self._m)
            self._m, self._h, self._shards = None, C.c_void_p(), []
        elif getattr(self, "_h", None) is not None and self._h.value:
            L.lib().klb_job_destroy(self._h)
            self._h = C.c_void_p()

    def _call(self, name, *args):
        """klb_multi_<name> when the job is sharded over devices, klb_job_<name> otherwise"""
        lib = L.lib()
        if self._m:
            return L.check(getattr(lib, "klb_multi_" + name)(self._m, *args))
        return L.check(getattr(lib, "klb_job_" + name)(self._h, *args))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference API
    def run(self):
        """run(job)        src/jobs/BasicMCJob.jl:212-244"""
        self._call("run")
        self.count = self.range.npoststeps
        if self.tuner.verbose:
            self._print_burnin_report()
        if self.outopts["destination"] == "iostream":
            # the samples are produced on the device; the CSV files of the reference's iostream destination are
            # written from the fetched NState when the run ends (README.md:117-146)
            from .iostream import write_job_output
            self.iostream = write_job_output(self, self._fetch_nstate())
        return self

    def _report_horizon(self):
        """transitions over which a verbose tuner reports: nadapt for DualAveragingMCTuner with HMC (iterate/HMC.jl:225-248),
        burnin otherwise -- NUTS keeps the burn-in condition with either tuner (iterate/NUTS.jl:406-447)"""
        if isinstance(self.tuner, DualAveragingMCTuner) and not isinstance(self.sampler, NUTS):
            return self.tuner.nadapt
        return self.range.burnin

    @property
    def burnin_rates(self):
        """(nchains, nperiods): the acceptance rate of every burn-in period of every chain, recorded by the kernels of
        a job with a verbose tuner -- the numbers the reference prints per period (iterate/HMC.jl:211-221,
        iterate/MALA.jl:138-148, iterate/MH.jl:126-139).  NaN where a period did not close."""
        if not self.tuner.verbose:
            raise L.KlaraError(L.KLB_ESTATE, "burn-in rates are recorded for verbose tuners only")
        nper = self._report_horizon() // self.tuner.period
        if nper == 0:
            return np.empty((self.nchains, 0))
        return self._fetch(L.OUT_TUNE_RATES, (self.nchains, nper))

    def _print_burnin_report(self, file=None):
        """println("Burnin iteration ", fmt_iter(totproposed), " of ", burnin, ": ", fmt_perc(100*rate), " % acceptance rate")
        once per period: fmt_iter = %<ndigits(burnin)>d, fmt_perc = %6.2f (src/format.jl, BasicMCJob.jl:90-101).  One
        chain prints the reference's line; a batch prints the mean over chains with the range."""
        rates = self.burnin_rates
        horizon = self._report_horizon()
        nd = len(str(horizon))
        for k in range(rates.shape[1]):
            r = rates[:, k]
            if np.isnan(r).all():
                continue
            line = "Burnin iteration %*d of %d: %6.2f %% acceptance rate" % (nd, (k + 1) * self.tuner.period, horizon,
                                                                           100 * float(np.nanmean(r)))
            if self.nchains > 1:
                line += " (mean of %d chains, %.2f .. %.2f)" % (self.nchains, 100 * np.nanmin(r), 100 * np.nanmax(r))
            print(line, file=file)

    def run_async(self):
        self._call("run_async")
        self.count = self.range.npoststeps

    def sync(self):
        self._call("sync")

    def run_host(self, x0=None, outputs=None, nslices=0):
        """reset(job, x0); run(job); output(job) in one pipelined call (klb_job_run_host): the chains go through in
        `nslices` slices on their own streams so that host->device copies, kernels and device->host copies overlap.
        `outputs` maps field codes (klara_b200._lib.OUT_*) to writable C-contiguous numpy arrays -- ideally views of
        pinned memory (klb_host_alloc) -- that receive the fields.  Results are identical to the three separate calls."""
        outputs = outputs or {}
        arr = (L.KlbHostField * max(1, len(outputs)))()
        for i, (field, buf) in enumerate(outputs.items()):
            if not (isinstance(buf, np.ndarray) and buf.flags.c_contiguous and buf.flags.writeable):
                raise TypeError("output buffers must be writable C-contiguous numpy arrays")
            arr[i].field, arr[i].host_dst, arr[i].nbytes = field, buf.ctypes.data, buf.nbytes
        xp = None
        if x0 is not None:
            x0 = np.ascontiguousarray(np.atleast_2d(np.asarray(x0, dtype=np.float64)))
            if x0.shape != (self.nchains, self.dim):
                raise AssertionError("initial value has shape %s, job has %s" % (x0.shape, (self.nchains, self.dim)))
            xp = _ptr(x0)
        if self._m:                                 # every device runs its shard's pipeline concurrently
            L.check(L.lib().klb_multi_run_host(self._m, xp, arr, len(outputs), nslices))
        else:
            L.check(L.lib().klb_job_run_host(self._h, xp, arr, len(outputs), nslices))
        self.count = self.range.npoststeps
        return self

    def reset(self, x=None):
        """reset(job) / reset(job, x)        src/jobs/BasicMCJob.jl:187-201"""
        if x is None:
            self._call("reset")
        else:
            x = np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=np.float64)))
            if x.shape != (self.nchains, self.dim):
                raise AssertionError("reset value has shape %s, job has %s" % (x.shape, (self.nchains, self.dim)))
            self._call("set_state", _ptr(x))
        self.count = 0
        return self

    def reset_synthetic(self):
        """reset(job, x0) with the device-generated synthetic initial value (SyntheticNormal)"""
        self._call("set_state_synthetic")
        self.count = 0
        return self

    def seek(self, t):
        """position the RNG streams: the next transition is number t + 1 (klb_job_seek)"""
        self._call("seek", int(t))
        return self

    def set_chunk(self, nt):
        for h, _, _ in self._shards:
            L.check(L.lib().klb_job_set_chunk(h, nt))

    def _fetch(self, field, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        self._call("output", field, _ptr(out), out.nbytes)
        return out

    def gathered(self, field, g=0):
        """device g's copy of the closing all-gather (klb_multi_gathered_output): the final state / log-target /
        tuner step / counters of ALL chains as that device holds them after run(job)"""
        if not self._m:
            raise L.KlaraError(L.KLB_ESTATE, "the job is not sharded over devices (ngpus = 1)")
        shape, dtype = {L.OUT_STATE: ((self.nchains, self.dim), np.float64), L.OUT_STATE_LOGTARGET: ((self.nchains,), np.float64),
                        L.OUT_TUNE_STEP: ((self.nchains,), np.float64), L.OUT_TUNE_COUNTERS: ((self.nchains, 3), np.int64)}[field]
        out = np.empty(shape, dtype=dtype)
        L.check(L.lib().klb_multi_gathered_output(self._m, g, field, _ptr(out), out.nbytes))
        return out

    def output(self):
        """output(job)        src/jobs/BasicMCJob.jl:279"""
        if self.outopts["destination"] == "none":
            return None
        if self.outopts["destination"] == "iostream":
            return self.iostream            # output(job) is the VariableIOStream in the reference
        return self._fetch_nstate()

    def _fetch_nstate(self):
        N, P, d = self.nchains, self.range.npoststeps, self.dim
        ns = BasicContMuvParameterNState(d, P)
        ns._job = self                       # mean(chain), ess(chain), acceptance(chain) find the device-resident samples through it
        mon = self.outopts["monitor"]
        if "value" in mon:
            ns.value = self._fetch(L.OUT_VALUE, (N, P, d))
        if "logtarget" in mon:
            ns.logtarget = self._fetch(L.OUT_LOGTARGET, (N, P))
        if "gradlogtarget" in mon:
            ns.gradlogtarget = self._fetch(L.OUT_GRADLOGTARGET, (N, P, d))
        diag = [k for k in self.outopts["diagnostics"]]
        if diag:
            # one diagnostic: (nchains, npost); several (NUTS: accept, ndoublings, a, na): (nchains, nkeys, npost), the
            # reference's nkeys x npost matrix per chain (an Array{Any} there), in the order of outopts[:diagnostics]; numpy
            # promotes the stack to float64 when :a is among the keys (every count is exactly representable)
            vals = [self._fetch(_DIAG_FIELDS[k][0], (N, P), _DIAG_FIELDS[k][1]) for k in diag]
            ns.diagnostickeys = diag
            ns.diagnosticvalues = vals[0] if len(vals) == 1 else np.stack(vals, axis=1)
        if self.single:
            for f in ("value", "logtarget", "gradlogtarget", "diagnosticvalues"):
                a = getattr(ns, f)
                if a is not None:
                    setattr(ns, f, a[0])
        return ns

    def ess(self, to_host=True):
        """ess(output(job)): effective sample size (IMSE) of every coordinate of every chain, computed on the
        device (src/stats/convergence/ess.jl:3-14).  Returns (nchains, dim), or None when to_host=False."""
        out = np.empty((self.nchains, self.dim)) if to_host else None
        for h, lo, n in self._shards:           # statistics are per chain: every shard computes its own
            L.check(L.lib().klb_job_ess(h, _ptr(out[lo:lo + n]) if to_host else None))
        if not to_host:
            return None
        return out[0] if self.single else out

    def _stat(self, code, per_chain=False):
        out = np.empty(self.nchains if per_chain else (self.nchains, self.dim))
        for h, lo, n in self._shards:
            L.check(L.lib().klb_job_stat(h, code, _ptr(out[lo:lo + n])))
        return out[0] if self.single else out

    def mean(self):
        """mean(output(job)): per coordinate of every chain (src/stats/mean.jl:7-11); (nchains, dim)"""
        return self._stat(L.STAT_MEAN)

    def mcvar(self, vtype="imse"):
        """mcvar(output(job), Val{vtype}) for vtype "iid" (var/len, src/stats/variance/mcvar.jl:5) or "imse"
        (Geyer's initial monotone sequence estimator, mcvar.jl:75-105); (nchains, dim)"""
        vtype = str(vtype).lstrip(":")
        if vtype not in ("iid", "imse"):
            raise ValueError("mcvar on the device supports :iid and :imse, got %r" % vtype)
        return self._stat(L.STAT_MCVAR_IID if vtype == "iid" else L.STAT_MCVAR_IMSE)

    def mcse(self, vtype="imse"):
        """mcse = sqrt(mcvar) (src/stats/variance/mcvar.jl:20,112)"""
        return np.sqrt(self.mcvar(vtype))

    def iact(self):
        """iact(output(job)) = mcvar(:imse)/mcvar(:iid) (src/stats/convergence/iact.jl:3-5); (nchains, dim)"""
        return self._stat(L.STAT_IACT)

    def acceptance(self, diagnostics=True):
        """acceptance(output(job); diagnostics): mean of the :accept diagnostic, or, with diagnostics=False, the
        fraction of saved states that differ from their predecessor (src/stats/acceptance.jl:3-14,28-34); one
        number per chain"""
        return self._stat(L.STAT_ACCEPTANCE if diagnostics else L.STAT_ACCEPTANCE_VALUE, per_chain=True)

    # -- job.pstate / job.sstate.tune
    @property
    def pstate_value(self):
        v = self._fetch(L.OUT_STATE, (self.nchains, self.dim))
        return v[0] if self.single else v

    @property
    def pstate_logtarget(self):
        v = self._fetch(L.OUT_STATE_LOGTARGET, (self.nchains,))
        return v[0] if self.single else v

    @property
    def tune(self):
        cnt = self._fetch(L.OUT_TUNE_COUNTERS, (self.nchains, 3), np.int64)
        base = BasicMCTune(self._fetch(L.OUT_TUNE_STEP, (self.nchains,)), cnt[:, 0].copy(), cnt[:, 1].copy(),
                           cnt[:, 2].copy(), self._fetch(L.OUT_TUNE_RATE, (self.nchains,)))
        if isinstance(self.tuner, DualAveragingMCTuner):
            return DualAveragingMCTune(base, self._fetch(L.OUT_TUNE_DA, (self.nchains, 8)))
        return base

    # -- introspection
    def plan(self):
        p = L.KlbPlan()
        L.check(L.lib().klb_job_plan(self._h, C.byref(p)))
        return p

    @property
    def launches(self):
        return sum(L.lib().klb_job_launches(h) for h, _, _ in self._shards)

    @property
    def last_run_ms(self):
        return max(L.lib().klb_job_last_run_ms(h) for h, _, _ in self._shards)

    def device_ptr(self, field):
        p, nb = C.c_void_p(), C.c_int64()
        L.check(L.lib().klb_job_device_ptr(self._h, field, C.byref(p), C.byref(nb)))
        return p.value, nb.value


def run(job):
    """run(job), or run(jobs::Vector) = map(run, jobs)        src/jobs/jobs.jl:212"""
    if isinstance(job, (list, tuple)):
        return [run(j) for j in job]
    return job.run()


def reset(job, x=None):
    return job.reset(x)


def output(job):
    return job.output()


def _job_of(x):
    """the reference's statistics take the chain (`chain = output(job); mean(chain)`, src/stats/*.jl); here the samples live
    on the device of the job that produced them, so the NState output(job) returned carries a reference to that job and
    either may be passed"""
    if isinstance(x, BasicContMuvParameterNState):
        if getattr(x, "_job", None) is None:
            raise TypeError("this NState did not come from output(job): device statistics need the job that holds the samples")
        return x._job
    return x


def ess(job):
    """ess(chain) for the job's monitored values, on the device        src/stats/convergence/ess.jl:3-14"""
    return _job_of(job).ess()


def mean(job):
    return _job_of(job).mean()


def mcvar(job, vtype="imse"):
    return _job_of(job).mcvar(vtype)


def mcse(job, vtype="imse"):
    return _job_of(job).mcse(vtype)


def iact(job):
    return _job_of(job).iact()


def acceptance(job, diagnostics=True):
    return _job_of(job).acceptance(diagnostics)


def diagnostics(nstate):
    """diagnostics(chain): Dict(zip(diagnostickeys, rows of diagnosticvalues))        src/nstates/ParameterNStates/ParameterNStates.jl:14-15
    key -> (npost,) array for a single chain, (nchains, npost) for a batch; e.g. the swiss NUTS example's
    `diags = diagnostics(chain); mean(diags[:a]./diags[:na])`"""
    keys, dv = list(nstate.diagnostickeys), nstate.diagnosticvalues
    if not keys:
        return {}
    if len(keys) == 1:
        return {keys[0]: dv}
    return {k: dv[..., q, :] for q, k in enumerate(keys)}
