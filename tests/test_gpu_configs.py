"""BASELINE.json's configurations at their STATED parameters, on the GPU at full width,
against the oracle on a chain subset (bit for bit) -- and the CUDA path against the independent numpy / libm twin
(oracle/twin.py) at north_star's tolerance: accept/reject identical, log-target / gradient / values within 1e-6."""
import numpy as np
import pytest

from helpers import SAMPLERS, ar1_precision, assert_same, build_pair, synthetic_x0
from oracle import oracle as O
from twin_helpers import check_against_twin, twin_cfg, twin_target

pytestmark = pytest.mark.gpu
SEED = 20240925


def _x0_subset(chains, dim):
    return np.stack([O.normals(SEED, int(c), 0, dim) for c in chains])


def _full_vs_subset(K, sampler, target, N, dim, nsteps, burnin, chains, smp, tuner=None, tparams=None, oracle_tuner=O.VANILLA,
                    **okw):
    """GPU: all N chains of the configuration (monitor logtarget + accept: the value matrix of the full job stays in
    HBM), x0 from the Philox streams (seed, chain, 0).  Oracle: the chains `chains` only, keyed by their global index.
    Compared bit for bit: every saved log-target and accept flag, the final state, its log-target, the tuner records."""
    x0 = np.empty((N, dim))
    for c0 in range(0, N, 4096):
        x0[c0:c0 + 4096] = synthetic_x0(SEED, min(4096, N - c0), dim, c0)
    p = K.BasicContMuvParameter("p", logtarget=target)
    job = K.BasicMCJob(K.likelihood_model(p, False), smp, K.BasicMCRange(nsteps=nsteps, burnin=burnin), {"p": x0}, tuner=tuner,
                       outopts={"monitor": ["logtarget"], "diagnostics": ["accept"]}, seed=SEED)
    job.run()
    out = job.output()
    xf, ltf, tn = job.pstate_value, job.pstate_logtarget, job.tune
    tcode = {"iso": O.ISO, "dense": O.DENSE, "rosen": O.ROSEN}[okw.pop("tname")]
    lo = 0
    # the oracle takes a contiguous block of global chain indices: run it once per contiguous run of `chains`
    chains = np.asarray(chains)
    breaks = np.nonzero(np.diff(chains) != 1)[0] + 1
    for blk in np.split(chains, breaks):
        cfg = O.make_config(SAMPLERS[sampler], tcode, len(blk), dim, nsteps, burnin, 1, okw["step"], okw.get("nleaps", 1),
                            oracle_tuner, okw.get("target_rate", 0.574), 7.0, 100, 0, 2, 1, SEED, int(blk[0]), 0, 0,
                            job.plan().nv, O.max_threads())
        ref = O.run(cfg, x0[blk], tparams)
        assert_same("logtarget", out.logtarget[blk], ref["logtarget"])
        assert_same("accept", out.diagnosticvalues[blk], ref["accept"])
        assert_same("final state", xf[blk], ref["x"])
        assert_same("final logtarget", ltf[blk], ref["logtarget_state"])
        assert_same("tune.step", tn.step[blk], ref["tune"]["step"])
        assert_same("tune.accepted", tn.accepted[blk], ref["tune"]["accepted"])
        assert_same("tune.totproposed", tn.totproposed[blk], ref["tune"]["totproposed"])
        lo += len(blk)
    return job, out


def test_c2_mala_stated_parameters(K):
    """C2: MALA(driftstep = 0.9), -z.z, 4096 chains x 128, nsteps 2000, burnin 1000        iterate/MALA.jl:78-152"""
    chains = np.concatenate([np.arange(0, 48), np.arange(4096 - 16, 4096)])
    job, out = _full_vs_subset(K, "MALA", K.IsoGaussian(), 4096, 128, 2000, 1000, chains, K.MALA(0.9), tname="iso", step=0.9)
    assert out.logtarget.shape == (4096, 1000)
    assert out.diagnosticvalues.mean() < 0.05          # driftstep 0.9 at d = 128 hardly ever accepts


def test_c3_hmc_stated_parameters_1024_chains(K):
    """C3: HMC(0.05, 10), -z.z, 65 536 chains x 1024, nsteps 200, burnin 100; oracle on 1024 of the chains
    (iterate/HMC.jl:124-224, samplers.jl:122-134)"""
    chains = np.concatenate([np.arange(0, 512), np.arange(30000, 30256), np.arange(65536 - 256, 65536)])
    job, out = _full_vs_subset(K, "HMC", K.IsoGaussian(), 65536, 1024, 200, 100, chains, K.HMC(0.05, 10), tname="iso",
                               step=0.05, nleaps=10)
    assert out.diagnosticvalues.mean() > 0.9
    # the monitored values of the same chains (a 1024-chain shard with global chain indices) against the oracle
    sub = np.arange(30000, 30256)
    x0 = _x0_subset(sub, 1024)
    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    shard = K.BasicMCJob(K.likelihood_model(p, False), K.HMC(0.05, 10), K.BasicMCRange(nsteps=200, burnin=100), {"p": x0},
                         outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept"]}, seed=SEED, chain_offset=30000)
    shard.run()
    cfg = O.make_config(O.HMC, O.ISO, 256, 1024, 200, 100, 1, 0.05, 10, O.VANILLA, 0.574, 7.0, 100, 0, 3, 1, SEED, 30000, 0, 0,
                        shard.plan().nv, O.max_threads())
    ref = O.run(cfg, x0)
    assert_same("value", shard.output().value, ref["value"])
    assert_same("shard logtarget == full job", shard.output().logtarget, out.logtarget[sub])


def test_c4_hmc_dense_stated_parameters(K):
    """C4: HMC(0.02, 20), -z'Cz with C = inv(AR(1), rho = 0.8), 16 384 chains x 512, nsteps 200, burnin 100 (DMMA path)"""
    C = ar1_precision(512)
    chains = np.concatenate([np.arange(0, 16), np.arange(16384 - 8, 16384)])
    job, out = _full_vs_subset(K, "HMC", K.DenseGaussian(C), 16384, 512, 200, 100, chains, K.HMC(0.02, 20), tname="dense",
                               tparams=C.reshape(-1), step=0.02, nleaps=20)
    assert out.diagnosticvalues.mean() > 0.5


def test_c5_mala_tuned_rosenbrock_stated_parameters(K):
    """C5: MALA(0.01) + AcceptanceRateMCTuner(0.574), paired Rosenbrock, 32 768 chains x 256, nsteps 2000, burnin 1000
    (iterate/MALA.jl:130-152, src/tuners/AcceptanceRateMCTuner.jl:46)"""
    chains = np.concatenate([np.arange(0, 48), np.arange(32768 - 16, 32768)])
    job, out = _full_vs_subset(K, "MALA", K.Rosenbrock(1.0, 100.0, 0.05), 32768, 256, 2000, 1000, chains, K.MALA(0.01),
                               tuner=K.AcceptanceRateMCTuner(0.574), tparams=np.array([1.0, 100.0, 0.05]),
                               oracle_tuner=O.ACCRATE, tname="rosen", step=0.01, target_rate=0.574)
    acc = out.diagnosticvalues.mean()
    assert 0.45 < acc < 0.7, acc                       # tuned towards 0.574
    assert len(np.unique(job.tune.step)) > 1000        # every chain has its own tuned step


# ------------------------------------------------------------------ CUDA path against the independent twin
TWIN_CASES = [
    ("HMC", "iso", 1024, dict(step=0.05, nleaps=10)),
    ("HMC", "iso", 700, dict(step=0.05, nleaps=7)),
    ("HMC", "shifted", 100, dict(step=0.08, nleaps=6)),
    ("HMC", "dense", 64, dict(step=0.05, nleaps=8)),
    ("HMC", "dense", 30, dict(step=0.05, nleaps=8)),
    ("HMC", "rosen", 96, dict(step=0.01, nleaps=6)),
    ("HMC", "logit", 4, dict(step=0.02, nleaps=5)),
    ("MALA", "iso", 128, dict(step=0.05)),
    ("MALA", "shifted", 33, dict(step=0.1)),
    ("MALA", "dense", 32, dict(step=0.02)),
    ("MALA", "rosen", 256, dict(step=0.002)),
    ("MALA", "logit", 4, dict(step=0.005)),
    ("MH", "iso", 2, dict(sigma=1.0)),
    ("MH", "iso", 1024, dict(sigma=0.02)),
    ("MH", "shifted", 65, dict(sigma=0.1)),
    ("MH", "dense", 16, dict(sigma=0.1)),
    ("MH", "rosen", 32, dict(sigma=0.05)),
    ("MH", "logit", 5, dict(sigma=0.05)),
]


@pytest.mark.parametrize("arith", ["reference", "fma"])
@pytest.mark.parametrize("sampler,target,dim,kw", TWIN_CASES, ids=["%s-%s-%d" % c[:3] for c in TWIN_CASES])
def test_cuda_path_agrees_with_independent_twin(K, sampler, target, dim, kw, arith):
    """GPU output vs the numpy / libm twin, both free-running from the same x0 and the same Philox streams: identical
    accept/reject sequences; values, log-targets and gradients within 1e-6 relative.  The twin's dot products are
    numpy's (BLAS order), its exp / log are libm's: this is the tolerance statement of north_star against an
    independently ordered computation, in both arithmetic modes of the kernels."""
    kw = dict(kw)
    nchains, nsteps = 6, 40
    sigma = np.full(dim, kw.pop("sigma")) if sampler == "MH" else None
    mon = ("value", "logtarget") if sampler == "MH" else ("value", "logtarget", "gradlogtarget")
    x0 = synthetic_x0(777, nchains, dim) * (0.3 if target == "dense" else 1.0)
    job, cfg, x0, tp, sg = build_pair(K, sampler, target, nchains=nchains, dim=dim, nsteps=nsteps, seed=777, sigma=sigma,
                                      monitor=mon, arith=arith, x0=x0, **kw)
    job.run()
    out = job.output()
    tcfg = twin_cfg(sampler, nsteps, seed=777, sigma=sigma, **kw)
    w = check_against_twin("cuda", tcfg, twin_target(target, dim, tp), x0, range(nchains), out.value, out.logtarget,
                           out.diagnosticvalues, grad=out.gradlogtarget)
    assert w["flips"] == 0 and w["transitions"] == nchains * nsteps
    assert w["value"] < 1e-9 and w["logtarget"] < 1e-9 and w["grad"] < 1e-9      # observed ~1e-13; the bar is 1e-6


@pytest.mark.parametrize("name,sampler,target,dim,nsteps,kw", [
    ("C2", "MALA", "iso", 128, 400, dict(step=0.9)),
    ("C3", "HMC", "iso", 1024, 100, dict(step=0.05, nleaps=10)),
    ("C4", "HMC", "dense", 512, 8, dict(step=0.02, nleaps=20)),
    ("C5", "MALA", "rosen", 256, 400, dict(step=0.01, tuner="accrate", target_rate=0.574, period=100)),
])
def test_cuda_path_vs_twin_teacher_forced_at_baseline_parameters(K, name, sampler, target, dim, nsteps, kw):
    """the four GPU configurations of BASELINE.json at their stated sampler parameters: every twin transition restarts
    from the GPU's previous state, so long chains cannot compound rounding differences"""
    nchains = 3
    job, cfg, x0, tp, sg = build_pair(K, sampler, target, nchains=nchains, dim=dim, nsteps=nsteps, seed=SEED, **dict(kw))
    job.run()
    out = job.output()
    w = check_against_twin(name, twin_cfg(sampler, nsteps, seed=SEED, **dict(kw)), twin_target(target, dim, tp), x0,
                           range(nchains), out.value, out.logtarget, out.diagnosticvalues, forced=True)
    assert w["transitions"] == nchains * nsteps and w["flips"] == 0


def test_cuda_tuners_agree_with_twin(K):
    """AcceptanceRateMCTuner and DualAveragingMCTuner records against the twin's libm arithmetic"""
    for sampler, tuner, kw in (("HMC", "accrate", dict(step=0.07, nleaps=4, target_rate=0.7, period=10, burnin=60)),
                               ("MALA", "accrate", dict(step=0.3, target_rate=0.7, period=10, burnin=60)),
                               ("HMC", "dualavg", dict(step=0.1, nleaps=8, target_rate=0.65, nadapt=30))):
        job, cfg, x0, tp, sg = build_pair(K, sampler, "iso", nchains=4, dim=20, nsteps=90, seed=99, tuner=tuner, **kw)
        job.run()
        out = job.output()
        w = check_against_twin("cuda-" + tuner, twin_cfg(sampler, 90, seed=99, tuner=tuner, **kw), twin_target("iso", 20, None),
                               x0, range(4), out.value, out.logtarget, out.diagnosticvalues, final_step=job.tune.step)
        assert w["flips"] == 0
