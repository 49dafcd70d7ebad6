"""GPU parity suite: the CUDA path (through the C ABI) against the CPU oracle, bit for bit, on the
same seeded inputs.  Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import assert_same, build_pair, compare_run, synthetic_x0

pytestmark = pytest.mark.gpu


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ primitives
@pytest.mark.parametrize("seed,chain,t", [(0, 0, 0), (20240925, 17, 3), (2**63 + 5, 2**32 - 1, 2**40 + 9)])
def test_device_normals_match_oracle(K, O, seed, chain, t):
    n = 1 << 17
    out = np.empty(n)
    K._lib.check(K._lib.lib().klb_debug_normals(0, seed, chain, t, n, _ptr(out)))
    assert_same("normals", out, O.normals(seed, chain, t, n))


def test_device_uniform_matches_oracle(K, O):
    for seed, chain, t in [(1, 0, 1), (99, 65535, 200), (2**64 - 1, 12345, 2**33)]:
        out = np.empty(1)
        K._lib.check(K._lib.lib().klb_debug_uniform(0, seed, chain, t, _ptr(out)))
        assert out[0] == O.uniform(seed, chain, t)
        assert 0.0 <= out[0] < 1.0


def test_device_exp_log_match_oracle(K, O):
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(-750, 710, 200000), rng.uniform(-3, 3, 200000),
                         np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 709.782712893384, 709.79, -745.2, -745.0,
                                   -708.4, 1e-300, -1e-300])])
    out = np.empty_like(xs)
    K._lib.check(K._lib.lib().klb_debug_math(0, 0, xs.size, _ptr(xs), _ptr(out)))
    assert_same("exp", out, np.array([O.exp(x) for x in xs]))
    xl = np.concatenate([rng.uniform(0, 1, 200000), 10.0 ** rng.uniform(-320, 308, 100000),
                         rng.uniform(0.98, 1.02, 100000),
                         np.array([0.0, -0.0, 1.0, np.inf, -1.0, np.nan, 5e-324, 2.2250738585072014e-308])])
    out = np.empty_like(xl)
    K._lib.check(K._lib.lib().klb_debug_math(0, 1, xl.size, _ptr(xl), _ptr(out)))
    assert_same("log", out, np.array([O.log(x) for x in xl]))


# ------------------------------------------------------------------ samplers vs oracle
DIMS = [2, 7, 64, 100, 128, 250, 512, 1024, 1500, 2048, 4096]


@pytest.mark.parametrize("arith", ["reference", "fma"])
@pytest.mark.parametrize("dim", DIMS)
def test_hmc_iso_bit_exact(K, dim, arith):
    eps = 0.3 / np.sqrt(dim) if dim > 4 else 0.2
    job, cfg, x0, tp, sg = build_pair(K, "HMC", "iso", nchains=37, dim=dim, nsteps=50, burnin=20, thinning=3,
                                      step=eps, nleaps=10, seed=20240925, arith=arith)
    out, ref = compare_run(job, cfg, x0, tp, sg)
    acc = out.diagnosticvalues.mean()
    assert 0.05 < acc <= 1.0


@pytest.mark.parametrize("arith", ["reference", "fma"])
@pytest.mark.parametrize("dim", [2, 33, 128, 256, 1000, 3000])
def test_mala_iso_bit_exact(K, dim, arith):
    job, cfg, x0, tp, sg = build_pair(K, "MALA", "iso", nchains=41, dim=dim, nsteps=80, burnin=30,
                                      step=0.9 / dim ** (1 / 3), seed=7, arith=arith,
                                      monitor=("value", "logtarget", "gradlogtarget"))
    compare_run(job, cfg, x0, tp, sg)


@pytest.mark.parametrize("arith", ["reference", "fma"])
@pytest.mark.parametrize("dim", [1, 2, 65, 512])
def test_mh_iso_bit_exact(K, dim, arith):
    sigma = np.linspace(0.05, 0.4, dim) if dim > 1 else np.array([0.7])
    job, cfg, x0, tp, sg = build_pair(K, "MH", "iso", nchains=29, dim=dim, nsteps=120, burnin=20, thinning=2,
                                      seed=99, arith=arith, sigma=sigma, verbose=True)
    compare_run(job, cfg, x0, tp, sg)


@pytest.mark.parametrize("sampler", ["HMC", "MALA", "MH"])
@pytest.mark.parametrize("target", ["shifted", "rosen"])
def test_other_targets_bit_exact(K, sampler, target):
    dim = 96
    step = {"HMC": 0.02, "MALA": 0.002, "MH": 0.1}[sampler]
    mon = ("value", "logtarget") if sampler == "MH" else ("value", "logtarget", "gradlogtarget")
    for arith in ("reference", "fma"):
        job, cfg, x0, tp, sg = build_pair(K, sampler, target, nchains=33, dim=dim, nsteps=60, burnin=10, step=step,
                                          nleaps=7, seed=4242, arith=arith, monitor=mon,
                                          sigma=np.full(dim, 0.02))
        compare_run(job, cfg, x0, tp, sg)


@pytest.mark.parametrize("arith", ["reference", "fma"])
@pytest.mark.parametrize("dim", [1, 2, 3, 7, 64, 65, 128, 130, 255, 256, 511, 512])
@pytest.mark.parametrize("sampler", ["HMC", "MALA", "MH"])
def test_dense_precision_target_bit_exact(K, sampler, dim, arith):
    """-z'Cz, -2Cz with C = inv(AR(1) covariance): the matrix-vector kernels (klb_dense.cuh) against the oracle;
    13 chains = one full CTA tile of 8 plus a ragged one"""
    step = {"HMC": 0.05, "MALA": 0.02, "MH": 0.1}[sampler]
    mon = ("value", "logtarget") if sampler == "MH" else ("value", "logtarget", "gradlogtarget")
    job, cfg, x0, tp, sg = build_pair(K, sampler, "dense", nchains=13, dim=dim, nsteps=30, burnin=6, thinning=2,
                                      step=step, nleaps=5, seed=99, arith=arith, monitor=mon,
                                      sigma=np.full(dim, 0.05), tuner="accrate", period=5, target_rate=0.6)
    out, ref = compare_run(job, cfg, x0, tp, sg)
    assert out.diagnosticvalues.mean() > 0.05


def test_dense_bivariate_normal_example(K, O):
    """doc/examples/BivariateNormal/MALA/function/analytical.jl: MALA(0.3), C = inv([1 .8; .8 1]), p0 = [1.25, 3.11],
    nsteps 10000, burnin 1000, monitor value/logtarget/gradlogtarget, diagnostics accept"""
    C = np.linalg.inv(np.array([[1.0, 0.8], [0.8, 1.0]]))
    C = (C + C.T) / 2
    p = K.BasicContMuvParameter("p", logtarget=K.DenseGaussian(C))
    job = K.BasicMCJob(K.likelihood_model(p, False), K.MALA(0.3), K.BasicMCRange(nsteps=10000, burnin=1000),
                       {"p": [1.25, 3.11]},
                       outopts={"monitor": ["value", "logtarget", "gradlogtarget"], "diagnostics": ["accept"]}, seed=7)
    K.run(job)
    chain = K.output(job)
    cfg = O.make_config(O.MALA, O.DENSE, 1, 2, 10000, 1000, step=0.3, monitor=7, diagnostics=1, seed=7, nv=job.plan().nv)
    ref = O.run(cfg, np.array([[1.25, 3.11]]), C.reshape(-1))
    assert_same("value", chain.value, ref["value"][0])
    assert_same("gradlogtarget", chain.gradlogtarget, ref["gradlogtarget"][0])
    assert_same("accept", chain.diagnosticvalues, ref["accept"][0])
    # exp(-z'Cz) is N(0, Sigma/2) with Sigma = [1 .8; .8 1]
    cov = np.cov(chain.value.T)
    assert abs(cov[0, 0] - 0.5) < 0.08 and abs(cov[0, 1] - 0.4) < 0.08 and abs(chain.value.mean()) < 0.1


def test_dense_bivariate_normal_example_through_the_hyperparameter_vertex(K, O):
    """the same example as the reference writes it: C = Hyperparameter(:C), closures of (p, v) with v[1] = C,
    nkeys = 2, model = GenericModel([C, p], isindexed=false), v0 = Dict(:C => inv(...), :p => [1.25, 3.11])"""
    C = np.linalg.inv(np.array([[1.0, 0.8], [0.8, 1.0]]))
    C = (C + C.T) / 2
    d = K.DenseGaussian()
    p = K.BasicContMuvParameter("p", logtarget=d, gradlogtarget=d.gradient, nkeys=2)
    model = K.GenericModel([K.Hyperparameter("C"), p], isindexed=False)
    job = K.BasicMCJob(model, K.MALA(0.3), K.BasicMCRange(nsteps=400, burnin=100), {"C": C, "p": [1.25, 3.11]},
                       outopts={"monitor": ["value", "logtarget", "gradlogtarget"], "diagnostics": ["accept"]}, seed=7)
    K.run(job)
    chain = K.output(job)
    cfg = O.make_config(O.MALA, O.DENSE, 1, 2, 400, 100, step=0.3, monitor=7, diagnostics=1, seed=7, nv=job.plan().nv)
    ref = O.run(cfg, np.array([[1.25, 3.11]]), C.reshape(-1))
    assert_same("value", chain.value, ref["value"][0])
    assert_same("gradlogtarget", chain.gradlogtarget, ref["gradlogtarget"][0])
    assert_same("accept", chain.diagnosticvalues, ref["accept"][0])


def test_dense_mma_and_dfma_kernels_agree(K, monkeypatch):
    """HMC on the dense target: the tensor-pipe kernel (DMMA, default for dim 64/128/256/512) and the DFMA
    register-tile kernel (KLB_DENSE_MMA=0) produce the same bits -- both follow the increasing-j fma chain"""
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("KLB_DENSE_MMA", flag)
        job, cfg, x0, tp, sg = build_pair(K, "HMC", "dense", nchains=37, dim=256, nsteps=12, burnin=2, step=0.05,
                                          nleaps=4, seed=17)
        job.run()
        outs.append((job.output().value, job.output().diagnosticvalues, job.plan().regs_per_thread))
    assert_same("value", outs[0][0], outs[1][0])
    assert_same("accept", outs[0][1], outs[1][1])


def test_hmc_warp_specialised_and_fused_kernels_agree(K, monkeypatch):
    """dim 257..1024 HMC defaults to the producer/consumer kernel (klb_hmc_ws.cuh); KLB_HMC_WS=0 selects the fused
    single-role kernel.  Same bits, including a ragged last CTA (nchains not a multiple of 4) and a masked dim."""
    for dim in (1024, 700, 512):
        outs = []
        for flag in ("1", "0"):
            monkeypatch.setenv("KLB_HMC_WS", flag)
            job, cfg, x0, tp, sg = build_pair(K, "HMC", "iso", nchains=37, dim=dim, nsteps=14, burnin=3, thinning=2,
                                              step=0.3 / np.sqrt(dim), nleaps=5, seed=23, tuner="accrate", period=4,
                                              target_rate=0.9)
            job.run()
            outs.append((job.output().value, job.output().diagnosticvalues, job.tune.step, job.plan().warps_per_block))
        assert outs[0][3] == 8 and outs[1][3] == 4
        assert_same("value", outs[0][0], outs[1][0])
        assert_same("accept", outs[0][1], outs[1][1])
        assert_same("step", outs[0][2], outs[1][2])


def test_dense_needs_symmetric_matrix(K):
    C = np.array([[1.0, 0.5], [0.25, 1.0]])
    p = K.BasicContMuvParameter("p", logtarget=K.DenseGaussian(C))
    with pytest.raises(K.KlaraError, match="symmetric"):
        K.BasicMCJob(K.likelihood_model(p, False), K.MALA(0.3), K.BasicMCRange(nsteps=10), {"p": [1.0, 2.0]})


@pytest.mark.parametrize("sampler,step", [("HMC", 0.01), ("MALA", 0.5)])
def test_acceptance_rate_tuner_bit_exact(K, sampler, step):
    """burn-in adaptation: step *= logistic_rate_score(rate - target) every `period` proposals while
    totproposed <= burnin; per-chain records must match (iterate/HMC.jl:203-224)"""
    job, cfg, x0, tp, sg = build_pair(K, sampler, "iso", nchains=48, dim=64, nsteps=260, burnin=200, step=step,
                                      nleaps=5, tuner="accrate", target_rate=0.574 if sampler == "MALA" else 0.8,
                                      period=25, seed=31337)
    out, ref = compare_run(job, cfg, x0, tp, sg)
    tn = job.tune
    assert (tn.totproposed == 225).all()       # period*(1 + floor(burnin/period)) = 25*9
    assert (tn.proposed == 60).all()           # keeps counting after burn-in
    assert not np.allclose(tn.step, step)      # adapted


@pytest.mark.parametrize("sampler,step", [("HMC", 0.12), ("MALA", 0.4)])
def test_acceptance_rate_tuner_erf_score_bit_exact(K, sampler, step):
    """AcceptanceRateMCTuner(targetrate, score=erf_rate_score): step *= erf(3*(rate - target)) + 1
    (src/tuners/AcceptanceRateMCTuner.jl:17,46); klb_erf is the same double-double series on both sides"""
    job, cfg, x0, tp, sg = build_pair(K, sampler, "iso", nchains=33, dim=40, nsteps=130, burnin=100, step=step, nleaps=4,
                                      seed=17, tuner="accrate", target_rate=0.65, period=10, score="erf")
    out, ref = compare_run(job, cfg, x0, tp, sg)
    assert len(np.unique(job.tune.step)) > 10 and (job.tune.step != step).all()
    logi, *_ = build_pair(K, sampler, "iso", nchains=33, dim=40, nsteps=130, burnin=100, step=step, nleaps=4,
                          seed=17, tuner="accrate", target_rate=0.65, period=10)
    logi.run()
    assert not np.array_equal(logi.tune.step, job.tune.step)       # the two score functions really differ


def test_verbose_vanilla_counters(K):
    job, cfg, x0, tp, sg = build_pair(K, "HMC", "iso", nchains=8, dim=16, nsteps=130, burnin=100, step=0.1,
                                      nleaps=4, tuner="vanilla", verbose=True, period=50, seed=5)
    compare_run(job, cfg, x0, tp, sg)
    tn = job.tune
    assert (tn.totproposed == 150).all() and (tn.proposed == 30).all()


def test_chunked_launches_are_invariant(K):
    """one launch per transition (lockstep) == one launch for the whole run"""
    res = []
    for chunk in (0, 1, 7):
        job, cfg, x0, tp, sg = build_pair(K, "HMC", "iso", nchains=21, dim=130, nsteps=40, burnin=11, thinning=2,
                                          step=0.03, nleaps=6, seed=77)
        job.set_chunk(chunk)
        out, ref = compare_run(job, cfg, x0, tp, sg)
        res.append(out.value)
    assert_same("chunk 1 vs whole", res[1], res[0])
    assert_same("chunk 7 vs whole", res[2], res[0])


def test_sharding_is_invariant(K):
    """chains [8, 24) of a 32-chain job computed as their own shard give the same bits (global chain
    index in the RNG counter): the multi-GPU invariance"""
    full, cfg, x0, tp, sg = build_pair(K, "MALA", "iso", nchains=32, dim=40, nsteps=30, burnin=5, step=0.2, seed=3)
    full.run()
    v = full.output().value
    shard, cfg2, x02, _, _ = build_pair(K, "MALA", "iso", nchains=16, dim=40, nsteps=30, burnin=5, step=0.2, seed=3,
                                        chain_offset=8)
    assert_same("x0 shard", x02, x0[8:24])
    shard.run()
    assert_same("shard", shard.output().value, v[8:24])


def test_reset_and_second_run(K):
    """run twice without reset overruns the NState (BoundsError in the reference); after reset(job) the chain
    continues from its current state with fresh randomness and zeroed tuner records"""
    job, cfg, x0, tp, sg = build_pair(K, "HMC", "iso", nchains=10, dim=20, nsteps=30, burnin=10, step=0.1, nleaps=3,
                                      seed=11, tuner="accrate", period=5)
    out1, ref1 = compare_run(job, cfg, x0, tp, sg)
    with pytest.raises(K.KlaraError):
        job.run()
    job.reset()
    tn = job.tune
    assert (tn.accepted == 0).all() and (tn.proposed == 0).all() and (tn.totproposed == 5).all()
    assert np.isnan(tn.rate).all() and (tn.step == 0.1).all()
    out2, ref2 = compare_run(job, cfg, ref1["x"], tp, sg, t0=30)
    assert not np.array_equal(out1.value, out2.value)
    # reset(job, x): restart from a new value
    job.reset(x0)
    assert_same("state after reset(job, x)", job.pstate_value, x0)


def test_nonfinite_initial_value_is_rejected(K):
    x0 = synthetic_x0(1, 6, 10)
    x0[4, 3] = np.inf
    with pytest.raises(K.KlaraError) as ei:
        build_pair(K, "HMC", "iso", nchains=6, dim=10, nsteps=5, x0=x0)
    assert ei.value.code == K._lib.KLB_ENOTFINITE and "chain 4" in str(ei.value)


def test_nan_proposals_reject(K):
    """a divergent trajectory (huge step) yields NaN/-Inf ratios, which must reject silently
    (iterate/HMC.jl:163-165: rand() < NaN is false)"""
    job, cfg, x0, tp, sg = build_pair(K, "HMC", "iso", nchains=5, dim=8, nsteps=12, step=1e200, nleaps=3, seed=2)
    out, ref = compare_run(job, cfg, x0, tp, sg)
    assert out.diagnosticvalues.sum() == 0
    assert_same("state unchanged", job.pstate_value, x0)


def test_readme_mh_example(K, O):
    """README.md:23-55: MH(ones(2)), -z.z, one chain, nsteps 10000, burnin 1000, v0 = [5.1, -0.9]"""
    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    model = K.likelihood_model(p, False)
    job = K.BasicMCJob(model, K.MH(np.ones(2)), K.BasicMCRange(nsteps=10000, burnin=1000), {"p": [5.1, -0.9]},
                       seed=2024)
    K.run(job)
    chain = K.output(job)
    assert chain.value.shape == (9000, 2)
    cfg = O.make_config(O.MH, O.ISO, 1, 2, 10000, 1000, seed=2024, nv=job.plan().nv)
    ref = O.run(cfg, np.array([[5.1, -0.9]]), None, np.ones(2))
    assert_same("README chain", chain.value, ref["value"][0])
    # target is N(0, 1/2 I)
    assert abs(chain.value.mean()) < 0.1 and abs(chain.value.var() - 0.5) < 0.1


def test_full_size_properties_c3(K):
    """BASELINE config C3 at full width (65 536 chains x 1024, HMC L=10) for a few transitions:
    size-independent properties -- stored log-target equals -z.z of the stored value, rejected
    transitions repeat the previous sample, acceptance is high at eps = 0.05/sqrt(d)-scale steps"""
    N, d = 65536, 1024
    x0 = np.random.default_rng(0).normal(size=(N, d)) * np.sqrt(0.5)
    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    job = K.BasicMCJob(K.likelihood_model(p, False), K.HMC(0.01, 10), K.BasicMCRange(nsteps=6, burnin=3),
                       {"p": x0}, outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept"]}, seed=1)
    job.run()
    out = job.output()
    v, lt, acc = out.value, out.logtarget, out.diagnosticvalues
    np.testing.assert_allclose(lt, -(v * v).sum(-1), rtol=1e-12)
    rej = acc[:, 1:] == 0
    assert np.array_equal(v[:, 1:][rej], v[:, :-1][rej])
    assert acc.mean() > 0.6
    assert_same("final state == last sample", job.pstate_value, v[:, -1])


def test_ess_on_device_matches_oracle(K, O):
    """ess(chain, :imse) per coordinate (src/stats/convergence/ess.jl, variance/mcvar.jl:75-105): device == oracle
    bit for bit on the values the job stored; MALA with a small step gives visibly correlated chains"""
    # npost = 300, 100, 300, 700: the shared-memory tile kernel with 32 / 128 / 32 coordinates per CTA and the
    # global-memory fallback
    for sampler, step, dim, nsteps in (("MALA", 0.05, 24, 400), ("HMC", 0.1, 130, 200), ("MH", 0.0, 7, 400),
                                       ("MALA", 0.1, 40, 800)):
        job, cfg, x0, tp, sg = build_pair(K, sampler, "iso", nchains=19, dim=dim, nsteps=nsteps, burnin=100, step=step,
                                          nleaps=4, seed=21, sigma=np.full(dim, 0.3))
        job.run()
        e_gpu = job.ess()
        v = job.output().value
        e_ref = O.ess(v)
        assert_same("ess", e_gpu, e_ref)
        assert np.isfinite(e_gpu).all() and (e_gpu > 1).all() and (e_gpu < 3 * 700).all()
    # a well-mixing chain (HMC, trajectory ~ a quarter period) has ESS of the order of the number of samples
    # (a longer trajectory makes the chain antithetic and the IMSE estimate exceeds n), a sticky one far less
    job, *_ = build_pair(K, "HMC", "iso", nchains=64, dim=16, nsteps=1100, burnin=100, step=0.25, nleaps=4, seed=5)
    job.run()
    assert 400 < job.ess().mean() < 1000
    job, *_ = build_pair(K, "MH", "iso", nchains=64, dim=16, nsteps=1100, burnin=100, sigma=np.full(16, 0.05), seed=5)
    job.run()
    assert job.ess().mean() < 60


def test_stats_on_device_match_oracle(K, O):
    """mean / mcvar(:iid) / mcvar(:imse) / ess / iact / acceptance of the stored output (src/stats/mean.jl:7-11,
    variance/mcvar.jl:5,75-105, convergence/{ess,iact}.jl, acceptance.jl): device == oracle bit for bit"""
    for sampler, step, dim, nsteps in (("MALA", 0.05, 24, 400), ("HMC", 0.1, 130, 200), ("MH", 0.0, 7, 400),
                                       ("MALA", 0.1, 40, 800)):
        job, cfg, x0, tp, sg = build_pair(K, sampler, "iso", nchains=19, dim=dim, nsteps=nsteps, burnin=100, step=step,
                                          nleaps=4, seed=22, sigma=np.full(dim, 0.3))
        job.run()
        out = job.output()
        ref = O.stats(out.value)
        assert_same("mean", job.mean(), ref["mean"])
        assert_same("mcvar iid", job.mcvar("iid"), ref["mcvar_iid"])
        assert_same("mcvar imse", job.mcvar(":imse"), ref["mcvar_imse"])
        assert_same("ess", job.ess(), ref["ess"])
        assert_same("iact", job.iact(), ref["iact"])
        assert_same("mcse", job.mcse("iid"), np.sqrt(ref["mcvar_iid"]))
        acc = out.diagnosticvalues
        assert_same("acceptance", job.acceptance(), O.acceptance(accept=acc))
        assert_same("acceptance from values", job.acceptance(diagnostics=False), O.acceptance(value=out.value))
        # an accepted continuous proposal always moves the chain: both forms agree up to the first sample
        n = acc.shape[1]
        assert np.array_equal(job.acceptance(diagnostics=False), ((acc[:, 1:] != 0).sum(axis=1) + 1) / n)
        assert np.allclose(job.iact() * job.ess(), n, rtol=1e-12)
    with pytest.raises(K.KlaraError):
        build_pair(K, "HMC", "iso", nchains=4, dim=8, nsteps=20, burnin=0)[0].mean()      # not run yet
    with pytest.raises(ValueError):
        job.mcvar("bm")


def test_iostream_destination(K, tmp_path):
    """README.md:117-146: :destination => :iostream writes value.csv / logtarget.csv / diagnosticvalues.csv"""
    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    outopts = {"monitor": ["value", "logtarget"], "diagnostics": ["accept"], "destination": "iostream",
               "filepath": str(tmp_path)}
    job = K.BasicMCJob(K.likelihood_model(p, False), K.MH(np.ones(2)), K.BasicMCRange(nsteps=300, burnin=100),
                       {"p": [5.1, -0.9]}, tuner=K.VanillaMCTuner(verbose=True), outopts=outopts, seed=3)
    K.run(job)
    stream = K.output(job)
    back = stream.read()
    ref = K.BasicMCJob(K.likelihood_model(p, False), K.MH(np.ones(2)), K.BasicMCRange(nsteps=300, burnin=100),
                       {"p": [5.1, -0.9]}, tuner=K.VanillaMCTuner(verbose=True),
                       outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept"]}, seed=3)
    chain = K.output(K.run(ref))
    assert_same("value.csv", back["value"], chain.value)
    assert_same("logtarget.csv", back["logtarget"], chain.logtarget)
    assert np.array_equal(back["diagnosticvalues"][:, 0], chain.diagnosticvalues.astype(bool))
    # batched job: one directory per chain
    job = K.BasicMCJob(K.likelihood_model(p, False), K.MALA(0.5), K.BasicMCRange(nsteps=20, burnin=5),
                       {"p": np.ones((3, 4))}, outopts={"destination": "iostream", "filepath": str(tmp_path / "b")},
                       seed=1)
    K.run(job)
    assert sorted(os.listdir(tmp_path / "b")) == ["chain1", "chain2", "chain3"]
    assert len(open(tmp_path / "b" / "chain2" / "value.csv").read().splitlines()) == 15


# ------------------------------------------------------------------ Bayesian logistic regression (klb_glm.cuh)
@pytest.mark.parametrize("arith", ["reference", "fma"])
@pytest.mark.parametrize("dim", [1, 3, 4, 9, 16])
@pytest.mark.parametrize("sampler", ["HMC", "MALA", "MH"])
def test_logit_target_bit_exact(K, sampler, dim, arith):
    """the closures of doc/examples/swiss/HMC/noadaptation/analytical.jl:11-20 on synthetic 200 x dim data:
    thread-per-chain kernels against the oracle (sequential reduction order, nv = 0); 150 chains = two full CTAs
    of 64 plus a ragged one"""
    step = {"HMC": 0.03, "MALA": 0.004, "MH": 0.0}[sampler]
    mon = ("value", "logtarget") if sampler == "MH" else ("value", "logtarget", "gradlogtarget")
    job, cfg, x0, tp, sg = build_pair(K, sampler, "logit", nchains=150, dim=dim, nsteps=40, burnin=12, thinning=3,
                                      step=step, nleaps=6, seed=77, arith=arith, monitor=mon,
                                      sigma=np.full(dim, 0.08), tuner="accrate", period=6, target_rate=0.7)
    assert job.plan().nv == 0 and job.plan().warps_per_chain == 0
    out, ref = compare_run(job, cfg, x0, tp, sg)
    assert 0.02 < out.diagnosticvalues.mean() <= 1.0


def test_logit_data_streamed_from_global_memory(K):
    """a data set too large for shared memory (6000 x 4 = 234 KB with y) is read through L1/L2: same bits"""
    rng = np.random.default_rng(3)
    from helpers import logit_data
    X, y, lam = logit_data(4, rng, ndata=6000, lam=10.0)
    x0 = synthetic_x0(5, 70, 4) * 0.1
    p = K.BasicContMuvParameter("p", logtarget=K.BayesLogit(X, y, lam))
    job = K.BasicMCJob(K.likelihood_model(p, False), K.HMC(0.005, 4), K.BasicMCRange(nsteps=8, burnin=2), {"p": x0},
                       outopts={"monitor": ["value", "logtarget", "gradlogtarget"], "diagnostics": ["accept"]}, seed=5)
    from oracle import oracle as O
    cfg = O.make_config(O.HMC, O.LOGIT, 70, 4, 8, 2, step=0.005, nleaps=4, monitor=7, diagnostics=1, seed=5,
                        nthreads=O.max_threads())
    compare_run(job, cfg, x0, O.logit_params(X, y, lam), None)


def test_swiss_example_model_form(K, O):
    """doc/examples/swiss/HMC/noadaptation/analytical.jl as written: a model with Hyperparameter(:λ), Data(:X),
    Data(:y) vertices before the parameter, whose v0 values reach the target in vertex order; HMC(0.35) needs the
    real (wide-posterior) swiss data, so the synthetic stand-in uses a smaller step.  One chain and a batch of
    chains give the same first chain."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as G
    X, y, lam = G.logit_data(4)
    t = K.BayesLogit()
    p = K.BasicContMuvParameter("p", loglikelihood=t.loglikelihood, logprior=t.logprior, gradlogtarget=t.gradient,
                                nkeys=4)
    model = K.likelihood_model([K.Hyperparameter("λ"), K.Data("X"), K.Data("y"), p], isindexed=False)
    v0 = {"λ": lam, "X": X, "y": y, "p": [5.1, -0.9, 8.2, -4.5]}
    outopts = {"monitor": ["value", "logtarget", "gradlogtarget"], "diagnostics": ["accept"]}
    job = K.BasicMCJob(model, K.HMC(0.05), K.BasicMCRange(nsteps=3000, burnin=1000), v0, outopts=outopts, seed=9)
    K.run(job)
    chain = K.output(job)
    cfg = O.make_config(O.HMC, O.LOGIT, 1, 4, 3000, 1000, step=0.05, nleaps=10, monitor=7, diagnostics=1, seed=9)
    ref = O.run(cfg, np.array([v0["p"]]), O.logit_params(X, y, lam))
    assert_same("value", chain.value, ref["value"][0])
    assert_same("gradlogtarget", chain.gradlogtarget, ref["gradlogtarget"][0])
    assert_same("accept", chain.diagnosticvalues, ref["accept"][0])
    assert 0.5 < K.acceptance(job) <= 1.0
    # posterior mean close to the mode (Newton), in units of the Laplace standard deviation
    b = np.zeros(4)
    for _ in range(50):
        mu = 1 / (1 + np.exp(-X @ b))
        H = X.T @ (X * (mu * (1 - mu))[:, None]) + np.eye(4) / lam
        b = b + np.linalg.solve(H, X.T @ (y - mu) - b / lam)
    sd = np.sqrt(np.diag(np.linalg.inv(H)))
    assert np.all(np.abs(K.mean(job) - b) < 0.5 * sd)


def test_swiss_nuts_dual_averaging_example(K, O):
    """doc/examples/swiss/NUTS/dualaveraging/analytical.jl as written (shorter range; synthetic stand-in for the data):
    NUTS(0.4, maxndoublings=7) with DualAveragingMCTuner(0.651, nadapt), monitor value / logtarget / gradlogtarget,
    diagnostics [:accept, :ndoublings, :a, :na]; the example's closing statistic mean(diags[:a]./diags[:na]) sits at the
    tuner's target.  One chain (the reference's own shape) against the oracle, bit for bit."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as G
    X, y, lam = G.logit_data(4)
    t = K.BayesLogit()
    p = K.BasicContMuvParameter("p", loglikelihood=t.loglikelihood, logprior=t.logprior, gradlogtarget=t.gradient, nkeys=4)
    model = K.likelihood_model([K.Hyperparameter("λ"), K.Data("X"), K.Data("y"), p], isindexed=False)
    v0 = {"λ": lam, "X": X, "y": y, "p": [5.1, -0.9, 8.2, -4.5]}
    outopts = {"monitor": ["value", "logtarget", "gradlogtarget"], "diagnostics": ["accept", "ndoublings", "a", "na"]}
    job = K.BasicMCJob(model, K.NUTS(0.4, maxndoublings=7), K.BasicMCRange(nsteps=300, burnin=100), v0,
                       tuner=K.DualAveragingMCTuner(0.651, 300), outopts=outopts, seed=9)
    K.run(job)
    chain = K.output(job)
    cfg = O.make_config(O.NUTS, O.LOGIT, 1, 4, 300, 100, step=0.4, tuner=O.DUALAVG, target_rate=0.651, nadapt=300, monitor=7,
                        diagnostics=15, seed=9, maxndoublings=7, nv=0)
    ref = O.run(cfg, np.array([v0["p"]]), O.logit_params(X, y, lam))
    assert_same("value", chain.value, ref["value"][0])
    assert_same("gradlogtarget", chain.gradlogtarget, ref["gradlogtarget"][0])
    diags = dict(zip(chain.diagnostickeys, chain.diagnosticvalues))          # diagnostics(chain)
    for k in ("accept", "ndoublings", "a", "na"):
        assert_same(k, diags[k], ref[k][0].astype(np.float64))
    assert abs(np.mean(diags["a"] / diags["na"]) - 0.651) < 0.08              # mean(diags[:a]./diags[:na])
    assert_same("tuned step", np.atleast_1d(job.tune.step), ref["tune"]["step"])


def test_logit_needs_its_data(K):
    p = K.BasicContMuvParameter("p", logtarget=K.BayesLogit())
    with pytest.raises(AssertionError, match="no data"):
        K.BasicMCJob(K.likelihood_model(p, False), K.MALA(0.1), K.BasicMCRange(nsteps=5), {"p": np.zeros(4)})
    X = np.ones((10, 17))
    p = K.BasicContMuvParameter("p", logtarget=K.BayesLogit(X, np.ones(10), 1.0))
    with pytest.raises(K.KlaraError) as ei:
        K.BasicMCJob(K.likelihood_model(p, False), K.MALA(0.1), K.BasicMCRange(nsteps=5), {"p": np.zeros(17)})
    assert ei.value.code == K._lib.KLB_EUNSUPPORTED


# ------------------------------------------------------------------ pipelined host-to-host run
@pytest.mark.parametrize("sampler,target,dim", [("HMC", "iso", 1024), ("MALA", "rosen", 96), ("MH", "iso", 7),
                                                ("HMC", "dense", 128), ("HMC", "logit", 4)])
def test_run_host_equals_separate_calls(K, sampler, target, dim):
    """klb_job_run_host (chains sliced over streams, copies overlapped with the kernels) gives exactly what
    set_state + run + output give, for any slice count, including slices that do not divide the chains"""
    L = K._lib
    N = 203
    kw = dict(nchains=N, dim=dim, nsteps=24, burnin=7, thinning=2, step={"HMC": 0.02, "MALA": 0.002, "MH": 0.1}[sampler],
              nleaps=5, seed=314, sigma=np.full(dim, 0.05), tuner="accrate", period=5, target_rate=0.7)
    ref_job, cfg, x0, tp, sg = build_pair(K, sampler, target, **kw)
    ref_job.run()
    ref = ref_job.output()
    for nslices in (1, 3, 16, 0):
        job, *_ = build_pair(K, sampler, target, **kw)
        # start from somewhere else: run_host must replace the state like reset(job, x0)
        job.reset(x0[::-1].copy())
        bufs = {L.OUT_VALUE: np.empty_like(ref.value), L.OUT_LOGTARGET: np.empty_like(ref.logtarget),
                L.OUT_ACCEPT: np.empty_like(ref.diagnosticvalues), L.OUT_STATE: np.empty((N, dim)),
                L.OUT_TUNE_STEP: np.empty(N), L.OUT_TUNE_COUNTERS: np.empty((N, 3), dtype=np.int64)}
        job.run_host(x0, bufs, nslices)
        assert_same("value", bufs[L.OUT_VALUE], ref.value)
        assert_same("logtarget", bufs[L.OUT_LOGTARGET], ref.logtarget)
        assert_same("accept", bufs[L.OUT_ACCEPT], ref.diagnosticvalues)
        assert_same("state", bufs[L.OUT_STATE], ref_job.pstate_value)
        assert_same("step", bufs[L.OUT_TUNE_STEP], ref_job.tune.step)
        assert_same("counters", bufs[L.OUT_TUNE_COUNTERS][:, 2], ref_job.tune.totproposed)
        assert_same("output() after run_host", job.output().value, ref.value)
        assert job.last_run_ms > 0


def test_run_host_continues_and_rejects_bad_starts(K):
    L = K._lib
    job, cfg, x0, tp, sg = build_pair(K, "HMC", "iso", nchains=50, dim=33, nsteps=10, burnin=2, step=0.1, nleaps=3, seed=8)
    two, *_ = build_pair(K, "HMC", "iso", nchains=50, dim=33, nsteps=10, burnin=2, step=0.1, nleaps=3, seed=8)
    job.run_host(x0, {}, 4)
    job.run_host(None, {}, 3)                       # x0 = NULL: reset(job) + run(job) from the current state
    two.run(); two.reset(); two.run()
    assert_same("second run", job.output().value, two.output().value)
    bad = x0.copy()
    bad[37, 5] = np.nan
    with pytest.raises(K.KlaraError) as ei:
        job.run_host(bad, {}, 4)
    assert ei.value.code == L.KLB_ENOTFINITE and "chain 37" in str(ei.value)
    with pytest.raises(K.KlaraError):
        job.run()                                   # no valid state any more
    with pytest.raises(K.KlaraError, match="bytes"):
        two.run_host(x0, {L.OUT_STATE: np.empty((3, 3))}, 2)


@pytest.mark.parametrize("target,dim", [("iso", 1024), ("iso", 130), ("logit", 4)])
def test_run_host_dual_averaging_slices(K, O, target, dim):
    """klb_job_run_host on a DualAveragingMCTuner job cut into several slices: every slice must index ITS chains'
    DualAveragingMCTune records (round-1 bug: slices shared the records of chains 0..nc-1).  Compared bit for bit
    with reset(job, x0) + run(job) and with the oracle started from reset!'s record (step = 1, HMC.jl:217-223)."""
    L = K._lib
    N = 101
    kw = dict(nchains=N, dim=dim, nsteps=36, burnin=10, thinning=2, step=0.6 / np.sqrt(dim) if target == "iso" else 0.02,
              nleaps=12 if target == "iso" else 5, seed=1618, tuner="dualavg", target_rate=0.7, nadapt=25, period=6, verbose=True)
    ref_job, cfg, x0, tp, sg = build_pair(K, "HMC", target, **kw)
    ref_job.reset(x0)                                  # what run_host(x0) does first: reset!(tune) -> step = 1
    ref_job.run()
    ref, ref_tn = ref_job.output(), ref_job.tune
    t_or, d_or = O.da_state(cfg, first=False)
    orc = O.run(cfg, x0, tp, sg, tune=t_or, da=d_or)
    assert_same("oracle value", ref.value, orc["value"])
    for nslices in (3, 16):
        job, *_ = build_pair(K, "HMC", target, **kw)
        bufs = {L.OUT_VALUE: np.empty_like(ref.value), L.OUT_ACCEPT: np.empty_like(ref.diagnosticvalues),
                L.OUT_TUNE_STEP: np.empty(N), L.OUT_TUNE_DA: np.empty((N, 8)), L.OUT_TUNE_COUNTERS: np.empty((N, 3), dtype=np.int64)}
        job.run_host(x0, bufs, nslices)
        assert_same("value", bufs[L.OUT_VALUE], orc["value"])
        assert_same("accept", bufs[L.OUT_ACCEPT], orc["accept"])
        assert_same("step", bufs[L.OUT_TUNE_STEP], orc["tune"]["step"])
        da = bufs[L.OUT_TUNE_DA]
        for i, name in enumerate(("lambda", "mu", "epsbar", "hbar", "hweight", "epsweight", "nleaps", "count")):
            assert_same("da." + name, da[:, i], orc["da"][name])
        assert_same("counters", bufs[L.OUT_TUNE_COUNTERS][:, 2], ref_tn.totproposed)
        assert (da[:, 7] == 36).all() and len(np.unique(da[:, 2])) > N // 2
        tn = job.tune                                  # the records left on the device are the same
        assert_same("tune.epsbar", tn.epsbar, orc["da"]["epsbar"])
    # a bad start must not brick the job (t_global stays 0, so reset / set_state are still the reference's)
    job, *_ = build_pair(K, "HMC", target, **kw)
    bad = x0.copy()
    bad[77, 0] = np.inf
    with pytest.raises(K.KlaraError) as ei:
        job.run_host(bad, {}, 4)
    assert ei.value.code == L.KLB_ENOTFINITE and "chain 77" in str(ei.value)
    assert job.plan().transitions_done == 0
    job.run_host(x0, {}, 5)
    assert_same("after a rejected start", job.output().value, orc["value"])


# ------------------------------------------------------------------ HMC, four consumer warps per chain (dim 1025..4096)
@pytest.mark.parametrize("arith", ["reference", "fma"])
@pytest.mark.parametrize("target,dim,tuner", [("iso", 4096, "accrate"), ("iso", 1025, "vanilla"), ("shifted", 2050, "accrate"),
                                              ("rosen", 3000, "vanilla"), ("iso", 3001, "vanilla"), ("rosen", 4096, "dualavg"), ("shifted", 1536, "dualavg")])
def test_hmc_team_kernel_bit_exact(K, target, dim, tuner, arith):
    """klb_hmc_ws_kernel<W = 4>: one chain per consumer warpgroup, masked and full dims, every tuner, every output
    (value, logtarget, gradient, accept flags), thinning, the verbose rate records      iterate/HMC.jl:124-248"""
    step = {"iso": 0.5 / np.sqrt(dim), "shifted": 0.5 / np.sqrt(dim), "rosen": 0.004}[target]
    job, cfg, x0, tp, sg = build_pair(K, "HMC", target, nchains=11, dim=dim, nsteps=36, burnin=12, thinning=3, step=step,
                                      nleaps=6, seed=8128 + dim, arith=arith, tuner=tuner, target_rate=0.7, nadapt=20,
                                      period=4, verbose=True, monitor=("value", "logtarget", "gradlogtarget"))
    assert job.plan().warps_per_block == 8
    out, ref = compare_run(job, cfg, x0, tp, sg)
    assert 0.05 < out.diagnosticvalues.mean() <= 1.0
    # the same chains through the fused 4-warp-team kernel of klb_kernels.cuh (KLB_HMC_WS=1 keeps only the one-warp
    # warp-specialised geometries): bit-identical by construction
    os.environ["KLB_HMC_WS"] = "1"
    try:
        job2, *_ = build_pair(K, "HMC", target, nchains=11, dim=dim, nsteps=36, burnin=12, thinning=3, step=step,
                              nleaps=6, seed=8128 + dim, arith=arith, tuner=tuner, target_rate=0.7, nadapt=20,
                              period=4, verbose=True, monitor=("value", "logtarget", "gradlogtarget"))
        job2.run()
        out2 = job2.output()
    finally:
        del os.environ["KLB_HMC_WS"]
    assert np.array_equal(out.value, out2.value) and np.array_equal(out.logtarget, out2.logtarget)
    assert np.array_equal(job.pstate_value, job2.pstate_value)


# ------------------------------------------------------------------ DualAveragingMCTuner (HMC)
@pytest.mark.parametrize("arith", ["reference", "fma"])
@pytest.mark.parametrize("target,dim", [("iso", 1024), ("iso", 700), ("iso", 64), ("iso", 7), ("iso", 2048),
                                        ("shifted", 100), ("rosen", 96), ("logit", 4), ("dense", 30), ("dense", 64),
                                        ("dense", 129), ("dense", 2)])
def test_dual_averaging_hmc_bit_exact(K, target, dim, arith):
    """DualAveragingMCTuner (src/tuners/DualAveragingMCTuner.jl:95-101, iterate/HMC.jl:125-127,142-144,225-248):
    per-chain step and per-chain nleaps = max(1, round(λ/step)); adaptation for nadapt transitions, then step = εbar.
    Warp-specialised kernel (dim 1024, 700), fused kernel, 4-warp teams (2048), thread-per-chain (logit), and the
    dense-precision tile kernel, where the chains of a CTA tile stop after their own number of steps."""
    step = {"iso": 0.8 / np.sqrt(dim), "shifted": 0.08, "rosen": 0.01, "logit": 0.02, "dense": 0.05}[target]
    job, cfg, x0, tp, sg = build_pair(K, "HMC", target, nchains=37, dim=dim, nsteps=70, burnin=20, thinning=2, step=step,
                                      nleaps=40 if target == "iso" else 6, seed=2718, arith=arith, tuner="dualavg", target_rate=0.651, nadapt=45,
                                      verbose=(dim % 2 == 0), period=10)
    out, ref = compare_run(job, cfg, x0, tp, sg)
    tn = job.tune
    assert (tn.count == 70).all() and len(np.unique(tn.step)) > 1 and (tn.step == tn.epsbar).all()
    assert 0.2 < out.diagnosticvalues.mean() <= 1.0
    if target == "iso":
        assert len(np.unique(tn.nleaps)) > 1            # chains really run different numbers of leapfrog steps


def test_dual_averaging_chunks_shards_and_reset_rule(K, O):
    L = K._lib
    kw = dict(nchains=24, dim=130, nsteps=40, burnin=10, step=0.05, nleaps=8, seed=99, tuner="dualavg", target_rate=0.8,
              nadapt=25)
    whole, cfg, x0, tp, sg = build_pair(K, "HMC", "iso", **kw)
    whole.run()
    chunked, *_ = build_pair(K, "HMC", "iso", **kw)
    chunked.set_chunk(7)
    chunked.run()
    assert_same("chunked", chunked.output().value, whole.output().value)
    assert_same("chunked step", chunked.tune.step, whole.tune.step)
    shard, *_ = build_pair(K, "HMC", "iso", **dict(kw, nchains=10), chain_offset=5)
    shard.run()
    assert_same("shard", shard.output().value, whole.output().value[5:15])
    # reset!(tune, ::HMC, ::DualAveragingMCTuner) throws in the reference once the job has run
    for call in (lambda: whole.reset(), lambda: whole.reset(x0), lambda: whole.run_host(x0, {}, 2)):
        with pytest.raises(K.KlaraError) as ei:
            call()
        assert ei.value.code == L.KLB_EUNSUPPORTED
    # before the first transition reset gives step = 1, mu = log(10)                    (src/samplers/HMC.jl:217-223)
    fresh, *_ = build_pair(K, "HMC", "iso", **kw)
    assert (fresh.tune.step == 0.05).all() and np.allclose(fresh.tune.mu, np.log(0.5), rtol=1e-15)
    fresh.reset()
    tn = fresh.tune
    assert (tn.step == 1.0).all() and np.allclose(tn.mu, np.log(10.0), rtol=1e-15) and (tn.lam == 8 * 0.05).all()
    t_ref, d_ref = O.da_state(cfg, first=False)
    assert_same("reset mu", tn.mu, d_ref["mu"][:24])
    # dual averaging is an HMC tuner
    with pytest.raises(K.KlaraError):
        build_pair(K, "MALA", "iso", nchains=4, dim=8, nsteps=5, tuner="dualavg")


# ------------------------------------------------------------------ randomized sweep
def _random_configs(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        sampler = rng.choice(["HMC", "MALA", "MH"])
        target = rng.choice(["iso", "shifted", "rosen", "dense", "logit"])
        if target == "logit":
            dim = int(rng.integers(1, 17))
        elif target == "dense":
            dim = int(rng.integers(1, 140))
        else:
            dim = int(rng.choice([rng.integers(1, 64), rng.integers(64, 600), rng.integers(600, 1400), rng.integers(1400, 4097)]))
            if target == "rosen":
                dim += dim & 1
        tuner = rng.choice(["vanilla", "accrate", "dualavg"])
        if tuner == "dualavg" and sampler != "HMC":
            continue
        if tuner == "accrate" and sampler == "MH":
            tuner = "vanilla"
        nsteps = int(rng.integers(3, 40))
        burnin = int(rng.integers(0, nsteps))
        mon = ["value"] + (["logtarget"] if rng.random() < 0.7 else []) + \
              (["gradlogtarget"] if sampler != "MH" and rng.random() < 0.4 else [])
        scale = {"iso": 1.0, "shifted": 1.0, "rosen": 0.15, "dense": 0.5, "logit": 0.15}[target]
        step = {"HMC": 0.6, "MALA": 0.5, "MH": 1.0}[sampler] * scale / max(1.0, dim) ** (0.25 if sampler == "HMC" else 1 / 3)
        out.append(dict(sampler=str(sampler), target=str(target), dim=dim, nchains=int(rng.integers(1, 70)), nsteps=nsteps,
                        burnin=burnin, thinning=int(rng.integers(1, 4)), step=float(step), nleaps=int(rng.integers(1, 9)),
                        tuner=str(tuner), target_rate=float(rng.uniform(0.3, 0.9)), period=int(rng.integers(2, 12)),
                        verbose=bool(rng.random() < 0.4), monitor=tuple(mon),
                        diagnostics=("accept",) if rng.random() < 0.8 else (), seed=int(rng.integers(0, 2**62)),
                        arith=str(rng.choice(["reference", "fma"])), chain_offset=int(rng.integers(0, 1000)),
                        nadapt=int(rng.integers(1, 50)), sigma=np.full(dim, 0.3 * scale / max(1.0, dim) ** 0.5)))
    return out


@pytest.mark.parametrize("k,cfgd", list(enumerate(_random_configs(40, 20261017))))
def test_random_configurations_bit_exact(K, k, cfgd):
    """40 seeded random draws over sampler x target x dim x tuner x range x arithmetic x monitor x shard offset:
    every field, the final state and the tuner records against the oracle, bit for bit"""
    job, cfg, x0, tp, sg = build_pair(K, rng_seed=k, **cfgd)
    compare_run(job, cfg, x0, tp, sg)


def test_device_erf_and_shared_divisor_division(K, O):
    """klb_erf on the device == the oracle's (same double-double series); DivBy (the MALA kernels' division by the drift
    step with the reciprocal refinement hoisted) == IEEE division, bit for bit, on random and on extreme operands"""
    rng = np.random.default_rng(9)
    xs = np.concatenate([rng.uniform(-6.5, 6.5, 20000), np.array([0.0, -0.0, 4.5, 4.500001, 6.0, -6.0, 7.0, np.inf, -np.inf, np.nan, 1e-300])])
    out = np.empty_like(xs)
    K._lib.check(K._lib.lib().klb_debug_math(0, 2, xs.size, _ptr(xs), _ptr(out)))
    assert_same("erf", out, np.array([O.erf(x) for x in xs]))
    n = 1 << 20
    a = np.concatenate([rng.uniform(0, 4, n) ** 2, 10.0 ** rng.uniform(-320, 308, n), rng.normal(size=n)])
    b = np.concatenate([10.0 ** rng.uniform(-4, 1, n), 10.0 ** rng.uniform(-320, 308, n), rng.uniform(1e-3, 2, n)])
    edge = np.array([0.0, -0.0, 5e-324, 2.2250738585072014e-308, 1e-292, 1e-291, 1.7976931348623157e308, np.inf, -np.inf, np.nan, 1.0, 3.0])
    ea, eb = np.meshgrid(edge, edge)
    a, b = np.concatenate([a, ea.ravel()]), np.concatenate([b, eb.ravel()])
    # significands of all ones / divisors just below a power of two: the classic hard cases of reciprocal-based division
    hard = np.nextafter(2.0 ** rng.integers(-30, 30, 4096).astype(np.float64), 0)
    a, b = np.concatenate([a, rng.uniform(0.5, 2, 4096), hard]), np.concatenate([b, hard, rng.uniform(0.5, 2, 4096)])
    pairs = np.empty(2 * a.size)
    pairs[0::2], pairs[1::2] = a, b
    out = np.empty_like(pairs)
    K._lib.check(K._lib.lib().klb_debug_math(0, 3, pairs.size, _ptr(pairs), _ptr(out)))
    with np.errstate(all="ignore"):
        want = a / b
    assert_same("shared-divisor division", out[0::2], want)
    assert_same("shared-divisor division (second slot)", out[1::2], want)


def test_verbose_tuner_records_and_prints_the_burnin_rates(K, O, capsys):
    """verbose tuners: the per-period burn-in acceptance rates the reference prints (iterate/HMC.jl:211-221,
    iterate/MH.jl:126-139) are recorded by the kernels (KLB_OUT_TUNE_RATES) and printed by run() in the reference's
    format; rate of period k = accepted / period of transitions k*period+1 .. (k+1)*period, from the accept flags"""
    for sampler, tuner, kw in (("HMC", "accrate", dict(step=0.08, nleaps=4)), ("MALA", "vanilla", dict(step=0.2)),
                               ("MH", "vanilla", dict(sigma=np.full(24, 0.3)))):
        job, cfg, x0, tp, sg = build_pair(K, sampler, "iso", nchains=7, dim=24, nsteps=70, burnin=0, seed=3, tuner=tuner,
                                          period=10, verbose=True, **kw)
        # burnin = 0 saves every transition: re-run the same chains with a burn-in of 45 and compare the records
        job.run()
        acc = job.output().diagnosticvalues
        capsys.readouterr()
        jb, cfgb, *_ = build_pair(K, sampler, "iso", nchains=7, dim=24, nsteps=70, burnin=45, seed=3, tuner=tuner,
                                  period=10, verbose=True, **kw)
        jb.run()
        text = capsys.readouterr().out.strip().splitlines()
        rates = jb.burnin_rates
        assert rates.shape == (7, 4) and len(text) == 4
        if tuner == "vanilla":          # no adaptation: the burn-in chain equals the head of the all-saved chain
            want = acc[:, :40].reshape(7, 4, 10).mean(-1)
            assert_same("burn-in rates", rates, want)
        assert text[0].startswith("Burnin iteration 10 of 45: ") and text[0].split(": ")[1].startswith("%6.2f" % (100 * rates[:, 0].mean()))
        assert " % acceptance rate" in text[3] and text[3].startswith("Burnin iteration 40 of 45")
    single, *_ = build_pair(K, "MH", "iso", nchains=1, dim=2, nsteps=30, burnin=20, seed=4, period=10, verbose=True,
                            sigma=np.ones(2), x0=np.array([[5.1, -0.9]]))
    single.run()
    out = capsys.readouterr().out.strip().splitlines()
    r = single.burnin_rates[0]
    assert out == ["Burnin iteration %2d of 20: %6.2f %% acceptance rate" % (10 * (k + 1), 100 * r[k]) for k in range(2)]


# ------------------------------------------------------------------ NUTS
@pytest.mark.parametrize("arith", ["reference", "fma"])
@pytest.mark.parametrize("target,dim,tuner,maxnd,maxdelta", [
    ("iso", 7, "vanilla", 5, 1000), ("iso", 64, "dualavg", 4, 1000), ("shifted", 100, "vanilla", 6, 1000),
    ("rosen", 96, "dualavg", 5, 1000), ("iso", 250, "vanilla", 3, 1000), ("shifted", 512, "dualavg", 4, 1000),
    ("iso", 1024, "vanilla", 4, 1000), ("rosen", 1000, "vanilla", 3, 1000), ("iso", 1500, "dualavg", 3, 1000),
    ("shifted", 4096, "vanilla", 2, 1000), ("iso", 33, "vanilla", 7, 2), ("iso", 2, "dualavg", 10, 1)])
def test_nuts_bit_exact(K, target, dim, tuner, maxnd, maxdelta, arith):
    """NUTS as the reference computes it (iterate/NUTS.jl:230-457, NUTS.jl:514-628, :781-927; the resolved aliasing of
    DESIGN.md 6b): every geometry, both tuners, trees that stop early (small maxδ, large steps), both diagnostics"""
    step = {"iso": 0.9 / np.sqrt(dim) ** 0.5, "shifted": 0.9 / np.sqrt(dim) ** 0.5, "rosen": 0.01}[target]
    if maxdelta < 10:
        step *= 3
    job, cfg, x0, tp, sg = build_pair(K, "NUTS", target, nchains=19, dim=dim, nsteps=40, burnin=12, thinning=2, step=step,
                                      seed=31337 + dim, arith=arith, tuner=tuner, target_rate=0.7, nadapt=25, period=5,
                                      verbose=(dim % 2 == 0), monitor=("value", "logtarget", "gradlogtarget"),
                                      diagnostics=("accept", "ndoublings") + (("a", "na") if tuner == "dualavg" else ()),
                                      maxdelta=maxdelta, maxndoublings=maxnd)
    out, ref = compare_run(job, cfg, x0, tp, sg)
    nd = ref["ndoublings"]
    assert nd.min() >= 1 and nd.max() <= maxnd
    if maxdelta < 10:
        assert nd.min() < maxnd                        # some trees stopped before the last doubling
    if tuner == "dualavg":                             # :a, :na: the leaves of the last doubling (iterate/NUTS.jl:393-399)
        assert out.diagnostickeys == ["accept", "ndoublings", "a", "na"] and out.diagnosticvalues.dtype == np.float64
        assert (ref["na"] >= 1).all() and (ref["na"] <= 2 ** (nd.astype(np.int64) - 1)).all()


@pytest.mark.parametrize("dim,tuner,maxnd,maxdelta,step,arith", [
    (4, "dualavg", 5, 1000, 0.15, "reference"), (4, "vanilla", 4, 1000, 0.2, "fma"), (3, "vanilla", 6, 1000, 0.1, "reference"),
    (2, "dualavg", 3, 1000, 0.3, "fma"), (7, "dualavg", 4, 1000, 0.1, "reference"), (8, "vanilla", 5, 3, 0.3, "reference"),
    (16, "vanilla", 3, 1000, 0.08, "fma"), (13, "dualavg", 4, 1000, 0.08, "reference")])
def test_nuts_logit_bit_exact(K, dim, tuner, maxnd, maxdelta, step, arith):
    """NUTS on the Bayesian logistic-regression target (doc/examples/swiss/NUTS/{noadaptation,dualaveraging}/analytical.jl):
    klb_glm_kernel<3, DP, FMA>, one thread per chain, every padded dimension, both tuners, trees that stop early"""
    job, cfg, x0, tp, sg = build_pair(K, "NUTS", "logit", nchains=70, dim=dim, nsteps=30, burnin=9, thinning=2, step=step,
                                      seed=5150 + dim, arith=arith, tuner=tuner, target_rate=0.651, nadapt=20, period=5,
                                      verbose=(dim % 2 == 0), monitor=("value", "logtarget", "gradlogtarget"),
                                      diagnostics=("accept", "ndoublings") + (("na", "a") if tuner == "dualavg" else ()),
                                      maxdelta=maxdelta, maxndoublings=maxnd)
    out, ref = compare_run(job, cfg, x0, tp, sg)
    nd = ref["ndoublings"]
    assert nd.min() >= 1 and nd.max() <= maxnd
    if maxdelta < 10:
        assert nd.min() < nd.max()                     # some trees stopped before the last doubling


def test_nuts_chunks_shards_and_run_host(K, O):
    """one launch per transition == one launch; 2 shards == 1 job; the pipelined host call == the three calls"""
    kw = dict(nchains=26, dim=130, nsteps=30, burnin=8, step=0.25, seed=77, tuner="dualavg", target_rate=0.65, nadapt=20,
              period=4, verbose=True, diagnostics=("accept", "ndoublings", "a", "na"), maxndoublings=4)
    whole, cfg, x0, tp, sg = build_pair(K, "NUTS", "iso", **kw)
    out, ref = compare_run(whole, cfg, x0, tp, sg)
    step, *_ = build_pair(K, "NUTS", "iso", **kw)
    step.set_chunk(1)
    step.run()
    assert_same("chunked value", step.output().value, out.value)
    assert_same("chunked diagnostics", step.output().diagnosticvalues, out.diagnosticvalues)
    a, *_ = build_pair(K, "NUTS", "iso", **dict(kw, nchains=10, x0=x0[:10]))
    b, *_ = build_pair(K, "NUTS", "iso", **dict(kw, nchains=16, x0=x0[10:], chain_offset=10))
    a.run(); b.run()
    assert_same("sharded value", np.concatenate([a.output().value, b.output().value]), out.value)
    assert_same("sharded steps", np.concatenate([a.tune.step, b.tune.step]), whole.tune.step)
    assert_same("sharded diagnostics", np.concatenate([a.output().diagnosticvalues, b.output().diagnosticvalues]), out.diagnosticvalues)
    # run_host = reset(job, x0) + run + output; with the vanilla tuner (a reset dual-averaging job restarts from step = 1,
    # NUTS.jl:320-325, so it would not retrace the constructor's run)
    kv = dict(kw, tuner="vanilla", diagnostics=("accept", "ndoublings"))
    plain, cfgv, _, _, _ = build_pair(K, "NUTS", "iso", **kv)
    outv, refv = compare_run(plain, cfgv, x0, tp, sg)
    host, *_ = build_pair(K, "NUTS", "iso", **kv)
    val = np.empty_like(outv.value)
    nd = np.empty((26, outv.value.shape[1]), dtype=np.uint8)
    host.run_host(x0, {K._lib.OUT_VALUE: val, K._lib.OUT_NDOUBLINGS: nd}, 3)
    assert_same("run_host value", val, outv.value)
    assert_same("run_host ndoublings", nd, refv["ndoublings"])
    # :a / :na through the sliced pipeline: one slice == three slices (fresh dual-averaging jobs; both restart the tuner alike)
    got = []
    for nslices in (1, 3):
        hj, *_ = build_pair(K, "NUTS", "iso", **kw)
        bufs = {K._lib.OUT_NUTS_A: np.empty((26, val.shape[1])), K._lib.OUT_NUTS_NA: np.empty((26, val.shape[1]), dtype=np.int32),
                K._lib.OUT_VALUE: np.empty_like(val)}
        hj.run_host(x0, bufs, nslices)
        got.append(bufs)
        hj.close()
    for f in got[0]:
        assert_same("run_host field %d, 1 vs 3 slices" % f, got[1][f], got[0][f])
    assert (got[0][K._lib.OUT_NUTS_NA] >= 1).all()


def test_nuts_validation(K):
    L = K._lib
    with pytest.raises(AssertionError, match="maxδ is not positive"):                         # NUTS.jl:235
        K.NUTS(0.1, maxdelta=0)
    with pytest.raises(K.KlaraError) as ei:                                                    # no sampler_state method
        build_pair(K, "NUTS", "iso", nchains=3, dim=8, nsteps=5, tuner="accrate")
    assert ei.value.code == L.KLB_EINVAL
    with pytest.raises(K.KlaraError) as ei:
        build_pair(K, "NUTS", "dense", nchains=3, dim=8, nsteps=5)
    assert ei.value.code == L.KLB_EUNSUPPORTED
    with pytest.raises(K.KlaraError) as ei:
        build_pair(K, "NUTS", "iso", nchains=3, dim=8, nsteps=5, maxndoublings=11)
    assert ei.value.code == L.KLB_EUNSUPPORTED
    with pytest.raises(KeyError):
        build_pair(K, "HMC", "iso", nchains=3, dim=8, nsteps=5, diagnostics=("ndoublings",))
    with pytest.raises(KeyError):                                                              # :a, :na need dual averaging (NUTS.jl:317)
        build_pair(K, "NUTS", "iso", nchains=3, dim=8, nsteps=5, diagnostics=("accept", "a"))
