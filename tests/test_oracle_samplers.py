"""CPU tests of the oracle's sampler semantics (the edge cases of SURVEY.md section 3.6) and of its
statistical sanity against theory: the target exp(-z.z) is N(0, I/2)."""
import numpy as np
import pytest


def _run(O, sampler, target=None, **kw):
    target = O.ISO if target is None else target
    n, d, nsteps = kw.pop("nchains", 8), kw.pop("dim", 4), kw.pop("nsteps", 100)
    sigma = kw.pop("sigma", None)
    x0 = kw.pop("x0", None)
    tparams = kw.pop("tparams", None)
    cfg = O.make_config(sampler, target, n, d, nsteps, nthreads=O.max_threads(), **kw)
    if x0 is None:
        x0 = np.stack([O.normals(cfg.seed, c, 0, d) for c in range(n)])
    return cfg, O.run(cfg, x0, tparams, sigma)


def test_tuner_window_hmc(O):
    """tune events after transitions period, 2*period, ... while totproposed <= burnin; totproposed starts at
    `period` and ends at period*(1 + floor(burnin/period)); counters keep accumulating after burn-in"""
    cfg, r = _run(O, O.HMC, nsteps=1300, burnin=1000, step=0.05, nleaps=3, tuner=O.ACCRATE, period=100,
                  target_rate=0.8, monitor=0, diagnostics=1, seed=5)
    t = r["tune"]
    assert (t["totproposed"] == 1100).all()
    assert (t["proposed"] == 300).all()
    assert (t["accepted"] <= 300).all() and (t["accepted"] == r["accept"].sum(1)).all()
    assert np.isnan(t["rate"]).all()                      # reset_burnin! leaves NaN, never recomputed after burn-in
    assert (t["step"] != 0.05).all()


def test_counters_only_with_accrate_or_verbose(O):
    for sampler in (O.HMC, O.MALA):
        cfg, r = _run(O, sampler, nsteps=50, burnin=20, step=0.1, nleaps=2, tuner=O.VANILLA, verbose=0, seed=1)
        assert (r["tune"]["proposed"] == 0).all() and (r["tune"]["accepted"] == 0).all()
        cfg, r = _run(O, sampler, nsteps=50, burnin=20, step=0.1, nleaps=2, tuner=O.VANILLA, verbose=1, period=10, seed=1)
        assert (r["tune"]["totproposed"] == 30).all() and (r["tune"]["proposed"] == 30).all()
    # MH: counters only when verbose; AcceptanceRateMCTuner never adapts MH (iterate/MH.jl has no tune! call)
    cfg, r = _run(O, O.MH, nsteps=60, burnin=30, tuner=O.ACCRATE, verbose=0, period=10, sigma=np.full(4, 0.5), seed=1)
    assert (r["tune"]["proposed"] == 0).all() and (r["tune"]["step"] == 1.0).all()
    cfg, r = _run(O, O.MH, nsteps=60, burnin=30, tuner=O.ACCRATE, verbose=1, period=10, sigma=np.full(4, 0.5), seed=1)
    assert (r["tune"]["totproposed"] == 40).all() and (r["tune"]["step"] == 1.0).all()


def test_nan_and_divergent_proposals_reject(O):
    cfg, r = _run(O, O.HMC, nsteps=10, step=1e200, nleaps=3, monitor=1, diagnostics=1, seed=2)
    assert r["accept"].sum() == 0
    x0 = np.stack([O.normals(2, c, 0, 4) for c in range(8)])
    np.testing.assert_array_equal(r["x"], x0)
    cfg, r = _run(O, O.MALA, nsteps=10, step=1e300, monitor=1, diagnostics=1, seed=2)
    assert r["accept"].sum() == 0


def test_nonfinite_start_is_rejected(O):
    x0 = np.zeros((3, 4)); x0[1, 2] = np.nan
    with pytest.raises(ValueError, match="chain 1"):
        _run(O, O.HMC, nchains=3, x0=x0, nsteps=3)


def test_saved_logtarget_matches_value(O):
    for sampler, kw in [(O.HMC, dict(step=0.1, nleaps=4)), (O.MALA, dict(step=0.3)), (O.MH, dict(sigma=np.full(6, 0.4)))]:
        cfg, r = _run(O, sampler, dim=6, nsteps=80, burnin=13, thinning=3, monitor=3, diagnostics=1, seed=9, **kw)
        np.testing.assert_allclose(r["logtarget"], -(r["value"] ** 2).sum(-1), rtol=1e-13)
        rej = r["accept"][:, 1:] == 0
        if cfg.thinning == 1:
            assert np.array_equal(r["value"][:, 1:][rej], r["value"][:, :-1][rej])


def test_chain_split_invariance(O):
    """N chains in one call == the same chains run in two calls with chain_offset (map(run, jobs) semantics)"""
    cfg, full = _run(O, O.MALA, nchains=10, dim=5, nsteps=40, step=0.2, monitor=1, seed=4)
    x0 = np.stack([O.normals(4, c, 0, 5) for c in range(10)])
    cfg2 = O.make_config(O.MALA, O.ISO, 6, 5, 40, step=0.2, monitor=1, seed=4, chain_offset=4)
    part = O.run(cfg2, x0[4:])
    assert np.array_equal(part["value"], full["value"][4:])


def test_continuation_equals_one_long_run(O):
    """two runs of 30 transitions (second starting at t0 = 30 from the first's state, tuner record carried)
    produce the same final state as one run of 60: the chain state persists between runs"""
    kw = dict(step=0.1, nleaps=3, monitor=0, diagnostics=0, seed=6, tuner=O.ACCRATE, period=7, burnin=0)
    x0 = np.stack([O.normals(6, c, 0, 4) for c in range(5)])
    cfg = O.make_config(O.HMC, O.ISO, 5, 4, 60, **kw)
    one = O.run(cfg, x0)
    cfg_a = O.make_config(O.HMC, O.ISO, 5, 4, 30, **kw)
    a = O.run(cfg_a, x0)
    cfg_b = O.make_config(O.HMC, O.ISO, 5, 4, 30, t0=30, **kw)
    b = O.run(cfg_b, a["x"], tune=a["tune"], logtarget=a["logtarget_state"])
    assert np.array_equal(one["x"], b["x"])


@pytest.mark.parametrize("sampler,kw,lo,hi", [
    ("HMC", dict(step=0.15, nleaps=8), 0.7, 1.0),
    ("MALA", dict(step=0.25), 0.4, 0.95),
    ("MH", dict(sigma=np.full(8, 0.35)), 0.15, 0.6),
])
def test_moments_of_the_isotropic_target(O, sampler, kw, lo, hi):
    """exp(-z.z) = N(0, I/2): sample mean -> 0, variance -> 0.5; acceptance in a plausible band"""
    code = {"HMC": O.HMC, "MALA": O.MALA, "MH": O.MH}[sampler]
    cfg, r = _run(O, code, nchains=64, dim=8, nsteps=3000, burnin=500, monitor=1, diagnostics=1, seed=77, **kw)
    v = r["value"]
    assert abs(v.mean()) < 0.02
    assert abs(v.var() - 0.5) < 0.02
    assert lo < r["accept"].mean() < hi


def test_rosenbrock_and_shifted_targets(O, K):
    """oracle targets == the Python descriptors (closed forms) to rounding"""
    rng = np.random.default_rng(0)
    x = rng.normal(size=10)
    mu = rng.normal(size=10)
    lt, g = O.eval_target(O.make_config(O.HMC, O.SHIFTED, 1, 10, 1), x, mu)
    t = K.ShiftedIsoGaussian(mu)
    assert lt == pytest.approx(t(x), rel=1e-13)
    np.testing.assert_allclose(g, t.gradient(x), rtol=1e-13)
    lt, g = O.eval_target(O.make_config(O.HMC, O.ROSEN, 1, 10, 1), x, np.array([1.0, 100.0, 0.05]))
    t = K.Rosenbrock()
    assert lt == pytest.approx(t(x), rel=1e-12)
    np.testing.assert_allclose(g, t.gradient(x), rtol=1e-12)
    # numerical gradient check of the Rosenbrock descriptor
    eps = 1e-6
    num = np.array([(t(x + eps * np.eye(10)[i]) - t(x - eps * np.eye(10)[i])) / (2 * eps) for i in range(10)])
    np.testing.assert_allclose(num, g, rtol=1e-5, atol=1e-6)


def test_reduction_order_is_the_documented_one(O):
    """orc_dot reproduces the canonical order of DESIGN.md literally (exact-arithmetic restatement: every term
    is accumulated by one fma, i.e. acc <- round(acc + a*b) with a single rounding)"""
    from fractions import Fraction

    def fma(a, b, c):
        return float(Fraction(a) * Fraction(b) + Fraction(c))

    rng = np.random.default_rng(1)
    for d, nv in [(5, 1), (64, 1), (100, 2), (1024, 16), (777, 16)]:
        a, b = rng.normal(size=d), rng.normal(size=d)
        pa = np.zeros(64 * nv); pb = np.zeros(64 * nv); pa[:d] = a; pb[:d] = b
        lanes = np.zeros(32)
        for l in range(32):
            acc = [0.0] * 4
            for m in range(nv):
                k = l + 32 * m
                acc[m & 3] = fma(pa[2 * k], pb[2 * k], acc[m & 3])
                acc[m & 3] = fma(pa[2 * k + 1], pb[2 * k + 1], acc[m & 3])
            lanes[l] = (acc[0] + acc[1]) + (acc[2] + acc[3])
        for s in (16, 8, 4, 2, 1):
            lanes = np.array([lanes[l] + lanes[l ^ s] for l in range(32)])
        assert O.dot(a, b, nv=nv) == lanes[0]


def _ess_numpy(v):
    """literal numpy restatement of mcvar(v, Val{:imse}) / mcvar(v, Val{:iid}) / ess
    (src/stats/variance/mcvar.jl:5,75-105, src/stats/convergence/ess.jl:3-5; StatsBase.autocov divides by n)"""
    n = len(v)
    z = v - v.mean()
    acv = np.array([np.dot(z[:n - k], z[k:]) / n for k in range(n)])
    maxlag = n - 1
    k = int(np.floor((maxlag - 1) / 2))
    m = k + 1
    g = np.zeros(k + 1)
    for j in range(k + 1):
        g[j] = acv[2 * j] + acv[2 * j + 1]
        if g[j] <= 0:
            m = j
            break
    for j in range(1, m):
        if g[j] > g[j - 1]:
            g[j] = g[j - 1]
    mcvar = (-acv[0] + 2 * g[:m].sum()) / n
    return n * (v.var(ddof=1) / n) / mcvar


def test_ess_follows_the_reference_estimator(O):
    rng = np.random.default_rng(3)
    n = 500
    ar = np.zeros((3, n, 4))
    e = rng.normal(size=(3, n, 4))
    for t in range(1, n):
        ar[:, t] = 0.6 * ar[:, t - 1] + e[:, t]
    got = O.ess(ar)
    for c in range(3):
        for i in range(4):
            assert got[c, i] == pytest.approx(_ess_numpy(ar[c, :, i]), rel=1e-10)
    # theory: AR(1) with rho = 0.6 has ESS ~ n (1-rho)/(1+rho) = n/4; iid noise has ESS ~ n
    assert 0.5 * n / 4 < got.mean() < 2 * n / 4
    assert 0.7 * n < O.ess(rng.normal(size=(4, n, 4))).mean() < 1.4 * n
    assert np.isnan(O.ess(rng.normal(size=(1, 3, 2)))).all()          # fewer than 4 samples: undefined


def test_stats_follow_the_reference_estimators(O):
    """mean (src/stats/mean.jl:9), mcvar(:iid) = var/len (mcvar.jl:5), mcvar(:imse) (mcvar.jl:75-105),
    ess = len*iid/imse (ess.jl:3), iact = imse/iid (iact.jl:3), acceptance (acceptance.jl:1-14) against literal
    numpy restatements"""
    rng = np.random.default_rng(11)
    n = 300
    ar = np.zeros((2, n, 3))
    e = rng.normal(size=(2, n, 3))
    for t in range(1, n):
        ar[:, t] = 0.5 * ar[:, t - 1] + e[:, t]
    st = O.stats(ar)
    for c in range(2):
        for i in range(3):
            v = ar[c, :, i]
            iid = v.var(ddof=1) / n
            ess = _ess_numpy(v)
            assert st["mean"][c, i] == pytest.approx(v.mean(), rel=1e-12, abs=1e-15)
            assert st["mcvar_iid"][c, i] == pytest.approx(iid, rel=1e-11)
            assert st["ess"][c, i] == pytest.approx(ess, rel=1e-10)
            assert st["mcvar_imse"][c, i] == pytest.approx(n * iid / ess, rel=1e-10)
            assert st["iact"][c, i] == pytest.approx(n / ess, rel=1e-10)
    assert np.array_equal(st["ess"], O.ess(ar))
    short = O.stats(rng.normal(size=(1, 3, 2)))
    assert np.isfinite(short["mean"]).all() and all(np.isnan(short[k]).all() for k in O.STAT_NAMES[1:])
    # acceptance(v::AbstractArray{Bool}) = mean(v); acceptance(values) counts the changes, the first sample included
    acc = rng.random((4, 50)) < 0.7
    assert np.array_equal(O.acceptance(accept=acc), acc.mean(axis=1))
    vals = np.cumsum(acc[:, :, None] * rng.normal(size=(4, 50, 3)), axis=1)
    expect = [(1 + sum((vals[c, t] != vals[c, t - 1]).any() for t in range(1, 50))) / 50 for c in range(4)]
    assert np.array_equal(O.acceptance(value=vals), np.array(expect))


def test_gibbs_oracle_is_run_plus_reset_per_sweep(O):
    """oracle/gibbs.py (BasicGibbsJob.jl:185-231): with a tuner that never adapts, S sweeps of an n-step dpjob are one
    S*n-step chain (state and RNG counter persist across reset(dpjob)), saved once per post-burn-in sweep"""
    from oracle import gibbs as OG
    cfg = O.make_config(O.HMC, O.ISO, 4, 10, 3, step=0.1, nleaps=3, monitor=1, seed=5)
    x0 = np.stack([O.normals(5, c, 0, 10) for c in range(4)])
    out = OG.run_gibbs({"a": dict(cfg=cfg, x0=x0)}, {"twice": lambda v: 2 * v["a"]}, ["a", "twice"], 6, 2, 2)
    long = O.run(O.make_config(O.HMC, O.ISO, 4, 10, 18, step=0.1, nleaps=3, monitor=1, seed=5), x0)
    assert np.array_equal(out["a"][:, 0], long["value"][:, 8]) and np.array_equal(out["a"][:, 1], long["value"][:, 14])
    assert np.array_equal(out["twice"], 2 * out["a"])
    # with the AcceptanceRate tuner the record restarts every sweep: the step never drifts beyond one sweep's tuning
    cfg = O.make_config(O.MALA, O.ISO, 4, 10, 4, burnin=4 - 1, step=0.5, tuner=O.ACCRATE, period=2, target_rate=0.5, monitor=1, seed=6)
    out2 = OG.run_gibbs({"a": dict(cfg=cfg, x0=x0)}, {}, ["a"], 5, 0, 1)
    assert out2["a"].shape == (4, 5, 10)


def test_tuned_samplers_hit_their_target_rate_and_the_posterior(O):
    """semantic check of the tuner branches that no reference KAT pins: after burn-in the acceptance rate
    sits near the tuner's target and the chain samples exp(-z.z) = N(0, I/2) -- an error shared by the oracle and the
    kernels (which agree bit for bit) would show up here"""
    d, n = 10, 64
    x0 = np.stack([O.normals(3, c, 0, d) for c in range(n)]) * 0.7
    # MALA + AcceptanceRateMCTuner(0.574), both score functions
    for score in (0, 1):
        cfg = O.make_config(O.MALA, O.ISO, n, d, 3000, burnin=2000, step=0.05, tuner=O.ACCRATE, target_rate=0.574,
                            score_k=3.0 if score else 7.0, period=100, monitor=1, diagnostics=1, seed=21, nthreads=O.max_threads(),
                            score=score)
        r = O.run(cfg, x0)
        assert abs(r["accept"].mean() - 0.574) < 0.08, r["accept"].mean()
        v = r["value"].reshape(-1, d)
        assert abs(v.mean()) < 0.03 and abs(v.var() - 0.5) < 0.05
        assert (r["tune"]["step"] != 0.05).all() and (r["tune"]["totproposed"] == 2100).all()
    # HMC + AcceptanceRateMCTuner(0.8)
    cfg = O.make_config(O.HMC, O.ISO, n, d, 1500, burnin=1000, step=0.6, nleaps=5, tuner=O.ACCRATE, target_rate=0.8,
                        period=50, monitor=1, diagnostics=1, seed=22, nthreads=O.max_threads())
    r = O.run(cfg, x0)
    assert abs(r["accept"].mean() - 0.8) < 0.1, r["accept"].mean()
    assert abs(r["value"].reshape(-1, d).var() - 0.5) < 0.05
    # HMC + DualAveragingMCTuner(0.65): adapts for nadapt transitions, then freezes the averaged step
    cfg = O.make_config(O.HMC, O.ISO, n, d, 1500, burnin=1000, step=0.1, nleaps=10, tuner=O.DUALAVG, target_rate=0.65,
                        monitor=1, diagnostics=1, seed=23, nthreads=O.max_threads(), nadapt=1000)
    r = O.run(cfg, x0)
    assert abs(r["accept"].mean() - 0.65) < 0.12, r["accept"].mean()
    assert abs(r["value"].reshape(-1, d).var() - 0.5) < 0.06
    assert (r["tune"]["step"] == r["da"]["epsbar"]).all() and (r["da"]["count"] == 1500).all()
