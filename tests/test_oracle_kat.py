"""CPU tests: the oracle against every known-answer value the reference's own tests hold for this path
 and against the published vectors of the primitives it is built on."""
import math

import numpy as np
import pytest


# ---- reference KATs (bit-exact) ------------------------------------------------------------
def test_logistic_kat(O):
    # test/common.jl:6
    assert O.lib().orc_logistic(0.7, 3., 4., 2.1, 1.4) == 1.4110527196983078


def test_logistic_rate_score_kat(O):
    # test/AcceptanceRateMCTuner.jl:8-9
    assert O.lib().orc_logistic_rate_score(0.25, 7.) == 1.7039056039366212
    assert O.lib().orc_logistic_rate_score(0.5, 11.) == 1.991859724568208


def test_erf_rate_score_kat(O):
    # test/AcceptanceRateMCTuner.jl:13-14
    assert O.lib().orc_erf_rate_score(-0.1, 3.) == 0.6713732405408726
    assert O.lib().orc_erf_rate_score(0.93, 2.) == 1.9914724883356396


def test_python_mirror_scores_match(K):
    assert K.logistic(0.7, 3, 4, 2.1, 1.4) == 1.4110527196983078
    assert K.logistic_rate_score(0.25) == 1.7039056039366212
    assert K.logistic_rate_score(0.5, 11) == 1.991859724568208


def test_function_defined_normal_target(O, K):
    """test/BasicContMuvParameter.jl:539-563: logtarget = -(x-mu).(x-mu), gradlogtarget = -2(x-mu) at
    pv = [-4.29, 2.91], mu = [2.2, 2.02]; 0.5*(lt - d*log(2pi)) == logpdf(MvNormal(mu, 1)), 0.5*glt == gradlogpdf"""
    pv, mu = np.array([-4.29, 2.91]), np.array([2.2, 2.02])
    cfg = O.make_config(O.HMC, O.SHIFTED, 1, 2, 1)
    lt, g = O.eval_target(cfg, pv, mu)
    assert lt == pytest.approx(-42.9122, rel=1e-14)
    np.testing.assert_allclose(g, [12.98, -1.78], rtol=1e-14)
    logpdf = -0.5 * (2 * math.log(2 * math.pi) + float((pv - mu) @ (pv - mu)))
    assert 0.5 * (lt - 2 * math.log(2 * math.pi)) == pytest.approx(logpdf, rel=1e-14)
    np.testing.assert_allclose(0.5 * g, -(pv - mu), rtol=1e-14)
    # the descriptor evaluated on the host agrees
    t = K.ShiftedIsoGaussian(mu)
    assert t(pv) == pytest.approx(lt, rel=1e-14)
    np.testing.assert_allclose(t.gradient(pv), g, rtol=1e-14)


def test_pdf_defined_targets(O):
    """test/BasicContMuvParameter.jl:39-80: MvNormal(mu, 1.) at the two test points, via the closed form
    logpdf = -0.5 (d log 2pi + |x-mu|^2); the unnormalised device target is 2*logpdf + d*log(2pi)"""
    for pv, mu in [([5.18, -7.76], [6.11, -8.5]), ([-11.87, -13.44], [-20.2, -18.91])]:
        pv, mu = np.array(pv), np.array(mu)
        lt, g = O.eval_target(O.make_config(O.HMC, O.SHIFTED, 1, 2, 1), pv, mu)
        logpdf = -0.5 * (2 * math.log(2 * math.pi) + float((pv - mu) @ (pv - mu)))
        assert 0.5 * (lt - 2 * math.log(2 * math.pi)) == pytest.approx(logpdf, rel=1e-13)
        np.testing.assert_allclose(0.5 * g, -(pv - mu), rtol=1e-13)


def test_nstate_column_layout(O):
    """test/ParameterNStates.jl:137-146: copy!(nstate, state, i) puts the state in column i of `value`
    (size x n) and entry i of `logtarget`: sample s of chain c sits at value[c, s, :]"""
    cfg = O.make_config(O.MH, O.ISO, 3, 4, 7, burnin=2, thinning=2, monitor=3, diagnostics=1, seed=3)
    x0 = np.arange(12, dtype=float).reshape(3, 4) / 10
    r = O.run(cfg, x0, sigma=np.full(4, 0.3))
    assert r["npost"] == 3 and r["value"].shape == (3, 3, 4) and r["logtarget"].shape == (3, 3)
    np.testing.assert_allclose(r["logtarget"], -(r["value"] ** 2).sum(-1), rtol=1e-15)
    np.testing.assert_array_equal(r["value"][:, -1], r["x"])       # last saved column = final pstate (nsteps in postrange)


def test_range_npoststeps(O, K):
    # src/ranges/BasicMCRange.jl:14-25: length((burnin+1):thinning:nsteps)
    for b, t, n in [(0, 1, 100), (1000, 1, 10000), (10, 3, 100), (5, 7, 6), (99, 100, 100)]:
        assert O.npoststeps(b, t, n) == len(range(b + 1, n + 1, t)) == K.BasicMCRange(burnin=b, thinning=t, nsteps=n).npoststeps


def test_klb_erf_is_correctly_rounded_almost_everywhere(O):
    """klb_erf (double-double series, shared by the oracle and the kernels) against mpmath: <= 0.5 ulp (+ 1e-6) on a
    dense sample of [-6.5, 6.5] and of tiny arguments; it reproduces the reference's erf_rate_score known answers
    (test_erf_rate_score_kat) and libm's erf wherever libm itself is correctly rounded"""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(-6.5, 6.5, 1500), rng.uniform(-1, 1, 1500), 10.0 ** rng.uniform(-300, 0, 200)])
    for x in xs:
        t = mp.erf(mp.mpf(float(x)))
        e = mp.floor(mp.log(abs(t), 2))
        assert float(abs(mp.mpf(O.erf(x)) - t) / mp.mpf(2) ** (e - 52)) < 0.5 + 1e-6
    assert O.erf(0.0) == 0.0 and O.erf(7.0) == 1.0 and O.erf(-7.0) == -1.0 and O.erf(math.inf) == 1.0 and math.isnan(O.erf(math.nan))
    assert sum(O.erf(x) != math.erf(x) for x in xs) < len(xs) // 8


# ---- primitives -----------------------------------------------------------------------------
def test_philox_random123_kat(O):
    """Philox4x32 round function + key schedule: the 10-round known-answer vectors of Random123 (kat_vectors)"""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert [int(v) for v in O.philox(ctr, key, 10)] == want
    # the contract runs the first 7 of those rounds (klb_math.h: KLB_PHILOX_ROUNDS); same code, same key schedule
    assert O.philox_rounds() == 7
    assert [int(v) for v in O.philox([0, 0, 0, 0], [0, 0], 7)] != [int(v) for v in O.philox([0, 0, 0, 0], [0, 0], 10)]


def test_no_fp_contraction_in_oracle(O):
    """the oracle must round a*b and +c separately in reference arithmetic (built with -ffp-contract=off)"""
    a, b = 1.0 + 2.0 ** -30, 1.0 - 2.0 ** -30
    # a*b = 1 - 2^-60 rounds to 1.0, so a*b - 1 = 0 un-fused; the fused result is -2^-60
    assert O.lib().orc_ma(a, b, -1.0, 0) == 0.0
    assert O.lib().orc_ma(a, b, -1.0, 1) == -(2.0 ** -60)


def test_exp_log_accuracy(O):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    rng = np.random.default_rng(11)

    def ulps(y, t):
        e = max(mp.floor(mp.log(abs(t), 2)), -1022)
        return float(abs(mp.mpf(y) - t) / mp.mpf(2) ** (e - 52))
    for x in np.concatenate([rng.uniform(-700, 700, 1500), rng.uniform(-2, 2, 1500)]):
        assert ulps(O.exp(x), mp.exp(mp.mpf(float(x)))) < 0.52
    for x in np.concatenate([rng.uniform(0, 1, 1500), 10.0 ** rng.uniform(-300, 300, 1000), rng.uniform(0.9, 1.1, 1500)]):
        assert ulps(O.log(x), mp.log(mp.mpf(float(x)))) < 2.0
    assert O.exp(0.0) == 1.0 and O.exp(-1000.0) == 0.0 and O.exp(1000.0) == math.inf and math.isnan(O.exp(math.nan))
    assert O.log(1.0) == 0.0 and O.log(0.0) == -math.inf and math.isnan(O.log(-1.0)) and O.log(math.inf) == math.inf
    assert O.log(5e-324) == pytest.approx(math.log(5e-324), rel=1e-15)
    assert O.exp(-744.0) == pytest.approx(math.exp(-744.0), rel=1e-12)


def test_exp_matches_libm_almost_everywhere(O):
    rng = np.random.default_rng(2)
    xs = rng.uniform(-50, 50, 20000)
    bad = sum(O.exp(x) != math.exp(x) for x in xs)
    assert bad < 50          # both are < 1 ulp; they may disagree only on hard-to-round arguments


def test_normals_distribution(O):
    sp = pytest.importorskip("scipy.stats")
    z = np.concatenate([O.normals(12345, c, 1, 1 << 16) for c in range(16)])
    assert abs(z.mean()) < 4 / math.sqrt(z.size)
    assert abs(z.var() - 1) < 6 * math.sqrt(2 / z.size)
    assert sp.kstest(z, "norm").pvalue > 1e-3
    edges = sp.norm.ppf(np.linspace(0, 1, 101))
    cnt, _ = np.histogram(z, edges)
    assert sp.chisquare(cnt).pvalue > 1e-3
    # tail beyond the ziggurat base strip r = 3.654...: exercised and correctly weighted
    r = 3.6541528853610088
    n_tail = int((np.abs(z) > r).sum())
    exp_tail = 2 * sp.norm.sf(r) * z.size
    assert abs(n_tail - exp_tail) < 5 * math.sqrt(exp_tail)


def test_uniform_range_and_streams(O):
    u = np.array([O.uniform(7, c, t) for c in range(50) for t in range(1, 41)])
    assert (u >= 0).all() and (u < 1).all() and len(set(u)) == u.size
    assert abs(u.mean() - 0.5) < 0.03
    # streams: different chains / transitions / seeds give different normals; same key repeats
    a = O.normals(1, 2, 3, 64)
    assert np.array_equal(a, O.normals(1, 2, 3, 64))
    assert not np.array_equal(a, O.normals(1, 2, 4, 64)) and not np.array_equal(a, O.normals(1, 3, 3, 64))
    assert not np.array_equal(a, O.normals(2, 2, 3, 64))
    # prefix property: element i does not depend on how many elements are drawn
    assert np.array_equal(a[:10], O.normals(1, 2, 3, 10))


# ------------------------------------------------------------------ Bayesian logistic regression target
def test_logit_target_matches_closed_form(O):
    """ploglikelihood + plogprior and pgradlogtarget of doc/examples/swiss/HMC/noadaptation/analytical.jl:11-20,
    evaluated in 40-digit arithmetic, against the oracle's restatement"""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    sys_path_golden()
    import make_golden as G
    for d in (1, 3, 4, 16):
        X, y, lam = G.logit_data(d)
        tp = O.logit_params(X, y, lam)
        rng = np.random.default_rng(d)
        for p in rng.normal(scale=2.0, size=(4, d)):
            for arith in (0, 1):
                cfg = O.make_config(O.HMC, O.LOGIT, 1, d, 10, arith=arith)
                assert cfg.nv == 0
                lt, g = O.eval_target(cfg, p, tp)
                Xp = [sum(mp.mpf(X[i, j]) * mp.mpf(p[j]) for j in range(d)) for i in range(X.shape[0])]
                ll = sum(xp * mp.mpf(y[i]) for i, xp in enumerate(Xp)) - sum(mp.log(1 + mp.exp(xp)) for xp in Xp)
                lp = -mp.mpf(0.5) * (sum(mp.mpf(v) ** 2 for v in p) / lam + d * mp.log(2 * mp.pi * lam))
                assert abs(mp.mpf(lt) - (ll + lp)) < 1e-12 * abs(ll + lp)
                r = [mp.mpf(y[i]) - 1 / (1 + mp.exp(-xp)) for i, xp in enumerate(Xp)]
                for j in range(d):
                    gj = sum(mp.mpf(X[i, j]) * r[i] for i in range(X.shape[0])) - mp.mpf(p[j]) / lam
                    assert abs(mp.mpf(g[j]) - gj) < 1e-12 * max(1, abs(gj))


def sys_path_golden():
    import os
    import sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if gdir not in sys.path:
        sys.path.insert(0, gdir)


def test_logit_descriptor_is_the_reference_closure(K, O):
    """the descriptor called as a plain function (numpy) agrees with the oracle: it is a valid host-side
    `loglikelihood` / `logprior` / `gradlogtarget` triple for stock Klara"""
    sys_path_golden()
    import make_golden as G
    X, y, lam = G.logit_data(4)
    t = K.BayesLogit(X, y, lam)
    cfg = O.make_config(O.MALA, O.LOGIT, 1, 4, 10)
    p = np.array([5.1, -0.9, 8.2, -4.5])                     # v0[:p] of the swiss examples
    lt, g = O.eval_target(cfg, p, O.logit_params(X, y, lam))
    assert t(p) == pytest.approx(lt, rel=1e-12)
    assert t.loglikelihood(p) + t.logprior(p) == t(p)
    np.testing.assert_allclose(t.gradient(p), g, rtol=1e-11)


def test_logit_posterior_sampling_sanity(O):
    """HMC / MALA / MH on the logistic-regression posterior: all three agree with each other and with the
    Laplace approximation around the posterior mode (Newton iterations in numpy)"""
    sys_path_golden()
    import make_golden as G
    d = 4
    X, y, lam = G.logit_data(d)
    tp = O.logit_params(X, y, lam)
    b = np.zeros(d)
    for _ in range(50):                                      # Newton: posterior mode and Hessian
        mu = 1 / (1 + np.exp(-X @ b))
        H = X.T @ (X * (mu * (1 - mu))[:, None]) + np.eye(d) / lam
        b = b + np.linalg.solve(H, X.T @ (y - mu) - b / lam)
    sd = np.sqrt(np.diag(np.linalg.inv(H)))
    means = {}
    for name, smp, kw in (("HMC", O.HMC, dict(step=0.1, nleaps=8)), ("MALA", O.MALA, dict(step=0.03)),
                          ("MH", O.MH, dict())):
        cfg = O.make_config(smp, O.LOGIT, 32, d, 1500, burnin=500, monitor=1, diagnostics=1, seed=5,
                            nthreads=O.max_threads(), **kw)
        x0 = np.tile(b, (32, 1))
        r = O.run(cfg, x0, tp, sigma=0.5 * sd if smp == O.MH else None)
        acc = r["accept"].mean()
        assert 0.15 < acc < 0.999, (name, acc)
        means[name] = r["value"].reshape(-1, d).mean(0)
        assert np.all(np.abs(means[name] - b) < 0.35 * sd), (name, means[name], b, sd)   # skewed posterior: mean != mode
    assert np.all(np.abs(means["HMC"] - means["MALA"]) < 0.1 * sd)
    assert np.all(np.abs(means["HMC"] - means["MH"]) < 0.15 * sd)


# ------------------------------------------------------------------ DualAveragingMCTuner
def test_dual_averaging_tune_formula(O):
    """tune!(tune, tuner, count, a) (src/tuners/DualAveragingMCTuner.jl:95-101) re-evaluated with libm along the
    accept probabilities the oracle saw: hweight, hbar, step, εweight, εbar after every transition"""
    cfg = O.make_config(O.HMC, O.ISO, 1, 8, 60, step=0.4, nleaps=3, tuner=O.DUALAVG, target_rate=0.651, nadapt=40,
                        monitor=3, diagnostics=1, seed=12, gamma=0.05, da_t0=10, kappa=0.75)
    x0 = O.normals(12, 0, 0, 8)[None]
    t0, d0 = O.da_state(cfg)
    assert t0["step"][0] == 0.4 and d0["lambda"][0] == 3 * 0.4 and d0["mu"][0] == pytest.approx(math.log(4.0), rel=1e-15)
    assert d0["epsbar"][0] == 1.0 and d0["hbar"][0] == 0.0 and t0["totproposed"][0] == 100 and d0["count"][0] == 0
    tr, dr = O.da_state(cfg, first=False)                                   # reset!: step = 1 (HMC.jl:218)
    assert tr["step"][0] == 1.0 and dr["mu"][0] == pytest.approx(math.log(10.0), rel=1e-15)
    # replay one transition at a time (nsteps = 1 jobs chained through state, tune and da records)
    tune, da, x, lt = t0, d0, x0, None
    step, hbar, epsbar, mu = 0.4, 0.0, 1.0, math.log(4.0)
    for i in range(1, 61):
        c1 = O.make_config(O.HMC, O.ISO, 1, 8, 1, step=0.4, nleaps=3, tuner=O.DUALAVG, target_rate=0.651, nadapt=40,
                           monitor=3, diagnostics=1, seed=12, t0=i - 1)
        nl_expect = max(1, int(round(1.2 / step)))                           # Python round = ties to even, like Julia's
        lt_before = -float(np.dot(x[0], x[0]))
        r = O.run(c1, x, tune=tune, da=da, logtarget=None if lt is None else lt)
        tune, da, x, lt = r["tune"], r["da"], r["x"], r["logtarget_state"]
        assert da["nleaps"][0] == nl_expect and da["count"][0] == i
        # the accept probability is not exported; recompute a = min(1, exp(H' - H)) from scratch in numpy
        z = O.normals(12, 0, i, 8)
        q, p = r_prev.copy() if i > 1 else x0[0].copy(), z.copy()
        h = 0.5 * step
        for _ in range(nl_expect):
            p = p + h * (-2 * q); q = q + step * p; p = p + h * (-2 * q)
        a = min(1.0, math.exp((-np.dot(q, q) - 0.5 * np.dot(p, p)) - (lt_before - 0.5 * np.dot(z, z))))
        if i <= 40:
            hw = 1 / (i + 10)
            hbar = (1 - hw) * hbar + hw * (0.651 - a)
            step = math.exp(mu - math.sqrt(i) * hbar / 0.05)
            ew = i ** (-0.75)
            epsbar = math.exp((1 - ew) * math.log(epsbar) + ew * math.log(step))
            assert da["hweight"][0] == pytest.approx(hw, rel=1e-15) and da["epsweight"][0] == pytest.approx(ew, rel=1e-13)
        else:
            step = epsbar                                                    # iterate/HMC.jl:246
        assert da["hbar"][0] == pytest.approx(hbar, rel=1e-9, abs=1e-12)
        assert tune["step"][0] == pytest.approx(step, rel=1e-9)
        assert da["epsbar"][0] == pytest.approx(epsbar, rel=1e-9)
        step, hbar, epsbar = float(tune["step"][0]), float(da["hbar"][0]), float(da["epsbar"][0])   # no drift
        r_prev = x[0].copy()
    assert 0.2 < step < 1.5


def test_dual_averaging_reaches_target_rate(O):
    """after adaptation the acceptance rate sits near the target (Hoffman & Gelman's criterion); verbose counters
    follow iterate/HMC.jl:129-133,229-243"""
    # (with few leapfrog steps nleaps = round(λ/step) makes the acceptance rate a step function of the step size and
    # the averaged εbar can land on either side of a jump; λ = 3 with ~6-9 steps is smooth enough)
    for target in (0.651, 0.8):
        cfg = O.make_config(O.HMC, O.ISO, 64, 64, 1200, burnin=600, step=0.1, nleaps=30, tuner=O.DUALAVG,
                            target_rate=target, nadapt=600, monitor=1, diagnostics=1, seed=5, verbose=1, period=100,
                            nthreads=O.max_threads())
        x0 = np.stack([O.normals(5, c, 0, 64) for c in range(64)])
        r = O.run(cfg, x0)
        assert abs(r["accept"].mean() - target) < 0.06
        assert np.all(r["da"]["count"] == 1200) and np.all(r["tune"]["step"] == r["da"]["epsbar"])
        # 6 periods of 100 inside nadapt were reset into totproposed (initially = period); the rest keeps counting
        assert np.all(r["tune"]["totproposed"] == 700) and np.all(r["tune"]["proposed"] == 600)
    cfg = O.make_config(O.HMC, O.ISO, 2, 4, 50, step=0.3, nleaps=2, tuner=O.DUALAVG, nadapt=10, monitor=1, seed=5)
    r = O.run(cfg, np.zeros((2, 4)) + 0.1)
    assert np.all(r["tune"]["proposed"] == 0) and np.all(r["tune"]["totproposed"] == 100)    # not verbose: no counters
