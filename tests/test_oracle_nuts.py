"""NUTS: the reference's multivariate transition, with its object aliasing, against the resolved state machine, and the C
oracle's restatement of that state machine against the Python one (CPU only).

1. oracle/nuts_alias.py::alias_transition follows iterate/NUTS.jl:230-400 and build_tree! (NUTS.jl:514-628, :781-927)
   statement by statement on objects with Julia's reference semantics (MuvNUTSState(pstate, pstate, pstate, pstate, ...),
   NUTS.jl:198-225).  simple_transition is what is left when the aliasing is resolved by hand.  Same draws, same
   floating-point operations: they must agree exactly, including the number of random numbers consumed.
2. oracle/klb_oracle.c::orc_iterate_nuts restates simple_transition on the repo's RNG contract and canonical reduction
   order.  Fed the same draws, the Python model must make the same decisions and land within 1e-9 of its values."""
import math

import numpy as np
import pytest

from oracle import nuts_alias as NA
from oracle import oracle as O


class Draws:
    def __init__(self, seed):
        self.r = np.random.default_rng(seed)
        self.count = 0

    def randn(self, d):
        self.count += 1
        return self.r.standard_normal(d)

    def rand(self):
        self.count += 1
        return float(self.r.random())

    def randbool(self):
        self.count += 1
        return bool(self.r.integers(0, 2))


def _fresh(tgt, x0):
    ps = NA.PState(x0.size)
    ps.value = x0.copy()
    tgt.gradlogtarget(ps)
    tgt.logtarget(ps)
    return ps


def test_resolved_state_machine_equals_the_aliased_reference_code():
    early = 0
    for trial in range(250):
        rng = np.random.default_rng(1000 + trial)
        d = int(rng.integers(1, 9))
        step = float(rng.choice([0.05, 0.2, 0.5, 0.9, 1.5]))
        maxnd = int(rng.integers(1, 7))
        maxdelta = int(rng.choice([1000, 1, 3]))
        da = bool(rng.integers(0, 2))
        tgt = NA.Target(rng.standard_normal(d) if rng.integers(0, 2) else None)
        x0 = rng.standard_normal(d)
        ps = _fresh(tgt, x0)
        ss = NA.SState(d, NA.Tune(step))
        q = _fresh(tgt, x0)
        st = dict(value=q.value, gradlogtarget=q.gradlogtarget, logtarget=q.logtarget)
        d1, d2 = Draws(trial), Draws(trial)
        for it in range(10):
            r1 = NA.alias_transition(ps, ss, tgt, maxdelta, maxnd, d1, da)
            r2 = NA.simple_transition(st, step, tgt, maxdelta, maxnd, d2, da)
            assert r1 == r2, (trial, it, r1, r2)
            assert d1.count == d2.count                                   # same number of random numbers consumed
            assert np.array_equal(ps.value, st["value"]) and np.array_equal(ps.gradlogtarget, st["gradlogtarget"])
            assert ps.logtarget == st["logtarget"] or (math.isnan(ps.logtarget) and math.isnan(st["logtarget"]))
            early += r1[1] < maxnd
        # every state of the sampler state is the one object the reference constructed       NUTS.jl:198-225
        assert ss.pstateplus is ss.pstateminus is ss.pstateprime is ss.pstatedprime
    assert early > 50                                                     # trees that stop early are exercised too


class OracleDraws:
    """the repo's RNG contract: normals of (seed, chain, t); uniforms q = 0, 1, 2, ... of the same stream"""

    def __init__(self, seed, chain, t):
        self.seed, self.chain, self.t, self.q = seed, chain, t, 0

    def randn(self, d):
        return O.normals(self.seed, self.chain, self.t, d)

    def rand(self):
        u = O.uniform_seq(self.seed, self.chain, self.t, self.q)
        self.q += 1
        return u

    def randbool(self):
        return self.rand() < 0.5


@pytest.mark.parametrize("target,dim,step,maxnd,maxdelta", [("iso", 5, 0.3, 5, 1000), ("iso", 70, 0.12, 4, 1000),
                                                            ("shifted", 9, 0.4, 6, 1000), ("iso", 3, 1.4, 5, 2),
                                                            ("shifted", 130, 0.1, 3, 1000),
                                                            ("logit", 4, 0.2, 5, 1000), ("logit", 7, 0.3, 5, 3)])
def test_c_oracle_follows_the_state_machine(target, dim, step, maxnd, maxdelta):
    N, nsteps, seed = 6, 25, 424242
    rng = np.random.default_rng(dim)
    mu = rng.standard_normal(dim) if target == "shifted" else None
    x0 = rng.standard_normal((N, dim))
    if target == "logit":                 # doc/examples/swiss/NUTS/*/analytical.jl: data-dependent target, sequential sums (nv = 0)
        X = rng.standard_normal((200, dim))
        X = (X - X.mean(0)) / X.std(0, ddof=1)
        y = (rng.uniform(size=200) < 1 / (1 + np.exp(-X @ rng.standard_normal(dim)))).astype(np.float64)
        cfg = O.make_config(O.NUTS, O.LOGIT, N, dim, nsteps, 0, step=step, monitor=3, diagnostics=3, seed=seed,
                            maxdelta=maxdelta, maxndoublings=maxnd, nv=0)
        ref = O.run(cfg, x0, tparams=O.logit_params(X, y, 100.0))
        tgt = NA.LogitTarget(X, y, 100.0)
    else:
        cfg = O.make_config(O.NUTS, O.SHIFTED if target == "shifted" else O.ISO, N, dim, nsteps, 0, step=step, monitor=3,
                            diagnostics=3, seed=seed, maxdelta=maxdelta, maxndoublings=maxnd)
        ref = O.run(cfg, x0, tparams=mu)
        tgt = NA.Target(mu)
    checked = 0
    for c in range(N):
        q = _fresh(tgt, x0[c])
        st = dict(value=q.value, gradlogtarget=q.gradlogtarget, logtarget=q.logtarget)
        for it in range(nsteps):
            # teacher forcing: every transition starts from the oracle's own state, so one borderline decision cannot
            # snowball; decisions are compared only where the transition is not within rounding of a threshold
            upd, j, _, _ = NA.simple_transition(st, step, tgt, maxdelta, maxnd, OracleDraws(seed, c, it + 1))
            same = upd == bool(ref["accept"][c, it]) and j == int(ref["ndoublings"][c, it])
            close = np.allclose(st["value"], ref["value"][c, it], rtol=1e-9, atol=1e-12)
            if same and close:
                assert abs(st["logtarget"] - ref["logtarget"][c, it]) <= 1e-9 * max(1.0, abs(st["logtarget"]))
                checked += 1
            st["value"] = ref["value"][c, it].copy()
            q = _fresh(tgt, st["value"])
            st["gradlogtarget"], st["logtarget"] = q.gradlogtarget, q.logtarget
    assert checked >= 0.97 * N * nsteps, checked                          # borderline comparisons are rare


@pytest.mark.parametrize("tuner", ["vanilla", "dualavg"])
def test_nuts_oracle_samples_the_target(tuner):
    """the aliased algorithm is not the textbook sampler, but on a Gaussian it still leaves the target's moments close"""
    N, d = 64, 8
    kw = dict(tuner=O.DUALAVG, nadapt=150, target_rate=0.65) if tuner == "dualavg" else {}
    cfg = O.make_config(O.NUTS, O.ISO, N, d, 400, 200, step=0.25, monitor=1, diagnostics=3, seed=11, maxndoublings=5,
                        nthreads=O.max_threads(), **kw)
    r = O.run(cfg, np.random.default_rng(3).standard_normal((N, d)) * 0.7)
    v = r["value"]
    assert abs(v.mean()) < 0.05 and 0.35 < v.var() < 0.65                 # N(0, I/2)
    assert (r["ndoublings"] >= 1).all() and (r["ndoublings"] <= 5).all()
    if tuner == "dualavg":
        assert np.isnan(r["da"]["lambda"]).all() and (r["da"]["count"] == 400).all()
        assert len(np.unique(r["tune"]["step"])) == N


@pytest.mark.parametrize("target,dim,step,maxnd", [("iso", 6, 0.35, 5), ("shifted", 40, 0.2, 4), ("logit", 4, 0.3, 6)])
def test_c_oracle_a_and_na_diagnostics(target, dim, step, maxnd):
    """:a and :na of NUTS with DualAveragingMCTuner (src/samplers/NUTS.jl:317,344; iterate/NUTS.jl:393-399): the sum of
    min(1, exp(H' - H0)) over the leaves of the LAST doubling and their number.  First transition of every chain (its step
    is still the sampler's), C oracle against the state machine fed the same draws."""
    N, seed = 48, 777
    rng = np.random.default_rng(dim)
    x0 = rng.standard_normal((N, dim))
    if target == "logit":
        X = rng.standard_normal((200, dim))
        X = (X - X.mean(0)) / X.std(0, ddof=1)
        y = (rng.uniform(size=200) < 1 / (1 + np.exp(-X @ rng.standard_normal(dim)))).astype(np.float64)
        tcode, tp, tgt, nv = O.LOGIT, O.logit_params(X, y, 100.0), NA.LogitTarget(X, y, 100.0), 0
    else:
        mu = rng.standard_normal(dim) if target == "shifted" else None
        tcode, tp, tgt, nv = (O.SHIFTED if mu is not None else O.ISO), mu, NA.Target(mu), None
    cfg = O.make_config(O.NUTS, tcode, N, dim, 1, 0, step=step, tuner=O.DUALAVG, target_rate=0.65, nadapt=10, monitor=3,
                        diagnostics=15, seed=seed, maxndoublings=maxnd, nv=nv)
    ref = O.run(cfg, x0, tparams=tp)
    assert ref["a"].shape == (N, 1) and ref["na"].dtype == np.int32
    checked = 0
    for c in range(N):
        q = _fresh(tgt, x0[c])
        st = dict(value=q.value, gradlogtarget=q.gradlogtarget, logtarget=q.logtarget)
        upd, j, a, na = NA.simple_transition(st, step, tgt, 1000, maxnd, OracleDraws(seed, c, 1), da=True)
        if upd == bool(ref["accept"][c, 0]) and j == int(ref["ndoublings"][c, 0]):
            assert na == int(ref["na"][c, 0])
            assert abs(a - ref["a"][c, 0]) <= 1e-9 * max(1.0, abs(a))
            assert 1 <= na <= 2 ** (j - 1) and 0.0 <= ref["a"][c, 0] <= na
            checked += 1
    assert checked >= 0.95 * N
    # the tuner was fed a/na of that doubling: tune! with count = 1        DualAveragingMCTuner.jl:95-101
    assert np.all(ref["da"]["nleaps"] == ref["na"][:, 0])
