#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: small seeded input/output vectors of the hot path.

Provenance: the reference (Klara.jl, Julia 0.6) cannot be executed in this image and never seeds its RNG,
so these vectors are produced by THIS repo's CPU oracle (oracle/klb_oracle.c), after it has passed the
reference's own known-answer tests (tests/test_oracle_kat.py), and every vector is cross-checked here, before it
is written, against the independent numpy / libm twin (oracle/twin.py: no shared source, libm exp/log, BLAS-order
dot products): accept/reject sequences identical, values and log-targets within 1e-6 relative.  They freeze the
RNG (Philox4x32-7) / reduction-order / arithmetic contract: the CPU suite checks the oracle
still reproduces them, the GPU suite checks the CUDA path reproduces them bit for bit.  They are regression
vectors of THIS repo's contract, not outputs of the reference.   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (sampler, target, nchains, dim, nsteps, kwargs, extra)
    "hmc_iso_d1024": ("HMC", "iso", 3, 1024, 12, dict(burnin=4, thinning=2, step=0.05, nleaps=10, monitor=3, diagnostics=1, seed=20240925), {}),
    "hmc_iso_d7_accrate": ("HMC", "iso", 5, 7, 90, dict(burnin=60, step=0.2, nleaps=4, tuner=O.ACCRATE, target_rate=0.8, period=20, monitor=3, diagnostics=1, seed=11), {}),
    "mala_iso_d128": ("MALA", "iso", 4, 128, 40, dict(burnin=10, step=0.18, monitor=7, diagnostics=1, seed=5), {}),
    "mala_rosen_d16_accrate": ("MALA", "rosen", 4, 16, 120, dict(burnin=100, step=0.01, tuner=O.ACCRATE, target_rate=0.574, period=25, monitor=3, diagnostics=1, seed=8), {}),
    "mh_readme_d2": ("MH", "iso", 1, 2, 200, dict(burnin=100, monitor=3, diagnostics=1, seed=2024), {"x0": [[5.1, -0.9]], "sigma": [1.0, 1.0]}),
    "hmc_shifted_d100_fma": ("HMC", "shifted", 3, 100, 20, dict(burnin=5, step=0.05, nleaps=6, monitor=3, diagnostics=1, seed=3, arith=1), {}),
    # Bayesian logistic regression (doc/examples/swiss/HMC/noadaptation/analytical.jl), synthetic 200 x 4 data
    "hmc_logit_d4": ("HMC", "logit", 6, 4, 60, dict(burnin=20, step=0.03, nleaps=10, monitor=7, diagnostics=1, seed=41), {}),
    "mala_logit_d3_accrate": ("MALA", "logit", 5, 3, 150, dict(burnin=100, step=0.01, tuner=O.ACCRATE, target_rate=0.574, period=25, monitor=3, diagnostics=1, seed=42), {}),
    "mh_logit_d4_fma": ("MH", "logit", 4, 4, 120, dict(burnin=40, thinning=2, monitor=3, diagnostics=1, seed=43, arith=1), {"sigma": [0.1, 0.15, 0.2, 0.1]}),
    # NUTS as the reference computes it; cross-checked against oracle/nuts_alias.py instead of the twin
    "nuts_iso_d70": ("NUTS", "iso", 4, 70, 30, dict(burnin=10, step=0.15, monitor=3, diagnostics=3, seed=1729, maxndoublings=4), {}),
    "nuts_shifted_d9_delta2": ("NUTS", "shifted", 5, 9, 40, dict(burnin=0, step=1.1, monitor=3, diagnostics=3, seed=1730, maxndoublings=6, maxdelta=2), {}),
    # doc/examples/swiss/NUTS/noadaptation/analytical.jl on the synthetic 200 x 4 data (thread-per-chain kernel)
    "nuts_logit_d4": ("NUTS", "logit", 6, 4, 30, dict(burnin=10, step=0.2, monitor=7, diagnostics=3, seed=1731, maxndoublings=5), {}),
}
SAMPLERS = {"MH": O.MH, "MALA": O.MALA, "HMC": O.HMC, "NUTS": O.NUTS}
TARGETS = {"iso": O.ISO, "shifted": O.SHIFTED, "rosen": O.ROSEN, "logit": O.LOGIT}


def logit_data(d, ndata=200, lam=100.0, seed=777):
    """deterministic stand-in for the swiss data (200 x 4 standardised covariates, 0/1 outcome), drawn from the
    oracle's own Philox streams so that it does not depend on numpy's generators"""
    X = np.stack([O.normals(seed, i, 0, d) for i in range(ndata)])
    X = (X - X.mean(0)) / X.std(0, ddof=1)
    beta = O.normals(seed, ndata, 0, d)
    u = np.array([O.uniform(seed, i, 1) for i in range(ndata)])
    y = (u < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
    return np.ascontiguousarray(X), y, lam


def build(name):
    smp, tgt, n, d, nsteps, kw, extra = CASES[name]
    cfg = O.make_config(SAMPLERS[smp], TARGETS[tgt], n, d, nsteps, **kw)
    x0 = np.array(extra["x0"], dtype=float) if "x0" in extra else np.stack([O.normals(cfg.seed, c, 0, d) for c in range(n)])
    tparams = None
    if tgt == "shifted":
        tparams = np.linspace(-1.0, 1.0, d)
    if tgt == "rosen":
        tparams = np.array([1.0, 100.0, 0.05])
    if tgt == "logit":
        tparams = O.logit_params(*logit_data(d))
    sigma = np.array(extra["sigma"], dtype=float) if "sigma" in extra else (np.full(d, 0.5) if smp == "MH" else None)
    return cfg, x0, tparams, sigma


def twin_crosscheck(name, cfg, x0, tparams, sigma, r):
    sys.path.insert(0, os.path.dirname(HERE))
    from twin_helpers import check_against_twin, twin_cfg, twin_target
    smp, tgt, n, d, nsteps, kw, extra = CASES[name]
    tuner = {O.VANILLA: "vanilla", O.ACCRATE: "accrate"}[kw.get("tuner", O.VANILLA)]
    tc = twin_cfg(smp, nsteps, burnin=kw.get("burnin", 0), thinning=kw.get("thinning", 1), step=kw.get("step", 0.1),
                  nleaps=kw.get("nleaps", 10), tuner=tuner, target_rate=kw.get("target_rate", 0.574),
                  period=kw.get("period", 100), seed=kw["seed"], sigma=sigma)
    w = check_against_twin(name, tc, twin_target(tgt, d, tparams), x0, range(n), r["value"], r["logtarget"], r["accept"],
                           final_step=r["tune"]["step"] if smp != "MH" else None)
    assert w["flips"] == 0
    return w


def nuts_crosscheck(name, cfg, x0, tparams, r):
    """every transition of every chain against the resolved state machine of oracle/nuts_alias.py (which
    tests/test_oracle_nuts.py pins to the aliasing-faithful model of the reference code), teacher-forced"""
    from oracle import nuts_alias as NA
    smp, tgt, n, d, nsteps, kw, extra = CASES[name]
    target = NA.LogitTarget(*logit_data(d)) if tgt == "logit" else NA.Target(tparams if tgt == "shifted" else None)

    class Draws:
        def __init__(self, chain, t):
            self.chain, self.t, self.q = chain, t, 0

        def randn(self, dd):
            return O.normals(cfg.seed, self.chain, self.t, dd)

        def rand(self):
            self.q += 1
            return O.uniform_seq(cfg.seed, self.chain, self.t, self.q - 1)

        def randbool(self):
            return self.rand() < 0.5

    def fresh(x):
        ps = NA.PState(d)
        ps.value = np.array(x, dtype=float)
        target.gradlogtarget(ps)
        target.logtarget(ps)
        return dict(value=ps.value, gradlogtarget=ps.gradlogtarget, logtarget=ps.logtarget)

    assert kw.get("burnin", 0) == 0 or kw.get("thinning", 1) == 1
    b = kw.get("burnin", 0)
    full = O.run(O.make_config(O.NUTS, TARGETS[tgt], n, d, nsteps, **dict(kw, burnin=0)), x0, tparams)   # every transition
    agree = 0
    for c in range(n):
        st = fresh(x0[c])
        for it in range(nsteps):
            upd, j, _, _ = NA.simple_transition(st, kw["step"], target, kw.get("maxdelta", 1000), kw["maxndoublings"], Draws(c, it + 1))
            ok = upd == bool(full["accept"][c, it]) and j == int(full["ndoublings"][c, it]) and \
                np.allclose(st["value"], full["value"][c, it], rtol=1e-9, atol=1e-12)
            agree += ok
            st = fresh(full["value"][c, it])
    assert agree >= 0.97 * n * nsteps, (name, agree)
    assert np.array_equal(full["value"][:, b:], r["value"])
    return {"transitions": n * nsteps, "agree": int(agree)}


if __name__ == "__main__":
    only = sys.argv[1:]
    for name in CASES:
        if only and name not in only:
            continue
        cfg, x0, tparams, sigma = build(name)
        r = O.run(cfg, x0, tparams, sigma)
        if CASES[name][0] == "NUTS":
            print(name, "state machine:", nuts_crosscheck(name, cfg, x0, tparams, r))
        else:
            print(name, "twin:", twin_crosscheck(name, cfg, x0, tparams, sigma, r))
        out = {"x0": x0, "x": r["x"], "logtarget_state": r["logtarget_state"], "tune": r["tune"]}
        for k in ("value", "logtarget", "gradlogtarget", "accept", "ndoublings"):
            if r[k] is not None:
                out[k] = r[k]
        if tparams is not None:
            out["tparams"] = tparams
        if sigma is not None:
            out["sigma"] = sigma
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
