"""Comparison of an implementation's output (the C oracle's or the CUDA path's) with the independent numpy / libm
twin (oracle/twin.py).  Tolerances are north_star's: accept/reject sequence identical, log-target / gradient /
values within 1e-6 relative."""
import math

import numpy as np

from oracle import twin as T

RTOL = 1e-6          # north_star: "log-target/gradient within 1e-6 relative fp64"


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b))) / scale if a.size else 0.0


def twin_target(name, dim, tparams):
    if name == "iso":
        return T.IsoGaussian()
    if name == "shifted":
        return T.ShiftedIsoGaussian(tparams)
    if name == "dense":
        return T.DenseGaussian(np.asarray(tparams).reshape(dim, dim))
    if name == "rosen":
        return T.Rosenbrock(*tparams)
    if name == "logit":
        lam, nd = float(tparams[0]), int(tparams[1])
        return T.BayesLogit(np.asarray(tparams[2:2 + nd * dim]).reshape(nd, dim), tparams[2 + nd * dim:], lam)
    raise KeyError(name)


def twin_cfg(sampler, nsteps, burnin=0, thinning=1, step=0.1, nleaps=10, tuner="vanilla", target_rate=0.574,
             period=100, verbose=False, seed=1234, sigma=None, nadapt=1000, **_):
    return dict(sampler={"MH": T.MH, "MALA": T.MALA, "HMC": T.HMC}[sampler], nsteps=nsteps, burnin=burnin,
                thinning=thinning, step=step, nleaps=nleaps,
                tuner={"vanilla": T.VANILLA, "accrate": T.ACCRATE, "dualavg": T.DUALAVG}[tuner],
                target_rate=target_rate, score_k=7.0, period=period, verbose=verbose, seed=seed, sigma=sigma,
                nadapt=nadapt, eps0bar=1.0, h0bar=0.0, gamma=0.05, da_t0=10, kappa=0.75)


def check_against_twin(name, cfg, target, x0, chains, value, logtarget, accept, grad=None, chain_offset=0, t0=0,
                       forced=False, final_step=None, min_margin=1e-9):
    """value / logtarget / accept (/ grad): the other implementation's saved output for the chains `chains`
    (arrays indexed [chain, saved step, ...]).  Free-running comparison by default; with forced=True (needs burnin 0,
    thinning 1, i.e. every transition saved) each twin transition restarts from the other side's previous state.
    Returns the worst relative errors seen."""
    worst = {"value": 0.0, "logtarget": 0.0, "grad": 0.0, "flips": 0, "transitions": 0}
    for ci, c in enumerate(chains):
        f = value[ci] if forced else None
        r = T.run_chain(cfg, target, x0[ci], chain_offset + c, t0=t0, forced=f)
        acc_other = np.asarray(accept[ci], dtype=bool)
        if forced:
            # transition i started from the other side's state i-1: compare the decision and the accepted proposal
            acc_twin, margins = r["accepts"].astype(bool), r["margins"]
            states = r["states"]
        else:
            burnin, thinning = cfg.get("burnin", 0), cfg.get("thinning", 1)
            sel = np.arange(burnin, cfg["nsteps"], thinning)
            acc_twin, margins, states = r["accept"].astype(bool), r["margins"][sel], r["value"]
        worst["transitions"] += len(acc_twin)
        diff = acc_twin != acc_other
        if diff.any():
            # a different decision is only legitimate when the test was a tie at the tolerance
            i = int(np.argmax(diff))
            assert margins[i] <= RTOL * max(1.0, abs(float(r["ratios"][i]))), \
                "%s chain %d: accept/reject differs from the twin at saved step %d with margin %g" % (name, c, i, margins[i])
            worst["flips"] += 1
            if not forced:
                continue                                  # free-running states legitimately part ways after a tie
            states = states.copy()
            states[diff] = value[ci][diff]
        ev = relerr(states, value[ci])
        el = relerr(r["state_lt"] if forced else r["logtarget"], logtarget[ci])
        assert ev <= RTOL, "%s chain %d: values differ from the twin by %g relative" % (name, c, ev)
        assert el <= RTOL, "%s chain %d: log-targets differ from the twin by %g relative" % (name, c, el)
        worst["value"], worst["logtarget"] = max(worst["value"], ev), max(worst["logtarget"], el)
        if grad is not None and cfg["sampler"] != T.MH and not forced:
            g_twin = np.array([target.grad(v) for v in r["value"]])
            eg = relerr(g_twin, grad[ci])
            assert eg <= RTOL, "%s chain %d: gradients differ from the twin by %g relative" % (name, c, eg)
            worst["grad"] = max(worst["grad"], eg)
        if final_step is not None:
            assert math.isclose(r["tune"].step, final_step[ci], rel_tol=RTOL), \
                "%s chain %d: tuned step %r vs twin %r" % (name, c, final_step[ci], r["tune"].step)
    return worst
