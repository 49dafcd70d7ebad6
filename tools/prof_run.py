#!/usr/bin/env python
"""Small driver for ncu captures: one BasicMCJob of the C3 shape with fewer chains / transitions.
    python tools/prof_run.py --nchains 9472 --nsteps 20 --burnin 10 --reps 3 [--sampler HMC] [--arith reference]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import klara_b200 as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nchains", type=int, default=9472)
ap.add_argument("--dim", type=int, default=1024)
ap.add_argument("--nsteps", type=int, default=20)
ap.add_argument("--burnin", type=int, default=10)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--sampler", default="HMC")
ap.add_argument("--arith", default="reference")
ap.add_argument("--step", type=float, default=0.05)
ap.add_argument("--nleaps", type=int, default=10)
ap.add_argument("--none", action="store_true", help="destination none")
ap.add_argument("--target", default="iso")
ap.add_argument("--accrate", type=float, default=0.0, help="AcceptanceRateMCTuner target (0 = Vanilla)")
a = ap.parse_args()
x0 = np.random.default_rng(0).standard_normal((a.nchains, a.dim)) * (0.3 if a.target == "dense" else 1.0)
def _target():
    if a.target == "dense":
        idx = np.arange(a.dim)
        C = np.linalg.inv(0.8 ** np.abs(idx[:, None] - idx[None, :]))
        return K.DenseGaussian((C + C.T) / 2)
    if a.target == "shifted":
        return K.ShiftedIsoGaussian(np.random.default_rng(5).standard_normal(a.dim))
    return {"iso": K.IsoGaussian(), "rosen": K.Rosenbrock()}[a.target]


p = K.BasicContMuvParameter("p", logtarget=_target())
tuner = K.AcceptanceRateMCTuner(a.accrate) if a.accrate > 0 else K.VanillaMCTuner()
smp = {"HMC": K.HMC(a.step, a.nleaps), "MALA": K.MALA(a.step), "MH": K.MH(np.full(a.dim, 0.02)), "NUTS": K.NUTS(a.step)}[a.sampler]
oo = {"destination": "none"} if a.none else {"monitor": ["value", "logtarget"], "diagnostics": ["accept"]}
job = K.BasicMCJob(K.likelihood_model(p, False), smp, K.BasicMCRange(nsteps=a.nsteps, burnin=a.burnin), {"p": x0},
                   tuner=tuner, outopts=oo, seed=1, arith=a.arith)
for r in range(a.reps):
    job.reset()
    job.run()
    ms = job.last_run_ms
    lf = a.nchains * a.nsteps * (a.nleaps if a.sampler == "HMC" else 31 if a.sampler == "NUTS" else 1)   # NUTS on a Gaussian: every tree runs 2^5 - 1 leaves
    print("rep %d: %.3f ms  %.4g %s/s" % (r, ms, lf / ms * 1e3, "leapfrog-steps" if a.sampler in ("HMC", "NUTS") else "transitions"))
if not a.none:
    print("accept rate %.3f" % job.output().diagnosticvalues.mean())
