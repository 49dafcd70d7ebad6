#!/usr/bin/env python
"""The ESS / statistics kernel variants (KLB_ESS_VARIANT, klb_aux.cu: lags per pass x raw / centred tile x series per CTA):
bit-identity with variant 0 and the oracle on small jobs with awkward lengths, then the time on the C3 output.

    python tools/ess_variants.py [--nchains 65536] [--variants 0,2,6,7,9,10,11]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import klara_b200 as K                      # noqa: E402
from oracle import oracle as O              # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nchains", type=int, default=65536)
ap.add_argument("--variants", default="0,2,6,7,9,10,11")
a = ap.parse_args()
variants = [int(v) for v in a.variants.split(",")]
L = K._lib


def same(x, y):
    return bool(np.array_equal(x.view(np.uint64), y.view(np.uint64)))


def stats(job):
    return [job.ess(), job.mean(), job.mcvar("iid"), job.mcvar("imse"), job.iact()]


# ---- parity: every variant against variant 0 (which the GPU suite compares with the oracle) and against the oracle itself
ok = {v: True for v in variants}
for npost, dim, step in ((100, 130, 0.05), (37, 64, 0.3), (5, 9, 0.1), (4, 3, 0.1), (131, 33, 0.02), (64, 1024, 0.05)):
    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    job = K.BasicMCJob(K.likelihood_model(p, False), K.HMC(step, 10), K.BasicMCRange(nsteps=npost + 3, burnin=3),
                       {"p": K.SyntheticNormal(24, dim)}, outopts={"monitor": ["value"]}, seed=7)
    job.run()
    val = job.output().value                                   # (nchains, npost, dim)
    ref = O.stats(val)                                         # keys: mean, mcvar_iid, mcvar_imse, ess, iact
    os.environ["KLB_ESS_VARIANT"] = "0"
    base = stats(job)
    refs = [ref[k] for k in ("ess", "mean", "mcvar_iid", "mcvar_imse", "iact")]
    for v in variants:
        os.environ["KLB_ESS_VARIANT"] = str(v)
        got = stats(job)
        good = all(same(g, b) for g, b in zip(got, base)) and all(same(g, np.ascontiguousarray(r)) for g, r in zip(got, refs))
        ok[v] = ok[v] and good
        if not good:
            print("variant %d differs: npost %d dim %d" % (v, npost, dim), flush=True)
    job.close()
print("parity:", {v: ("ok" if ok[v] else "FAIL") for v in variants}, flush=True)

# ---- time on the C3 output
p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
job = K.BasicMCJob(K.likelihood_model(p, False), K.HMC(0.05, 10), K.BasicMCRange(nsteps=200, burnin=100),
                   {"p": K.SyntheticNormal(a.nchains, 1024)}, outopts={"monitor": ["value"]}, seed=20240925)
job.run()
gb = a.nchains * 100 * 1024 * 8 / 1e9
for v in variants:
    os.environ["KLB_ESS_VARIANT"] = str(v)
    job.ess(to_host=False)
    best = 1e9
    for r in range(3):
        t = time.perf_counter(); job.ess(to_host=False); best = min(best, (time.perf_counter() - t) * 1e3)
    print("variant %d: %.2f ms  (%.1f GB of samples -> %.0f GB/s) parity %s" % (v, best, gb, gb / best * 1e3, "ok" if ok[v] else "FAIL"), flush=True)
job.close()
