#!/usr/bin/env python
"""NUTS on the logistic-regression target (klb_glm_kernel<3, DP, FMA>) against the CPU oracle, bit for bit: every padded
dimension, both tuners, both arithmetic modes, trees that stop early.  No torch, no pytest: starts in a second or two
(the same comparisons are tests/test_gpu_parity.py::test_nuts_logit_bit_exact)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

t0 = time.time()
import klara_b200 as K                      # noqa: E402
from helpers import build_pair, compare_run  # noqa: E402

CASES = [  # dim, tuner, maxnd, maxdelta, step, arith
    (4, "dualavg", 5, 1000, 0.15, "reference"), (4, "vanilla", 4, 1000, 0.2, "fma"), (3, "vanilla", 6, 1000, 0.1, "reference"),
    (2, "dualavg", 3, 1000, 0.3, "fma"), (7, "dualavg", 4, 1000, 0.1, "reference"), (8, "vanilla", 5, 3, 0.3, "reference"),
    (16, "vanilla", 3, 1000, 0.08, "fma"), (13, "dualavg", 4, 1000, 0.08, "reference"),
]
nok = 0
for dim, tuner, maxnd, maxdelta, step, arith in CASES:
    job, cfg, x0, tp, sg = build_pair(K, "NUTS", "logit", nchains=70, dim=dim, nsteps=30, burnin=9, thinning=2, step=step,
                                      seed=5150 + dim, arith=arith, tuner=tuner, target_rate=0.651, nadapt=20, period=5,
                                      verbose=(dim % 2 == 0), monitor=("value", "logtarget", "gradlogtarget"),
                                      diagnostics=("accept", "ndoublings") + (("na", "a") if tuner == "dualavg" else ()),
                                      maxdelta=maxdelta, maxndoublings=maxnd)
    try:
        out, ref = compare_run(job, cfg, x0, tp, sg)
        nd = ref["ndoublings"]
        print("ok   dim %2d %-8s maxnd %d maxdelta %4d %-9s ndoublings %d..%d accept %.2f regs %d"
              % (dim, tuner, maxnd, maxdelta, arith, nd.min(), nd.max(), ref["accept"].mean(), job.plan().regs_per_thread), flush=True)
        nok += 1
    except AssertionError as e:
        print("FAIL dim %2d %-8s maxnd %d %-9s: %s" % (dim, tuner, maxnd, arith, e), flush=True)
    job.close()
print("nuts_glm_check: %d of %d cases bit-exact, %.1f s" % (nok, len(CASES), time.time() - t0), flush=True)
sys.exit(0 if nok == len(CASES) else 1)
