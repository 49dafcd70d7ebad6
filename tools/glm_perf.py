import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import klara_b200 as K
import make_golden as G
X, y, lam = G.logit_data(4)
ONLY = sys.argv[1] if len(sys.argv) > 1 else None
for smp, name in ((K.HMC(0.05, 10), "HMC"), (K.MALA(0.02), "MALA"), (K.MH(np.full(4, 0.1)), "MH"), (K.NUTS(0.1, maxndoublings=4), "NUTS")):
    if ONLY and name != ONLY:
        continue
    N = 148 * 64 * 16
    x0 = np.random.default_rng(0).normal(size=(N, 4)) * 0.3
    p = K.BasicContMuvParameter("p", logtarget=K.BayesLogit(X, y, lam))
    job = K.BasicMCJob(K.likelihood_model(p, False), smp, K.BasicMCRange(nsteps=200, burnin=100), {"p": x0},
                       outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept"]}, seed=1)
    job.run(); job.reset(); job.run()
    ms = job.last_run_ms
    evals = N * 200 * (10 if name == "HMC" else 15 if name == "NUTS" else 1)    # NUTS: 2^4 - 1 leaves per transition unless a tree stops early
    pl = job.plan()
    print(name, "N", N, "ms", ms, "target-evals/s %.3e" % (evals / ms * 1e3), "regs", pl.regs_per_thread, "bps", pl.blocks_per_sm,
          "acc", K.acceptance(job).mean(), flush=True)
