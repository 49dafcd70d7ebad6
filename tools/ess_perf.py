#!/usr/bin/env python
"""ESS / statistics kernels on the C3 output: python tools/ess_perf.py [--nchains 65536]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import klara_b200 as K
ap = argparse.ArgumentParser(); ap.add_argument("--nchains", type=int, default=65536); a = ap.parse_args()
L = K._lib; lib = L.lib()
p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
job = K.BasicMCJob(K.likelihood_model(p, False), K.HMC(0.05, 10), K.BasicMCRange(nsteps=200, burnin=100),
                   {"p": K.SyntheticNormal(a.nchains, 1024)}, outopts={"monitor": ["value"]}, seed=20240925)
job.run()
job.ess(to_host=False)
for r in range(3):
    t = time.perf_counter(); job.ess(to_host=False); ms = (time.perf_counter() - t) * 1e3
    print("klb_job_ess: %.2f ms  (%.1f GB of samples -> %.0f GB/s)" % (ms, a.nchains * 100 * 1024 * 8 / 1e9, a.nchains * 100 * 1024 * 8 / ms / 1e6))
