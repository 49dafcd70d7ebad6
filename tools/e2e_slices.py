#!/usr/bin/env python
"""End-to-end time of klb_job_run_host for one GPU's shard of C3 as a function of the slice count.
    python tools/e2e_slices.py --nchains 8192       # the shard of an 8-GPU run
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import klara_b200 as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nchains", type=int, default=8192)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
L = K._lib
lib = L.lib()
N, d, P = a.nchains, 1024, 100


def pinned(shape, dtype=np.float64):
    nb = int(np.prod(shape)) * np.dtype(dtype).itemsize
    h = C.c_void_p()
    L.check(lib.klb_host_alloc(C.byref(h), nb))
    ct = {np.float64: C.c_double, np.uint8: C.c_uint8}[dtype]
    return np.ctypeslib.as_array(C.cast(h, C.POINTER(ct)), shape=shape)


x0 = pinned((N, d))
x0[:] = np.random.default_rng(0).standard_normal((N, d))
st, lt, ac = pinned((N, d)), pinned((N, P)), pinned((N, P), np.uint8)
p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
job = K.BasicMCJob(K.likelihood_model(p, False), K.HMC(0.05, 10), K.BasicMCRange(nsteps=200, burnin=100), {"p": x0},
                   outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept"]}, seed=1)
job.run()
print("device-resident run: %.3f ms" % job.last_run_ms)
for ns in (1, 2, 4, 8, 16, 0):
    bufs = {L.OUT_STATE: st, L.OUT_LOGTARGET: lt, L.OUT_ACCEPT: ac}
    job.run_host(x0, bufs, ns)
    t = time.perf_counter()
    for _ in range(a.reps):
        job.run_host(x0, bufs, ns)
    ms = (time.perf_counter() - t) / a.reps * 1e3
    print("nslices %2d: %.3f ms per run_host  (%.3g leapfrog-steps/s)" % (ns, ms, N * 2000 / ms * 1e3))
