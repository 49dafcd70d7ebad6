#!/usr/bin/env python
"""Small configurations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
Each run is also compared with the oracle, so a sanitizer-clean run that computes garbage cannot pass."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import klara_b200 as K  # noqa: E402
from helpers import build_pair, compare_run  # noqa: E402

L = K._lib
CASES = [
    ("ws-hmc iso 1024 (warp-specialised producer/consumer, named barriers)", "HMC", "iso", dict(nchains=10, dim=1024, nsteps=5, burnin=2, step=0.05, nleaps=4)),
    ("ws-hmc iso 700 (masked)", "HMC", "iso", dict(nchains=6, dim=700, nsteps=4, burnin=1, step=0.05, nleaps=3)),
    ("ws-hmc dual averaging", "HMC", "iso", dict(nchains=6, dim=1024, nsteps=5, burnin=1, step=0.03, nleaps=3, tuner="dualavg", nadapt=3)),
    ("ws-hmc 4 consumer warps per chain 2048 (team barrier, double-buffered exchange)", "HMC", "iso", dict(nchains=3, dim=2048, nsteps=5, burnin=1, step=0.03, nleaps=3)),
    ("ws-hmc 4 consumer warps per chain 3000 rosen, dual averaging", "HMC", "rosen", dict(nchains=3, dim=3000, nsteps=5, burnin=1, step=0.004, nleaps=3, tuner="dualavg", nadapt=3)),
    ("fused hmc 64", "HMC", "shifted", dict(nchains=9, dim=64, nsteps=6, burnin=2, step=0.1, nleaps=3)),
    ("mala rosen 256 + tuner", "MALA", "rosen", dict(nchains=9, dim=256, nsteps=12, burnin=8, step=0.01, tuner="accrate", period=4)),
    ("mh iso 1024", "MH", "iso", dict(nchains=5, dim=1024, nsteps=6, burnin=2, sigma=np.full(1024, 0.02))),
    ("dense DMMA clusters (TMA multicast, mbarriers) 128", "HMC", "dense", dict(nchains=40, dim=128, nsteps=3, burnin=1, step=0.05, nleaps=3)),
    ("dense DMMA 512", "HMC", "dense", dict(nchains=20, dim=512, nsteps=2, burnin=0, step=0.02, nleaps=2)),
    ("dense DFMA tile mala 30", "MALA", "dense", dict(nchains=11, dim=30, nsteps=4, burnin=1, step=0.02)),
    ("glm hmc", "HMC", "logit", dict(nchains=70, dim=4, nsteps=4, burnin=1, step=0.02, nleaps=3)),
]
CASES += [
    ("nuts iso 2048, four warps per chain, dual averaging (team barrier between the reads and the writer's stores)", "NUTS", "iso",
     dict(nchains=5, dim=2048, nsteps=6, burnin=2, step=0.1, tuner="dualavg", nadapt=4, period=2, verbose=True,
          diagnostics=("accept", "ndoublings"), maxndoublings=3)),
    ("nuts shifted 100", "NUTS", "shifted", dict(nchains=5, dim=100, nsteps=6, burnin=2, step=0.2, diagnostics=("accept", "ndoublings"), maxndoublings=4)),
    ("nuts rosen 1000, dual averaging", "NUTS", "rosen", dict(nchains=5, dim=1000, nsteps=6, burnin=2, step=0.02, tuner="dualavg", nadapt=4,
                                                         diagnostics=("accept", "ndoublings"), maxndoublings=3)),
]
CASES += [
    ("glm nuts (thread per chain), dual averaging, :a / :na", "NUTS", "logit",
     dict(nchains=70, dim=4, nsteps=5, burnin=1, step=0.15, tuner="dualavg", nadapt=3, diagnostics=("accept", "ndoublings", "a", "na"), maxndoublings=3)),
    ("glm nuts dim 13 (padded to 16), trees that stop early", "NUTS", "logit",
     dict(nchains=40, dim=13, nsteps=4, burnin=1, step=0.3, diagnostics=("accept", "ndoublings"), maxndoublings=3, maxdelta=3)),
]
ONLY = os.environ.get("KLB_SANITIZE_ONLY", "")          # substring filter, e.g. KLB_SANITIZE_ONLY=nuts
if ONLY:
    CASES = [c for c in CASES if ONLY in c[0]]
for name, smp, tgt, kw in CASES:
    job, cfg, x0, tp, sg = build_pair(K, smp, tgt, seed=31, **kw)
    compare_run(job, cfg, x0, tp, sg)
    print("ok:", name, flush=True)
# pipelined host-to-host call: slice streams, async copies
job, cfg, x0, tp, sg = build_pair(K, "HMC", "iso", nchains=50, dim=1024, nsteps=4, burnin=1, step=0.05, nleaps=3, seed=5)
val = np.empty((50, 3, 1024))
job.run_host(x0, {L.OUT_VALUE: val}, 4)
ref, *_ = build_pair(K, "HMC", "iso", nchains=50, dim=1024, nsteps=4, burnin=1, step=0.05, nleaps=3, seed=5)
ref.run()
assert np.array_equal(val, ref.output().value)
print("ok: klb_job_run_host, 4 slices", flush=True)
# post-hoc statistics kernels: the shipping one-warp-per-CTA window kernel and the variants, series lengths that end inside a trip
from oracle import oracle as O  # noqa: E402
for npost, dim in ((3, 1024), (37, 130), (100, 33)):
    sj, *_ = build_pair(K, "HMC", "iso", nchains=7, dim=dim, nsteps=npost + 2, burnin=2, step=0.1, nleaps=3, seed=8, monitor=("value",))
    sj.run()
    want = O.stats(sj.output().value)
    for variant in ("", "0", "2", "6", "9", "10"):
        if variant:
            os.environ["KLB_ESS_VARIANT"] = variant
        else:
            os.environ.pop("KLB_ESS_VARIANT", None)
        for got, key in ((sj.ess(), "ess"), (sj.mean(), "mean"), (sj.mcvar("imse"), "mcvar_imse")):
            assert np.array_equal(got.view(np.uint64), np.ascontiguousarray(want[key]).view(np.uint64)), (npost, dim, variant, key)
    os.environ.pop("KLB_ESS_VARIANT", None)
    sj.acceptance()
    sj.close()
print("ok: ess / stats / acceptance kernels (window kernel and variants)", flush=True)
