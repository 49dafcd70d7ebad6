#!/bin/bash
# parity of the whole suite file (odd dense dims, dual averaging on dense, team kernel) + shifted-target leapfrog blocking
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
for lib in klara.jl_b200/lib/libklara_b200_blk0.so ""; do
  echo "== KLB_LIB_PATH='$lib' (blk0 = generic leapfrog loop for the shifted target, empty = blocked step)"
  for dim in 1024 4096; do
    n=$((148*128*1024*4/dim))
    KLB_LIB_PATH=$lib timeout 300 python tools/prof_run.py --target shifted --dim $dim --nchains $n --nsteps 40 --burnin 20 --step 0.02 --reps 3 | tail -2
    KLB_LIB_PATH=$lib timeout 300 python tools/prof_run.py --target shifted --dim $dim --nchains $n --nsteps 40 --burnin 20 --step 0.02 --reps 3 --arith fma | tail -2 | head -1
  done
done
echo "== dense tile kernel, dual averaging vs vanilla (HMC, d = 128 | 129)"
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
import klara_b200 as K
for dim in (128, 129):
    idx = np.arange(dim); C = np.linalg.inv(0.8 ** np.abs(idx[:, None] - idx[None, :])); C = (C + C.T) / 2
    x0 = np.random.default_rng(0).standard_normal((148 * 8 * 8, dim)) * 0.3
    for tuner in (K.VanillaMCTuner(), K.DualAveragingMCTuner(0.65, 100)):
        p = K.BasicContMuvParameter("p", logtarget=K.DenseGaussian(C))
        job = K.BasicMCJob(K.likelihood_model(p, False), K.HMC(0.05, 10), K.BasicMCRange(nsteps=200, burnin=100), {"p": x0},
                           tuner=tuner, outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept"]}, seed=1)
        job.run()
        print(dim, type(tuner).__name__, "%.2f ms" % job.last_run_ms, "accept %.3f" % job.output().diagnosticvalues.mean(),
              "regs", job.plan().regs_per_thread, flush=True)
PY
