#!/bin/bash
# team (W = 4) warp-specialised HMC kernel: parity, sanitizers, d = 4096 / 2048 before and after
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "team_kernel or hmc_iso_bit_exact or dual_averaging" 2>&1 | tail -5
for ws in 1 ""; do
  echo "== KLB_HMC_WS='$ws' (1 = fused 4-warp-team kernel, empty = warp-specialised team kernel)"
  for dim in 4096 2048 1536; do
    KLB_HMC_WS=$ws timeout 300 python tools/prof_run.py --dim $dim --nchains 18944 --nsteps 40 --burnin 20 --step 0.02 --reps 3 | tail -2
    KLB_HMC_WS=$ws timeout 300 python tools/prof_run.py --dim $dim --nchains 18944 --nsteps 40 --burnin 20 --step 0.02 --reps 3 --arith fma | tail -2 | head -1
  done
done
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_run.py > gpurun_out/r2p_racecheck.log 2>&1; tail -4 gpurun_out/r2p_racecheck.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r2p_memcheck.log 2>&1; tail -3 gpurun_out/r2p_memcheck.log
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_run.py > gpurun_out/r2p_synccheck.log 2>&1; tail -3 gpurun_out/r2p_synccheck.log
