#!/bin/bash
# round-2 GPU call K: RNG software-pipelined into the MALA / MH passes (A/B), ESS main-loop rewrite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "mala or mh or ess or stats or random or other_targets or run_host or multi or twin or c2 or c5 or golden or readme" > gpurun_out/r2k_pytest.log 2>&1
tail -3 gpurun_out/r2k_pytest.log
{
for v in _nopipe ""; do
  echo "== variant '$v'"
  export KLB_LIB_PATH=$PWD/klara.jl_b200/lib/libklara_b200$v.so
  python tools/prof_run.py --sampler MALA --target rosen --dim 256 --nchains 32768 --nsteps 2000 --burnin 1000 --step 0.01 --accrate 0.574 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MALA --dim 128 --nchains 4096 --nsteps 2000 --burnin 1000 --step 0.9 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MH --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MALA --step 0.02 --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MH --dim 256 --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
done
unset KLB_LIB_PATH
python tools/ess_perf.py
} > gpurun_out/r2k_timings.txt 2>&1
cat gpurun_out/r2k_timings.txt
