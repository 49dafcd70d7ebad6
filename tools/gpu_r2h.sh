#!/bin/bash
# round-2 GPU call H: parity suite on the final kernels (shared-divisor division, occupancy targets), MALA / MH timings
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1
tail -4 gpurun_out/r2h_pytest.log
{
for v in "" _nhs5 _nhs6; do
  echo "== variant '$v'"
  export KLB_LIB_PATH=$PWD/klara.jl_b200/lib/libklara_b200$v.so
  python tools/prof_run.py --sampler MALA --target rosen --dim 256 --nchains 32768 --nsteps 2000 --burnin 1000 --step 0.01 --accrate 0.574 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MALA --dim 128 --nchains 4096 --nsteps 2000 --burnin 1000 --step 0.9 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MH --dim 256 --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
done
unset KLB_LIB_PATH
python tools/prof_run.py --sampler MH --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
python tools/prof_run.py --sampler MALA --step 0.02 --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
python tools/prof_run.py --sampler MALA --step 0.03 --dim 512 --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
python tools/prof_run.py --sampler MALA --target dense --dim 512 --step 0.002 --nchains 16384 --nsteps 40 --burnin 20 --reps 2 | tail -2 | head -1
python tools/glm_perf.py MALA
} > gpurun_out/r2h_timings.txt 2>&1
cat gpurun_out/r2h_timings.txt
