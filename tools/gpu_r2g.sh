#!/bin/bash
# round-2 GPU call G: occupancy targets of the MALA / MH kernels at NV = 8 / 16, GLM two-row kernel, ESS re-timing
mkdir -p gpurun_out
{
for v in "" _nh2 _nh3 _nh4; do
  echo "== variant '$v'"
  export KLB_LIB_PATH=$PWD/klara.jl_b200/lib/libklara_b200$v.so
  python tools/prof_run.py --sampler MH --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MH --dim 512 --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MALA --step 0.02 --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MALA --step 0.03 --dim 512 --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
done
unset KLB_LIB_PATH
python tools/glm_perf.py
python tools/ess_perf.py
} > gpurun_out/r2g_timings.txt 2>&1
cat gpurun_out/r2g_timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:klb_glm_kernel -c 1 -o gpurun_out/r2g_prof_glm python tools/glm_perf.py HMC > gpurun_out/r2g_prof_glm.log 2>&1
