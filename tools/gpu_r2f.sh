#!/bin/bash
# round-2 GPU call F: full parity suite on the final kernels, secondary-kernel timings and ncu columns
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
tail -4 gpurun_out/r2f_pytest.log
{ python tools/glm_perf.py; python tools/ess_perf.py; } > gpurun_out/r2f_timings.txt 2>&1
cat gpurun_out/r2f_timings.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:klb_glm_kernel -c 1 -o gpurun_out/r2f_prof_glm python tools/glm_perf.py HMC > gpurun_out/r2f_prof_glm.log 2>&1
timeout 600 $NCU -k regex:klb_ess_tile -s 1 -c 1 -o gpurun_out/r2f_prof_ess python tools/ess_perf.py --nchains 16384 > gpurun_out/r2f_prof_ess.log 2>&1
timeout 600 $NCU -k regex:klb_chain_kernel -s 1 -c 1 -o gpurun_out/r2f_prof_mala_c5 python tools/prof_run.py --sampler MALA --target rosen --dim 256 --nchains 32768 --nsteps 100 --burnin 50 --step 0.01 --accrate 0.574 --reps 2 > gpurun_out/r2f_prof_mala.log 2>&1
timeout 600 $NCU -k regex:klb_chain_kernel -s 1 -c 1 -o gpurun_out/r2f_prof_mh python tools/prof_run.py --sampler MH --nchains 65536 --nsteps 20 --burnin 10 --reps 2 > gpurun_out/r2f_prof_mh.log 2>&1
timeout 600 $NCU -k regex:klb_dense_mma -s 1 -c 1 -o gpurun_out/r2f_prof_dense python tools/prof_run.py --target dense --nchains 16384 --dim 512 --nsteps 4 --burnin 2 --step 0.02 --nleaps 20 --reps 2 > gpurun_out/r2f_prof_dense.log 2>&1
timeout 600 $NCU -k regex:klb_chain_kernel -s 1 -c 1 -o gpurun_out/r2f_prof_hmc4096 python tools/prof_run.py --dim 4096 --nchains 8192 --nsteps 10 --burnin 5 --step 0.02 --reps 2 > gpurun_out/r2f_prof_hmc4096.log 2>&1
tail -2 gpurun_out/r2f_prof_*.log
python tools/prof_run.py --dim 4096 --nchains 16384 --nsteps 40 --burnin 20 --step 0.02 --reps 2
ls -la gpurun_out | grep r2f
