#!/bin/bash
# round-2 GPU call A: scheduling-control experiments + regression tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
tools/_bin/fp64_batch_test > gpurun_out/r2a_batch.txt 2>&1
for c in fp64_yield fp64_yield_y fp64_yield_s2 fp64_yield_ys2; do tools/_bin/fp64_yield_run tools/_bin/$c.cubin; done > gpurun_out/r2a_yield.txt 2>&1
for v in "" _y _s2 _ys2; do
  echo "== variant '$v'"
  KLB_LIB_PATH=$PWD/klara.jl_b200/lib/libklara_b200$v.so python tools/prof_run.py --nchains 65536 --nsteps 40 --burnin 20 --reps 3
  KLB_LIB_PATH=$PWD/klara.jl_b200/lib/libklara_b200$v.so python tools/prof_run.py --nchains 65536 --nsteps 40 --burnin 20 --reps 3 --arith fma
done > gpurun_out/r2a_variants.txt 2>&1
python tools/e2e_slices.py --nchains 8192 > gpurun_out/r2a_slices.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
tail -3 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_yield.txt gpurun_out/r2a_variants.txt gpurun_out/r2a_slices.txt
