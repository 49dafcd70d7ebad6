#!/bin/bash
# round-2 GPU call C: Philox4x32-7 contract + scheduling post-pass: parity suite, timings, sanitizers
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1
tail -5 gpurun_out/r2c_pytest.log
{
python tools/prof_run.py --nchains 65536 --nsteps 40 --burnin 20 --reps 3
python tools/prof_run.py --nchains 65536 --nsteps 40 --burnin 20 --reps 2 --arith fma
python tools/prof_run.py --sampler MALA --target rosen --dim 256 --nchains 32768 --nsteps 2000 --burnin 1000 --step 0.01 --accrate 0.574 --reps 2
python tools/prof_run.py --sampler MALA --dim 128 --nchains 4096 --nsteps 2000 --burnin 1000 --step 0.9 --reps 2
python tools/prof_run.py --sampler MH --nchains 65536 --nsteps 200 --burnin 100 --reps 2
python tools/prof_run.py --sampler HMC --dim 512 --nchains 65536 --nsteps 40 --burnin 20 --reps 2
} > gpurun_out/r2c_timings.txt 2>&1
cat gpurun_out/r2c_timings.txt
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/r2c_sanitizer_$tool.log python tools/sanitize_run.py > gpurun_out/r2c_sanitizer_$tool.out 2>&1
  echo "== $tool rc=$?"; tail -3 gpurun_out/r2c_sanitizer_$tool.out; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/r2c_sanitizer_$tool.log | tail -5
done
