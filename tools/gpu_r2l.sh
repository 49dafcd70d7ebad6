#!/bin/bash
# round-2 GPU call L: dense MMA kernel with the light stage-release arrive, MALA kernel without the stack copy, parity suite
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1
tail -3 gpurun_out/r2l_pytest.log
{
python tools/prof_run.py --sampler MALA --target rosen --dim 256 --nchains 32768 --nsteps 2000 --burnin 1000 --step 0.01 --accrate 0.574 --reps 2 | tail -2 | head -1
python tools/prof_run.py --target dense --nchains 16384 --dim 512 --nsteps 20 --burnin 10 --step 0.02 --nleaps 20 --reps 3
python tools/prof_run.py --target dense --nchains 16384 --dim 256 --nsteps 20 --burnin 10 --step 0.02 --nleaps 20 --reps 2 | tail -2
python tools/prof_run.py --target dense --nchains 16384 --dim 128 --nsteps 20 --burnin 10 --step 0.02 --nleaps 20 --reps 2 | tail -2
python tools/prof_run.py --nchains 65536 --nsteps 40 --burnin 20 --reps 2 | tail -2 | head -1
} > gpurun_out/r2l_timings.txt 2>&1
cat gpurun_out/r2l_timings.txt
timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/r2l_racecheck.log python tools/sanitize_run.py > gpurun_out/r2l_racecheck.out 2>&1
grep -E "RACECHECK SUMMARY" gpurun_out/r2l_racecheck.log; grep -E "Error:|Warning:" gpurun_out/r2l_racecheck.log | sed -E 's/\+0x[0-9a-f]+//g' | sort | uniq -c | head
timeout 600 ncu --set full --clock-control none --import-source on -k regex:klb_dense_mma -s 1 -c 1 -o gpurun_out/r2l_prof_dense python tools/prof_run.py --target dense --nchains 16384 --dim 512 --nsteps 4 --burnin 2 --step 0.02 --nleaps 20 --reps 2 > gpurun_out/r2l_prof_dense.log 2>&1
