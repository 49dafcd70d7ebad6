#!/bin/bash
# round-2 GPU call D: multi-GPU layer on one device, the new bench line, racecheck re-run, ncu evidence
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -k "multi or gather or synthetic or peaks or run_host or warp_specialised" -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1
tail -15 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2d_bench.err; cut -c1-3000 gpurun_out/r2d_bench.json
timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/r2d_racecheck.log python tools/sanitize_run.py > gpurun_out/r2d_racecheck.out 2>&1
grep -E "RACECHECK SUMMARY" gpurun_out/r2d_racecheck.log
KLB_DENSE_CLUSTER=1 timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/r2d_racecheck_nocluster.log python tools/sanitize_run.py > gpurun_out/r2d_racecheck_nocluster.out 2>&1
grep -E "RACECHECK SUMMARY" gpurun_out/r2d_racecheck_nocluster.log
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches_bench_steps2_warmup1.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e-full > gpurun_out/r2d_bench_under_ncu.log 2>&1
# the headline kernel: the bench launch (65 536 chains x 200 transitions), full set, once
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:klb_hmc_ws -s 1 -c 1 -o gpurun_out/r2d_prof_ws \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e-full --no-configs > gpurun_out/r2d_prof_ws.log 2>&1
ls -la gpurun_out/ | tail -12
