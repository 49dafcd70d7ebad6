#!/bin/bash
# round-2 GPU call W (the last one): whole GPU suite on the final tree, ESS variants, memcheck of the kernels added last,
# ncu captures of the ESS window kernel and the NUTS kernel (as far as the remaining budget goes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?"; tail -14 gpurun_out/r2w_pytest.log
timeout 40 python tools/ess_variants.py 2>&1 | tee gpurun_out/r2w_ess_variants.log | tail -9
KLB_SANITIZE_ONLY="glm nuts" timeout 70 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r2w_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -c "^ok" gpurun_out/r2w_memcheck.log; grep "SUMMARY" gpurun_out/r2w_memcheck.log
timeout 60 ncu --set full --clock-control none --import-source on -k regex:klb_ess_win -c 1 -f -o gpurun_out/r2w_ess python tools/ess_perf.py > gpurun_out/r2w_ncu_ess.log 2>&1
echo "ncu ess rc=$?"; ls -la gpurun_out/r2w_ess.ncu-rep 2>/dev/null
timeout 60 ncu --set full --clock-control none --import-source on -k regex:klb_nuts_kernel -s 1 -c 1 -f -o gpurun_out/r2w_nuts python tools/prof_run.py --sampler NUTS --nchains 18944 --nsteps 10 --burnin 5 --reps 2 > gpurun_out/r2w_ncu_nuts.log 2>&1
echo "ncu nuts rc=$?"; ls -la gpurun_out/r2w_nuts.ncu-rep 2>/dev/null
