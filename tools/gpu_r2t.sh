#!/bin/bash
# NUTS: parity, whole GPU suite, sanitizers on every kernel family (NUTS included)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nuts" 2>&1 | tail -8
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; tail -3 gpurun_out/r2t_pytest.log
for tool in racecheck memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/r2t_$tool.log 2>&1; grep -c "^ok" gpurun_out/r2t_$tool.log; grep "SUMMARY" gpurun_out/r2t_$tool.log
done
