#!/bin/bash
# round-2 GPU call J: final regression -- parity suite, smoke(), the bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1
tail -4 gpurun_out/r2j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; tail -2 gpurun_out/r2j_smoke.log
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2j_bench.err; grep '^{' gpurun_out/r2j_bench.json | cut -c1-1200
