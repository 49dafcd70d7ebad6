#!/bin/bash
# round-2 GPU call N: run_host in two halves / over devices, full parity suite
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1
tail -4 gpurun_out/r2n_pytest.log
python tools/e2e_slices.py --nchains 8192 | tail -3
