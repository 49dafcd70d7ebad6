#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1
tail -4 gpurun_out/r2n_pytest.log
