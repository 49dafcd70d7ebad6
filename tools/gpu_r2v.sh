#!/bin/bash
# round-2 GPU call V (short): NUTS on the logit target with the :a / :na diagnostics, every NUTS test, the ESS kernel variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 40 python tools/nuts_glm_check.py 2>&1 | tail -4
timeout 90 python -m pytest tests -m gpu -q -x -k "nuts or NUTS" 2>&1 | tail -6
timeout 60 python tools/ess_variants.py 2>&1 | tee gpurun_out/r2v_ess_variants.log | tail -14
timeout 30 python tools/glm_perf.py NUTS 2>&1 | tail -2
