#!/bin/bash
# NUTS kernels: parity against the oracle, then the whole GPU suite, sanitizers on a small NUTS case
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nuts" 2>&1 | tail -25
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; tail -3 gpurun_out/r2s_pytest.log
cat > /tmp/nuts_san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import klara_b200 as K
from helpers import build_pair, compare_run
for tgt, dim, tuner in (("iso", 2048, "dualavg"), ("shifted", 100, "vanilla"), ("rosen", 1000, "dualavg")):
    job, cfg, x0, tp, sg = build_pair(K, "NUTS", tgt, nchains=5, dim=dim, nsteps=6, burnin=2, step=0.02 if tgt == "rosen" else 0.1, seed=3,
                                      tuner=tuner, nadapt=4, period=2, verbose=True, diagnostics=("accept", "ndoublings"), maxndoublings=3)
    compare_run(job, cfg, x0, tp, sg)
    print("ok: nuts", tgt, dim, tuner, flush=True)
PY
for tool in racecheck memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python /tmp/nuts_san.py > gpurun_out/r2s_nuts_$tool.log 2>&1; grep "^ok\|SUMMARY" gpurun_out/r2s_nuts_$tool.log
done
echo "== NUTS throughput (iso, d = 1024, 18944 chains, 40 transitions, maxndoublings 5 = 31 leapfrog steps per transition)"
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import klara_b200 as K
for dim, n in ((1024, 18944), (256, 75776), (4096, 4736)):
    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    job = K.BasicMCJob(K.likelihood_model(p, False), K.NUTS(0.05), K.BasicMCRange(nsteps=40, burnin=20), {"p": K.SyntheticNormal(n, dim)},
                       outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept", "ndoublings"]}, seed=1)
    job.run(); job.reset(); job.run()
    o = job.output()
    nd = o.diagnosticvalues[:, 1, :].astype(np.int64)
    leaps = float((2 ** nd - 1).mean()) * 40 * n          # the stored half stands for the whole run
    print(dim, n, "%.2f ms" % job.last_run_ms, "leapfrog-steps/s %.3e" % (leaps / job.last_run_ms * 1e3), "accept %.3f" % o.diagnosticvalues[:, 0, :].mean(),
          "regs", job.plan().regs_per_thread, flush=True)
PY
