#!/bin/bash
# ncu columns of the team kernel (d = 4096) and the shifted-target kernel (d = 1024); whole GPU suite; smoke; bench at N = 1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:klb_hmc_ws -s 1 -c 1 -o gpurun_out/r2r_team4096 python tools/prof_run.py --dim 4096 --nchains 9472 --nsteps 20 --burnin 10 --step 0.02 --reps 2 > gpurun_out/r2r_ncu1.log 2>&1
timeout 600 $NCU -k regex:klb_hmc_ws -s 1 -c 1 -o gpurun_out/r2r_shifted1024 python tools/prof_run.py --target shifted --dim 1024 --nchains 37888 --nsteps 20 --burnin 10 --step 0.02 --reps 2 > gpurun_out/r2r_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; tail -3 gpurun_out/r2r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2r_bench_ref.json 2> gpurun_out/r2r_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2r_bench_n1.json", "gpurun_out/r2r_bench_ref.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "value %.4g" % d["value"], "ms", d.get("ms_per_step"), "e2e %.4g" % d["e2e"]["value"], "parity", (d.get("parity") or {}).get("bit_exact"),
                  "launches", d.get("gpu_launches"), "roofline", (d.get("roofline") or {}).get("frac"))
            for k, c in (d.get("configs") or {}).items():
                print("    %s %.4g %.2f ms frac %.3f" % (k, c["value"], c["ms_per_run"], c["roofline"]["frac"]))
PY
