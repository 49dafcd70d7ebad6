#!/bin/bash
# round-2 GPU call U (last one of the round, ~6 min of box time): what the driver runs at round end, on the final tree --
# bench.py with its defaults (the NUTS entry of `configs` included), smoke(), the whole GPU suite with durations.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python bench.py > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2u_bench_n1.err
grep '^{' gpurun_out/r2u_bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'full', (d['e2e_full_output'] or {}).get('value'), 'frac %.3f'%d['roofline']['frac'], 'parity', d['parity']['bit_exact'], d['parity']['final_state_checksum'], 'cpu', d['cpu_baseline'] and '%.4g'%d['cpu_baseline']['value'], 'launches', d['gpu_launches'], d['clocks'])
for k,v in (d['configs'] or {}).items(): print('   ', k, v.get('error') or ('%.4g %.2f ms frac %.3f acc %.3f'%(v['value'], v['ms_per_run'], v['roofline']['frac'], v['accept_rate'])))
"
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 330 python -m pytest tests -m gpu -q --durations=25 > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?"; tail -32 gpurun_out/r2u_pytest.log
