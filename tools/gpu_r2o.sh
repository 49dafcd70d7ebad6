#!/bin/bash
# round-2 GPU call O (8 GPUs): final bench lines at N = 8, 4, 2 with the shipping library
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2o_bench_n8.json 2> gpurun_out/r2o_bench_n8.err
echo "n8 rc=$?"; tail -2 gpurun_out/r2o_bench_n8.err
timeout 600 $T --nproc-per-node 4 --master-port 29552 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2o_bench_n4.json 2> gpurun_out/r2o_bench_n4.err
echo "n4 rc=$?"
timeout 600 $T --nproc-per-node 2 --master-port 29553 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err
echo "n2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err
echo "n1 rc=$?"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q | tail -2
for n in 1 2 4 8; do grep '^{' gpurun_out/r2o_bench_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'value %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'full', d['e2e_full_output'].get('value'), 'parity', d['parity']['bit_exact'], d['parity'].get('g_invariance',{}).get('equal'), d['parity']['final_state_checksum'], d.get('details', d['config'])['closing_all_gather'][:12])
for k,v in d['configs'].items(): print('   ', k, '%.4g'%v['value'], '%.2f ms'%v['ms_per_run'], 'frac %.3f'%v['roofline']['frac'])
"; done
