#!/bin/bash
# round-2 GPU call I (8 GPUs): multi-GPU layer on 8 devices, bench at N = 8 (both closing all-gathers) and N = 4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "real_devices or ipc" > gpurun_out/r2i_pytest.log 2>&1
tail -3 gpurun_out/r2i_pytest.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2i_bench_n8_p2p.json 2> gpurun_out/r2i_bench_n8_p2p.err
echo "n8 p2p rc=$?"; tail -2 gpurun_out/r2i_bench_n8_p2p.err; grep '^{' gpurun_out/r2i_bench_n8_p2p.json | cut -c1-400
timeout 600 $T --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3 --gather nccl --no-configs --no-e2e-full --no-cpu > gpurun_out/r2i_bench_n8_nccl.json 2> gpurun_out/r2i_bench_n8_nccl.err
echo "n8 nccl rc=$?"; grep '^{' gpurun_out/r2i_bench_n8_nccl.json | cut -c1-400
timeout 600 $T --nproc-per-node 4 --master-port 29543 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e-full --no-cpu > gpurun_out/r2i_bench_n4_p2p.json 2> gpurun_out/r2i_bench_n4_p2p.err
echo "n4 rc=$?"; grep '^{' gpurun_out/r2i_bench_n4_p2p.json | cut -c1-400
