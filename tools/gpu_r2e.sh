#!/bin/bash
# round-2 GPU call E (2 GPUs): multi-GPU layer over real NVLink, bench at N = 2 with both closing all-gathers
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2e_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
tail -5 gpurun_out/r2e_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $T bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e-full > gpurun_out/r2e_bench_n2_p2p.json 2> gpurun_out/r2e_bench_n2_p2p.err
echo "p2p rc=$?"; tail -3 gpurun_out/r2e_bench_n2_p2p.err; cut -c1-1500 gpurun_out/r2e_bench_n2_p2p.json
timeout 600 $T bench.py --gpus 2 --steps 10 --warmup 3 --gather nccl --no-configs --no-e2e-full --no-cpu > gpurun_out/r2e_bench_n2_nccl.json 2> gpurun_out/r2e_bench_n2_nccl.err
echo "nccl rc=$?"; tail -3 gpurun_out/r2e_bench_n2_nccl.err; cut -c1-600 gpurun_out/r2e_bench_n2_nccl.json
