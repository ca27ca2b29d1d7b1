#!/bin/bash
# 2 GPUs: correctness of the new Kernel A variants + kernel micro-benchmark (one-shot vs pipelined vs NCCL).
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_kernelA_2gpu.log) 2>&1
export DEAR_TIMEOUT_S=180
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== tests: kernels direct (one-shot, pipelined, NVLS), engine"
timeout 600 python -m pytest tests/test_kernels_direct.py tests/test_gpu_fused.py -m gpu -x -q --timeout 280 2>&1 | tail -15
echo "=== kernel bench P=2 auto (pipe >= 2 MB)"
timeout 300 $TR --master-port 29801 tools/kernel_bench.py --sizes-mb 1,4,24,64,392 --out gpurun_out/kernel_bench_p2_r2_auto.json 2>&1 | grep -E '^\{|rror'
echo "=== kernel bench P=2 one-shot forced"
DEAR_RS_ALGO=oneshot timeout 300 $TR --master-port 29802 tools/kernel_bench.py --sizes-mb 4,24,64,392 --nccl 0 --out gpurun_out/kernel_bench_p2_r2_oneshot.json 2>&1 | grep -E '^\{|rror'
echo "=== kernel bench P=2 pipe grid 48 / 64"
DEAR_RS_GRID=48 timeout 300 $TR --master-port 29803 tools/kernel_bench.py --sizes-mb 24,64,392 --nccl 0 2>&1 | grep -E '^\{|rror'
DEAR_RS_GRID=64 timeout 300 $TR --master-port 29804 tools/kernel_bench.py --sizes-mb 24,64,392 --nccl 0 2>&1 | grep -E '^\{|rror'
echo "=== kernel bench P=2 pipe stripe 4 MB / 16 MB"
DEAR_STRIPE_MB=4 timeout 300 $TR --master-port 29805 tools/kernel_bench.py --sizes-mb 24,64,392 --nccl 0 2>&1 | grep -E '^\{|rror'
DEAR_STRIPE_MB=16 timeout 300 $TR --master-port 29806 tools/kernel_bench.py --sizes-mb 24,64,392 --nccl 0 2>&1 | grep -E '^\{|rror'
echo "=== done"
