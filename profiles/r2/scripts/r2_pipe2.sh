#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_pipe2.log) 2>&1
export DEAR_TIMEOUT_S=180
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== tests"
timeout 600 python -m pytest tests/test_kernels_direct.py tests/test_gpu_fused.py -m gpu -x -q --timeout 280 2>&1 | tail -8
for g in 32 48 64; do
echo "=== kernel bench P=2 pipe grid $g"
DEAR_RS_GRID=$g timeout 300 $TR --master-port 2970$((g/16)) tools/kernel_bench.py --sizes-mb 4,24,64,392 --nccl 0 2>&1 | grep -E '^\{|rror' | cut -c1-250
done
echo "=== done"
