#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_oneshot_grid.log) 2>&1
export DEAR_TIMEOUT_S=180
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for g in 48 64 96 128; do
echo "=== kernel bench P=2 one-shot grid $g"
DEAR_RS_ALGO=oneshot DEAR_RS_GRID=$g DEAR_AG_GRID=$g timeout 300 $TR --master-port 298$g tools/kernel_bench.py --sizes-mb 4,24,64,392 --nccl 0 2>&1 | grep -E '^\{|rror' | cut -c1-330
done
echo "=== done"
