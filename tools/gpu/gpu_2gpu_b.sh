#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_2gpu_b.log) 2>&1
export DEAR_TIMEOUT_S=120
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== pytest fused bn"; timeout 600 python -m pytest tests/test_fused_bn.py -m gpu -q --timeout 300 2>&1 | tail -6
echo "=== bench dear 2 GPUs fp32 fused graph"; timeout 300 $TR --master-port 29901 bench.py --gpus 2 --steps 30 --warmup 10 --graph 1 2>&1 | grep -E '"metric"|Error' | tee gpurun_out/bench_dear_2gpu_fused_graph.json | cut -c1-330
echo "=== bench dear 2 GPUs fp32 fused eager"; timeout 300 $TR --master-port 29902 bench.py --gpus 2 --steps 30 --warmup 10 --no-e2e 2>&1 | grep -E '"metric"|Error' | cut -c1-330
echo "=== bench dear 2 GPUs BERT bf16 graph"; timeout 300 $TR --master-port 29903 bench.py --gpus 2 --model bert --steps 15 --warmup 6 --graph 1 --no-e2e 2>&1 | grep -E '"metric"|Error' | cut -c1-330
echo "=== done"
