#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_bo.log) 2>&1
export DEAR_TIMEOUT_S=120
echo "=== pytest rebucket on gpu"; timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 300 -k "rebucket" 2>&1 | tail -4
echo "=== BERT-base dear-bo 2 GPUs"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29931 benchmarks/bert_benchmark.py --model bert_base --batch-size 64 --sentence-len 64 --dtype bf16 --method dear-bo --num-warmup-batches 60 --num-iters 3 --num-batches-per-iter 10 2>&1 | grep -E "BO Tuning|Total|Iter #|Error|error|Tensor fusion" | tail -22
echo "=== BERT-base dear (25 MB) 2 GPUs"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29932 benchmarks/bert_benchmark.py --model bert_base --batch-size 64 --sentence-len 64 --dtype bf16 --method dear --num-warmup-batches 10 --num-iters 3 --num-batches-per-iter 10 2>&1 | grep -E "Total|Error|error" | tail -3
echo "=== done"
