#!/bin/bash
# rotated CUDA-graph body (update + all-gather overlap the forward inside the graph): correctness + gain
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_tenth.log) 2>&1
export DEAR_TIMEOUT_S=120
echo "=== graph tests"; timeout 240 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 200 -k "graph" 2>&1 | tail -6
B="timeout 200 python bench.py --model bert --steps 20 --warmup 8"
echo "=== bert natural graph body"; $B --overlap-update 0 --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_overlap0.json | cut -c1-200
echo "=== bert rotated graph body";  $B --overlap-update 1 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_overlap1.json | cut -c1-1100
echo "=== resnet50 rotated graph body"; timeout 200 python bench.py --overlap-update 1 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_default_overlap1.json | cut -c1-500
echo "=== engine tests"; timeout 200 python -m pytest tests/test_gpu_fused.py tests/test_adam.py -m gpu -q --timeout 150 -k "not graph" 2>&1 | tail -5
echo "=== done"
