#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_tc2.log) 2>&1
export DEAR_TIMEOUT_S=180
echo "=== tcgen05 tests"; timeout 300 python -m pytest tests/test_tc_gemm.py -m gpu -q --timeout 120 -x 2>&1 | tail -12
echo "=== FFN op micro-benchmark"; timeout 300 python tools/bert_ops_bench.py --sections gemm --json gpurun_out/bert_ops_bench_r2_clusters.json 2>&1 | grep -E "handwritten|eager|cublas|failed|rror"
echo "=== graph + LR scheduler tests (capture-private tables)"; timeout 400 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 280 -x -k "graph or scheduler" 2>&1 | tail -5
echo "=== bert 1 GPU default / hand-written FFN"
B="timeout 300 python bench.py --no-e2e"
$B --model bert --steps 30 --warmup 8 2>&1 | grep -E '"metric"|rror' | cut -c1-200
$B --model bert --steps 30 --warmup 8 --tc-ffn 1 2>&1 | grep -E '"metric"|rror' | cut -c1-200
echo "=== resnet50 1 GPU e2e"
timeout 300 python bench.py --steps 40 --warmup 8 2>&1 | grep -E '"metric"|rror' | cut -c1-1000
echo "=== done"
