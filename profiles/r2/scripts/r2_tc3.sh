#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_tc3.log) 2>&1
echo "=== tcgen05 tests"; timeout 200 python -m pytest tests/test_tc_gemm.py -m gpu -q --timeout 100 -x 2>&1 | tail -4
echo "=== FFN op micro-benchmark (direct epilogue, multicast clusters)"; timeout 200 python tools/bert_ops_bench.py --sections gemm,ffn --json gpurun_out/bert_ops_bench_r2_final.json 2>&1 | grep -E "handwritten|eager|cublas|failed|rror"
echo "=== done"
