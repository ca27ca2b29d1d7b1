#!/bin/bash
# 1-GPU validation of the Adam epilogue, the tcgen05 FFN GEMMs and the fused dropout+add+LayerNorm; BERT ablation
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_seventh.log) 2>&1
export DEAR_TIMEOUT_S=120
echo "=== new tests"; timeout 420 python -m pytest tests/test_tc_gemm.py tests/test_fused_ln.py tests/test_adam.py tests/test_kernels_direct.py tests/test_fused_bn.py -m gpu -q --timeout 200 2>&1 | tail -25
echo "=== op micro-benchmarks"; timeout 200 python tools/bert_ops_bench.py --json gpurun_out/bert_ops_bench.json 2>&1 | tail -45
B="timeout 240 python bench.py --model bert --steps 20 --warmup 8"
echo "=== bert eager-ops baseline";   $B --fused-ln 0 --tc-ffn 0 --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_ablate_none.json | cut -c1-200
echo "=== bert fused ln";             $B --fused-ln 1 --tc-ffn 0 --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_ablate_ln.json | cut -c1-200
echo "=== bert tc ffn";               $B --fused-ln 0 --tc-ffn 1 --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_ablate_tc.json | cut -c1-200
echo "=== bert fused ln + tc ffn";    $B --fused-ln 1 --tc-ffn 1 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_fused_1gpu.json | cut -c1-900
echo "=== bert fused, cublas down";   DEAR_TC_DOWN=0 $B --fused-ln 1 --tc-ffn 1 --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_ablate_tc_cublas_down.json | cut -c1-200
echo "=== bert adamw";                $B --optimizer adamw --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_adamw_1gpu.json | cut -c1-200
echo "=== default bench"; timeout 240 python bench.py 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_default_1gpu.json | cut -c1-600
echo "=== full gpu test-suite"; timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -8
echo "=== done"
