#!/bin/bash
# 1 GPU: hand-written tcgen05 FFN kernels after the epilogue rework (+ MN-major weight operand), VGG-16 / BERT benches.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_1gpu_b.log) 2>&1
export DEAR_TIMEOUT_S=180
echo "=== tcgen05 tests"; timeout 300 python -m pytest tests/test_tc_gemm.py -m gpu -q --timeout 120 2>&1 | tail -12
echo "=== FFN op micro-benchmark"; timeout 300 python tools/bert_ops_bench.py --sections gemm --json gpurun_out/bert_ops_bench_r2.json 2>&1 | grep -E "handwritten|eager|cublas|failed|rror"
B="timeout 300 python bench.py --no-e2e"
echo "=== bert 1 GPU default / hand-written FFN"
$B --model bert --steps 30 --warmup 8 2>&1 | grep -E '"metric"|rror' | cut -c1-200
$B --model bert --steps 30 --warmup 8 --tc-ffn 1 2>&1 | grep -E '"metric"|rror' | cut -c1-200
echo "=== bert 1 GPU without direct wgrad"
DEAR_DIRECT_WGRAD=0 $B --model bert --steps 30 --warmup 8 2>&1 | grep -E '"metric"|rror' | cut -c1-200
echo "=== vgg16 1 GPU graph"
timeout 300 python bench.py --model vgg16 --steps 30 --warmup 8 2>&1 | grep -E '"metric"|rror' | tee gpurun_out/bench_vgg16_dear_1gpu_r2.json | cut -c1-200
echo "=== vgg16 reference 1 GPU"
$B --model vgg16 --impl reference --steps 15 --warmup 5 2>&1 | grep -E '"metric"|rror' | tee gpurun_out/bench_vgg16_reference_1gpu_r2.json | cut -c1-200
echo "=== ncu: hand-written FFN kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ffn_hw_kernel -s 4 -c 2 -f -o gpurun_out/prof_tc_ffn_hw python tools/bert_ops_bench.py --sections gemm --json /dev/null 2>&1 | tail -3
echo "=== done"
