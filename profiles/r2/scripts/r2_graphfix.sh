#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_graphfix.log) 2>&1
export DEAR_TIMEOUT_S=180
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_grad_accumulation.py tests/test_gpu_nccl_backend.py tests/test_tc_gemm.py -m gpu -x -q --timeout 280 2>&1 | tail -40
echo "=== bench resnet50 1 GPU (default: graph, rotated body)"
timeout 300 python bench.py --steps 40 --warmup 8 2>&1 | grep -E '"metric"|rror' | cut -c1-900
echo "=== bench bert 1 GPU bf16"
timeout 300 python bench.py --model bert --steps 30 --warmup 8 --no-e2e 2>&1 | grep -E '"metric"|rror' | cut -c1-400
echo "=== bench bert 1 GPU fp32 / fp32 no fused ln / reference fp32 / reference bf16"
timeout 300 python bench.py --model bert --dtype fp32 --steps 10 --warmup 5 --no-e2e 2>&1 | grep -E '"metric"|rror' | cut -c1-300
timeout 300 python bench.py --model bert --dtype fp32 --fused-ln 0 --steps 10 --warmup 5 --no-e2e 2>&1 | grep -E '"metric"|rror' | cut -c1-300
timeout 300 python bench.py --model bert --dtype fp32 --impl reference --steps 10 --warmup 5 --no-e2e 2>&1 | grep -E '"metric"|rror' | cut -c1-300
timeout 300 python bench.py --model bert --dtype bf16 --impl reference --steps 20 --warmup 5 --no-e2e 2>&1 | grep -E '"metric"|rror' | cut -c1-300
echo "=== done"
