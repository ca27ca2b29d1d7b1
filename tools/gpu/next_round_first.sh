#!/bin/bash
# First GPU call of the next round: everything that was written after this round's GPU
# budget was spent.  1 GPU, ~4 minutes.  Each step is wrapped in `timeout`; the hand-written kernel traps instead
# of hanging if its barrier protocol is wrong.
mkdir -p gpurun_out
exec > >(tee gpurun_out/next_round_first.log) 2>&1
export DEAR_TIMEOUT_S=120 DEAR_TEST_UNVALIDATED=1 DEAR_TC_EXPERIMENTAL=1
echo "=== hand-written tcgen05 kernel"; timeout 120 python -m pytest tests/test_tc_gemm.py -m gpu -q --timeout 100 -k handwritten 2>&1 | tail -15
echo "=== its speed vs cuBLAS + GELU and the CUTLASS-collective variants"; timeout 200 python tools/bert_ops_bench.py --sections gemm --json gpurun_out/bert_ops_bench_hw.json 2>&1 | grep -E "up_gelu|up_gemm|failed"
echo "=== engine on the nccl backend (world 1), gradient accumulation on the fused path"
timeout 300 python -m pytest tests/test_gpu_nccl_backend.py tests/test_grad_accumulation.py -m gpu -q --timeout 150 2>&1 | tail -8
echo "=== graph capture deferred until the BO tuner has settled"; timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 250 -k 'bo_tuner or adamw' 2>&1 | tail -6
echo "=== BERT with the hand-written FFN kernels (only meaningful if the tests above passed)"
B="timeout 200 python bench.py --model bert --steps 20 --warmup 8 --no-e2e"
$B 2>&1 | grep -E '"metric"|Error|error' | cut -c1-200
DEAR_TC_FFN_IMPL=hw DEAR_TC_DOWN=0 $B --tc-ffn 1 2>&1 | grep -E '"metric"|Error|error' | cut -c1-200
echo "=== done"
