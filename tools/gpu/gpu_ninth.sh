#!/bin/bash
# bias-fused LN / bias+GELU kernels, attention backward restructure (unbind), SDPA backends; BERT; full GPU test-suite
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_ninth.log) 2>&1
export DEAR_TIMEOUT_S=120
echo "=== new kernel tests"; timeout 300 python -m pytest tests/test_fused_ln.py -m gpu -q --timeout 200 2>&1 | tail -12
echo "=== op micro-benchmarks"; timeout 240 python tools/bert_ops_bench.py --sections ln,lg,attn --json gpurun_out/bert_ops_bench_v3.json 2>&1 | grep -v '^  *"tile\|^  *"ffn_\|^  *"linear_bias\|^ *\]' | tail -40
B="timeout 240 python bench.py --model bert --steps 20 --warmup 8"
echo "=== bert default (fused ln + bias fusions, cuBLAS ffn)"; $B 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_fused_1gpu.json | cut -c1-1000
echo "=== bert default, efficient-attention backend"; DEAR_SDPA_BACKEND=efficient $B --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_sdpa_efficient.json | cut -c1-200
echo "=== bert eager ops"; $B --fused-ln 0 --tc-ffn 0 --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_ablate_none.json | cut -c1-200
echo "=== bert step kernel breakdown"; timeout 200 python tools/profile_step.py --model bert --dtype bf16 --steps 3 --warmup 5 --top 24 --out gpurun_out/step_profile_bert_p1_v2 2>&1 | grep -E '"name"|us_per_step|per_step"' | paste - - - | cut -c1-200 | head -30
rm -f gpurun_out/step_profile_bert_p1_v2.rank0.trace.json
echo "=== default bench"; timeout 240 python bench.py 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_default_1gpu.json | cut -c1-400
echo "=== full gpu test-suite"; timeout 700 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8
echo "=== done"
