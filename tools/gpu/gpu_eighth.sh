#!/bin/bash
# tcgen05 GEMM configuration sweep (tile / cluster / scheduler, fast-GELU epilogue), BERT with the best ones,
# per-kernel breakdown of the BERT step, full GPU test-suite
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_eighth.log) 2>&1
export DEAR_TIMEOUT_S=120
echo "=== op micro-benchmarks (all variants)"; timeout 300 python tools/bert_ops_bench.py --json gpurun_out/bert_ops_bench_v2.json 2>&1 | grep -v "^ *\"tile" | tail -70
eval $(python - <<'PY'
import json
try:
    r = json.load(open("gpurun_out/bert_ops_bench_v2.json"))["us"]
    def best(prefix):
        c = {k: v for k, v in r.items() if k.startswith(prefix + "_v") and v == v and v != float("inf")}
        return min(c, key=c.get).rsplit("_v", 1)[1] if c else "0"
    print("export DEAR_TC_UP_VARIANT=%s DEAR_TC_DOWN_VARIANT=%s DEAR_TC_DGELU_VARIANT=%s" % (
        best("up_gelu_tcgen05"), best("down_tcgen05"), best("dgrad_dgelu_tcgen05")))
except Exception as e:
    print("echo variant selection failed: %s" % e)
PY
)
echo "variants: up=$DEAR_TC_UP_VARIANT down=$DEAR_TC_DOWN_VARIANT dgelu=$DEAR_TC_DGELU_VARIANT"
echo "=== tc/ln/adam tests"; timeout 300 python -m pytest tests/test_tc_gemm.py tests/test_adam.py -m gpu -q --timeout 200 2>&1 | tail -12
B="timeout 240 python bench.py --model bert --steps 20 --warmup 8"
echo "=== bert fused ln, eager ffn";          $B --fused-ln 1 --tc-ffn 0 --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_ablate_ln.json | cut -c1-200
echo "=== bert fused ln + tc ffn (best)";     $B --fused-ln 1 --tc-ffn 1 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_fused_1gpu.json | cut -c1-900
echo "=== bert fused ln + tc ffn, cublas down"; DEAR_TC_DOWN=0 $B --fused-ln 1 --tc-ffn 1 --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_ablate_tc_cublas_down.json | cut -c1-200
echo "=== bert step kernel breakdown (fused ln, eager ffn)"; timeout 200 python tools/profile_step.py --model bert --dtype bf16 --steps 3 --warmup 5 --extra "--fused-ln 1 --tc-ffn 0" --out gpurun_out/step_profile_bert_p1 2>&1 | tail -150 | cut -c1-220
rm -f gpurun_out/step_profile_bert_p1.rank0.trace.json
echo "=== full gpu test-suite"; timeout 700 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8
echo "=== done"
