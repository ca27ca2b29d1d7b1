#!/bin/bash
# Kernel-level profile of one training iteration (reference: horovod/prof.sh — `nvprof --metrics inst_fp_32` on the
# BERT and ResNet-50 drivers, summed by horovod/extract_profilings.py).  Here: ncu launch list with the duration and the
# FP32 / tensor-pipe instruction counters per kernel, aggregated by tools/ncu_summary.py.
#   dnn=resnet50 bs=64 scripts/prof.sh          dnn=bert bs=32 scripts/prof.sh
# ONE GPU (ncu replays every kernel: never run it under a multi-rank launch; tools/ncu_multi_gpu.sh profiles the fused
# communication kernels at world size N).  Numbers printed by a run under ncu are not benchmark values.
dnn="${dnn:-resnet50}"; bs="${bs:-64}"; method="${method:-dear}"; dtype="${dtype:-fp32}"; senlen="${senlen:-64}"
here="$(cd "$(dirname "$0")/.." && pwd)"
out="${out:-$here/logs/prof}"; mkdir -p "$out"
if [[ "$dnn" == bert* ]]; then
  driver="$here/benchmarks/bert_benchmark.py"; extra="--sentence-len $senlen"
else
  driver="$here/benchmarks/imagenet_benchmark.py"; extra=""
fi
metrics="gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,sm__inst_executed_pipe_tensor.sum"
csv="$out/${dnn}-bs${bs}-${dtype}.csv"
MASTER_ADDR=127.0.0.1 MASTER_PORT="${MASTER_PORT:-29533}" RANK=0 LOCAL_RANK=0 WORLD_SIZE=1 \
ncu --metrics "$metrics" --clock-control none --csv --log-file "$csv" \
  "${PY:-python}" "$driver" --model "$dnn" --batch-size "$bs" --method "$method" --dtype "$dtype" \
  --num-warmup-batches 0 --num-batches-per-iter 1 --num-iters 1 $extra "$@" > "$out/${dnn}-bs${bs}-${dtype}.log" 2>&1
"${PY:-python}" "$here/tools/ncu_summary.py" "$csv" "${top:-25}" | tee "$out/${dnn}-bs${bs}-${dtype}.summary.txt"
