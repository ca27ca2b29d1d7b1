#!/bin/bash
# Bayesian buffer-size tuner (dopt_rsag_bo) on BERT-base at N GPUs: 3 tuned runs + 3 runs at the 25 MB default.
#   gpurun --gpus N -- 'bash profiles/r2/scripts/r2_bo.sh N'
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee -a gpurun_out/bo_tuner_bert_base_p$N.log) 2>&1
export DEAR_TIMEOUT_S=180
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
ARGS="benchmarks/bert_benchmark.py --model bert_base --batch-size 64 --sentence-len 64 --dtype bf16"
for rep in 1 2 3; do
  echo "=== BERT-base dear-bo $N GPUs, run $rep"
  timeout 400 $TR --master-port $((29930+rep)) $ARGS --method dear-bo --num-warmup-batches 60 --num-iters 5 --num-batches-per-iter 10 2>&1 | grep -E "BO Tuning|Total|rror|Tensor fusion groups" | tail -24
done
for rep in 1 2 3; do
  echo "=== BERT-base dear (25 MB) $N GPUs, run $rep"
  timeout 300 $TR --master-port $((29940+rep)) $ARGS --method dear --num-warmup-batches 20 --num-iters 5 --num-batches-per-iter 10 2>&1 | grep -E "Total|rror" | tail -2
done
echo "=== done"
