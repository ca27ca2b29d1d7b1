#!/bin/bash
# ncu capture of the fused kernels AT WORLD SIZE N: rank 0 runs under ncu (kernel replay), ranks 1..N-1 run plain.
#   tools/ncu_multi_gpu.sh N <kernel-regex> <out-prefix> [extra kernel_bench args]
# Why kernel replay works here: the cross-GPU flags are monotonic epochs and the peers run AHEAD of the slow profiled
# rank, so their flags for this launch are already in rank 0's signal pad when ncu saves memory for the replay; rank
# 0's own signals are idempotent.  Every spin is bounded (DEAR_SPIN_TIMEOUT_S), so a surprise ends in an error, not a hang.
N=${1:-2}; KREG=${2:-rs_kernel}; OUT=${3:-gpurun_out/prof_multi}; shift 3
export MASTER_ADDR=127.0.0.1 MASTER_PORT=${MASTER_PORT:-29960} WORLD_SIZE=$N LOCAL_WORLD_SIZE=$N DEAR_SPIN_TIMEOUT_S=40 DEAR_TIMEOUT_S=240
ARGS="tools/kernel_bench.py --nccl 0 --iters 4 $*"
pids=""
for r in $(seq 1 $((N-1))); do
  RANK=$r LOCAL_RANK=$r timeout 400 python $ARGS > /dev/null 2>&1 &
  pids="$pids $!"
done
RANK=0 LOCAL_RANK=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:$KREG -s 6 -c 1 \
  --metrics nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes.sum.per_second,nvltx__bytes.sum.per_second \
  -f -o $OUT python $ARGS 2>&1 | tail -5
for p in $pids; do wait $p; done

ls -la $OUT.*
