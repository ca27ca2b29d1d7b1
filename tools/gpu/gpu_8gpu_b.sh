#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_8gpu_b.log) 2>&1
export DEAR_TIMEOUT_S=180
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "=== bench dear 8 GPUs fp32 fused graph"; timeout 300 $TR --master-port 29911 bench.py --gpus 8 --steps 30 --warmup 10 --graph 1 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_dear_8gpu_fused_graph.json | cut -c1-400
echo "=== bench dear 8 GPUs fp32 fused eager"; timeout 300 $TR --master-port 29912 bench.py --gpus 8 --steps 30 --warmup 10 --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_dear_8gpu_fused_eager.json | cut -c1-400
echo "=== bench dear 8 GPUs BERT bf16 graph"; timeout 300 $TR --master-port 29913 bench.py --gpus 8 --model bert --steps 15 --warmup 6 --graph 1 --no-e2e 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_bert_dear_8gpu_graph.json | cut -c1-400
echo "=== step profile BERT 8 GPUs (eager)"; timeout 300 $TR --master-port 29914 tools/profile_step.py --model bert --steps 3 --out gpurun_out/step_profile_bert_p8 2>&1 | tail -32
echo "=== done"
