#!/bin/bash
# 8 GPUs, session B (final configuration, lean: the round's GPU budget is almost spent): kernel bench + the three models.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_8gpu_b.log) 2>&1
export DEAR_TIMEOUT_S=120
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
J='"metric"|rror'
echo "=== bench resnet50 8 GPUs"
timeout 120 $TR --master-port 29821 bench.py --gpus 8 --steps 60 --warmup 10 2>&1 | grep -E "$J" | tee gpurun_out/bench_resnet50_dear_8gpu_r2_final.json | cut -c1-1300
echo "=== bench bert-large bf16 8 GPUs"
timeout 120 $TR --master-port 29824 bench.py --gpus 8 --model bert --steps 40 --warmup 10 2>&1 | grep -E "$J" | tee gpurun_out/bench_bert_dear_8gpu_r2_final.json | cut -c1-1300
echo "=== bench vgg16 8 GPUs"
timeout 120 $TR --master-port 29823 bench.py --gpus 8 --model vgg16 --steps 40 --warmup 10 2>&1 | grep -E "$J" | tee gpurun_out/bench_vgg16_dear_8gpu_r2_final.json | cut -c1-1300
echo "=== kernel bench P=8 (final plan)"
timeout 100 $TR --master-port 29811 tools/kernel_bench.py --sizes-mb 24,64,392 --iters 10 --out gpurun_out/kernel_bench_p8_r2_final.json 2>&1 | grep -E '^\{|rror' | cut -c1-800
echo "=== done"
