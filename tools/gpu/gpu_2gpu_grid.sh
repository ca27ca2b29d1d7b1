#!/bin/bash
# how many CTAs should the spinning kernels hold when they overlap with compute? (P=2)
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_2gpu_grid.log) 2>&1
export DEAR_TIMEOUT_S=120
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
p=29800
for g in 16 32 96; do
  for m in bert vgg16; do
    p=$((p+1))
    echo "=== bench dear 2 GPUs $m grid=$g"
    DEAR_RS_GRID=$g DEAR_AG_GRID=$g timeout 300 $TR --master-port $p bench.py --gpus 2 --model $m --steps 15 --warmup 6 --no-e2e 2>&1 | grep -E '"metric"|Error' | cut -c1-330
  done
done
echo "=== bench dear 2 GPUs resnet50 grid=32"; timeout 300 $TR --master-port 29820 bench.py --gpus 2 --steps 20 --warmup 8 --no-e2e 2>&1 | grep -E '"metric"|Error' | cut -c1-330
echo "=== bench dear 2 GPUs bert threshold 64 grid=32"; timeout 300 $TR --master-port 29821 bench.py --gpus 2 --model bert --threshold 64 --steps 15 --warmup 6 --no-e2e 2>&1 | grep -E '"metric"|Error' | cut -c1-330
echo "=== kernel bench P=2 grid=32"; timeout 300 $TR --master-port 29822 tools/kernel_bench.py --nccl 0 --sizes-mb 4,24,64,392 --out gpurun_out/kernel_bench_p2_g32.json 2>&1 | grep bucket_mb
echo "=== done"
