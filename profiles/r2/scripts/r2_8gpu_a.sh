#!/bin/bash
# 8 GPUs, session A: NVLink probe, multi-rank correctness, Kernel A/B micro-benchmarks, the three model benchmarks.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_8gpu_a.log) 2>&1
export DEAR_TIMEOUT_S=180
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
J='"metric"|rror'
echo "=== probe 8 devices"
timeout 100 build/p2p_probe 8 392 32,64 2>&1 | grep -v "^#"
timeout 60 build/p2p_probe 8 24 32,64 2>&1 | grep -E "ldg |tma4k|stg_na"
echo "=== correctness at 8 ranks (fused engine, pipelined Kernel A at 4 ranks)"
timeout 400 python -m pytest tests/test_gpu_fused.py tests/test_kernels_direct.py -m gpu -q --timeout 300 -k "multi_rank or pipelined or nvls or picked" 2>&1 | tail -4
echo "=== kernel bench P=8 auto (one-shot < 128 MB <= pipe)"
timeout 200 $TR --master-port 29811 tools/kernel_bench.py --sizes-mb 1,4,24,64,392 --out gpurun_out/kernel_bench_p8_r2_auto.json 2>&1 | grep -E '^\{|rror' | cut -c1-700
echo "=== kernel bench P=8 pipe everywhere"
DEAR_PIPE_MIN_MB=2 timeout 200 $TR --master-port 29812 tools/kernel_bench.py --sizes-mb 24,64,392 --nccl 0 --out gpurun_out/kernel_bench_p8_r2_pipe.json 2>&1 | grep -E '^\{|rror' | cut -c1-330
echo "=== kernel bench P=8 one-shot everywhere, grids 32 / 96"
DEAR_RS_ALGO=oneshot DEAR_RS_GRID=32 DEAR_AG_GRID=32 timeout 200 $TR --master-port 29813 tools/kernel_bench.py --sizes-mb 24,64,392 --nccl 0 2>&1 | grep -E '^\{|rror' | cut -c1-330
DEAR_RS_ALGO=oneshot DEAR_RS_GRID=96 DEAR_AG_GRID=96 timeout 200 $TR --master-port 29814 tools/kernel_bench.py --sizes-mb 24,64,392 --nccl 0 2>&1 | grep -E '^\{|rror' | cut -c1-330
echo "=== kernel bench P=8 pipe grid 96"
DEAR_PIPE_MIN_MB=2 DEAR_RS_GRID=96 timeout 200 $TR --master-port 29815 tools/kernel_bench.py --sizes-mb 24,64,392 --nccl 0 2>&1 | grep -E '^\{|rror' | cut -c1-330
echo "=== bench resnet50 8 GPUs (default: graph, rotated body = update overlaps forward)"
timeout 300 $TR --master-port 29821 bench.py --gpus 8 --steps 60 --warmup 10 2>&1 | grep -E "$J" | tee gpurun_out/bench_resnet50_dear_8gpu_r2.json | cut -c1-1200
echo "=== bench resnet50 8 GPUs natural body (round-1 configuration)"
timeout 300 $TR --master-port 29822 bench.py --gpus 8 --steps 60 --warmup 10 --overlap-update 0 --no-e2e 2>&1 | grep -E "$J" | tee gpurun_out/bench_resnet50_dear_8gpu_r2_natural.json | cut -c1-300
echo "=== bench vgg16 8 GPUs graph"
timeout 300 $TR --master-port 29823 bench.py --gpus 8 --model vgg16 --steps 40 --warmup 10 2>&1 | grep -E "$J" | tee gpurun_out/bench_vgg16_dear_8gpu_r2.json | cut -c1-1200
echo "=== bench bert-large bf16 8 GPUs graph"
timeout 300 $TR --master-port 29824 bench.py --gpus 8 --model bert --steps 40 --warmup 10 2>&1 | grep -E "$J" | tee gpurun_out/bench_bert_dear_8gpu_r2.json | cut -c1-1200
echo "=== reference arms at 8 GPUs: vgg16 (DeAR fp32), bert bf16 (DDP)"
timeout 300 $TR --master-port 29825 bench.py --gpus 8 --model vgg16 --impl reference --steps 20 --warmup 5 --no-e2e 2>&1 | grep -E "$J" | tee gpurun_out/bench_vgg16_reference_8gpu_r2.json | cut -c1-300
timeout 300 $TR --master-port 29826 bench.py --gpus 8 --model bert --impl reference --steps 20 --warmup 5 --no-e2e 2>&1 | grep -E "$J" | tee gpurun_out/bench_bert_reference_bf16_8gpu_r2.json | cut -c1-300
echo "=== done"
