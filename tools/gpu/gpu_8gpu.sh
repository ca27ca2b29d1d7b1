#!/bin/bash
# 8-GPU validation on one NVSwitch node: correctness at P=8, both bench arms, kernel bus bandwidth (P2P and NVLS)
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_8gpu.log) 2>&1
nvidia-smi -L | head -8
export DEAR_TIMEOUT_S=180
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== pytest P=8 correctness"; timeout 400 python -m pytest tests/test_gpu_fused.py -m gpu -x -q --timeout 300 -k "multi_rank and 8" 2>&1 | tail -4
echo "=== bench dear 8 GPUs fp32 CL"; timeout 400 $TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 8 2>&1 | grep -E '"metric"|Error|error' | tee gpurun_out/bench_dear_8gpu.json
echo "=== bench reference 8 GPUs"; timeout 400 $TR --nproc-per-node 8 --master-port 29702 bench.py --impl reference --gpus 8 --steps 20 --warmup 8 2>&1 | grep -E '"metric"|unavailable|Error' | tee gpurun_out/bench_reference_8gpu.json
echo "=== kernel bench P=8 (IPC, P2P)"; timeout 300 $TR --nproc-per-node 8 --master-port 29703 tools/kernel_bench.py --sizes-mb 1,4,24,64,392 --out gpurun_out/kernel_bench_p8_ipc.json 2>&1 | grep bucket_mb
echo "=== kernel bench P=8 (VMM + NVLS multicast)"; DEAR_PROVIDER=vmm DEAR_MULTICAST=1 timeout 300 $TR --nproc-per-node 8 --master-port 29704 tools/kernel_bench.py --nccl 0 --sizes-mb 1,4,24,64,392 --out gpurun_out/kernel_bench_p8_vmm_mc.json 2>&1 | grep -E "bucket_mb|Error|error" | head -8
echo "=== bench dear 8 GPUs vgg16"; timeout 400 $TR --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --model vgg16 --steps 15 --warmup 6 --no-e2e 2>&1 | grep -E '"metric"|Error' | tee gpurun_out/bench_vgg16_dear_8gpu.json
echo "=== bench reference 8 GPUs vgg16"; timeout 400 $TR --nproc-per-node 8 --master-port 29706 bench.py --impl reference --gpus 8 --model vgg16 --steps 15 --warmup 6 --no-e2e 2>&1 | grep -E '"metric"|Error' | tee gpurun_out/bench_vgg16_reference_8gpu.json
echo "=== bench dear 8 GPUs BERT-large bf16"; timeout 400 $TR --nproc-per-node 8 --master-port 29707 bench.py --gpus 8 --model bert --steps 15 --warmup 6 --no-e2e 2>&1 | grep -E '"metric"|Error' | tee gpurun_out/bench_bert_dear_8gpu.json
echo "=== bench reference 8 GPUs BERT-large fp32"; timeout 400 $TR --nproc-per-node 8 --master-port 29708 bench.py --impl reference --gpus 8 --model bert --steps 15 --warmup 6 --no-e2e 2>&1 | grep -E '"metric"|Error' | tee gpurun_out/bench_bert_reference_8gpu.json
echo "=== bench dear 4 GPUs fp32 CL"; timeout 400 $TR --nproc-per-node 4 --master-port 29709 bench.py --gpus 4 --steps 20 --warmup 8 --no-e2e 2>&1 | grep -E '"metric"|Error' | tee gpurun_out/bench_dear_4gpu.json
echo "=== bench dear 8 GPUs fp32 CL, multicast"; DEAR_PROVIDER=vmm DEAR_MULTICAST=1 timeout 400 $TR --nproc-per-node 8 --master-port 29710 bench.py --gpus 8 --steps 20 --warmup 8 --no-e2e 2>&1 | grep -E '"metric"|Error' | tee gpurun_out/bench_dear_8gpu_mc.json
echo "=== done"
