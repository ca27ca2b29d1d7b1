#!/bin/bash
# 2-GPU validation: tests over real NVLink P2P, kernel micro-benchmarks (IPC and VMM+NVLS), bench both arms
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_2gpu.log) 2>&1
nvidia-smi -L; nvidia-smi topo -m | head -8
export DEAR_TIMEOUT_S=120
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== pytest gpu (2 GPUs)"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -8
echo "=== kernel bench P=2 (IPC)"; timeout 300 $TR --master-port 29601 tools/kernel_bench.py --out gpurun_out/kernel_bench_p2_ipc.json 2>&1 | grep -v Warning | tail -8
echo "=== kernel bench P=2 (VMM + multicast)"; DEAR_PROVIDER=vmm DEAR_MULTICAST=1 timeout 300 $TR --master-port 29602 tools/kernel_bench.py --nccl 0 --out gpurun_out/kernel_bench_p2_vmm_mc.json 2>&1 | grep -v Warning | tail -12
echo "=== kernel bench P=2 (VMM, no multicast)"; DEAR_PROVIDER=vmm timeout 300 $TR --master-port 29603 tools/kernel_bench.py --nccl 0 --sizes-mb 4,24 --out gpurun_out/kernel_bench_p2_vmm.json 2>&1 | grep -v Warning | tail -6
echo "=== kernel bench P=1"; timeout 200 python tools/kernel_bench.py --sizes-mb 4,24,64 --out gpurun_out/kernel_bench_p1.json 2>&1 | tail -4
echo "=== bench dear 2 GPUs fp32 CL"; timeout 400 $TR --master-port 29604 bench.py --gpus 2 --steps 20 --warmup 8 2>&1 | grep '"metric"' | tee gpurun_out/bench_dear_2gpu.json
echo "=== bench reference 2 GPUs"; timeout 400 $TR --master-port 29605 bench.py --impl reference --gpus 2 --steps 20 --warmup 8 2>&1 | grep -E '"metric"|unavailable|Error' | tee gpurun_out/bench_reference_2gpu.json
echo "=== bench dear 2 GPUs bf16 CL graph"; timeout 400 $TR --master-port 29606 bench.py --gpus 2 --steps 20 --warmup 8 --dtype bf16 --graph 1 --no-e2e 2>&1 | grep -E '"metric"|Error' | tee gpurun_out/bench_dear_2gpu_bf16_graph.json
echo "=== done"
