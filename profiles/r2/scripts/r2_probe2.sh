#!/bin/bash
# 2 GPUs: NVLink access-pattern probe (LDG vs TMA bulk, grid sweep) for the fused kernels' pull / push phases.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_probe2.log) 2>&1
nvidia-smi topo -m | head -8
for mb in 24 392; do
  echo "=== probe ndev=2 bucket=${mb}MB"
  timeout 120 build/p2p_probe 2 $mb 16,32,64,128
done
echo "=== done"
