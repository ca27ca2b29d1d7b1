#!/bin/bash
# 2 GPUs: full multi-GPU tests (incl. NVLS, direct wgrad), multi-GPU ncu capture of Kernel A / B, BO tuner.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_2gpu_c.log) 2>&1
export DEAR_TIMEOUT_S=180
echo "=== tests"
timeout 900 python -m pytest tests/test_kernels_direct.py tests/test_gpu_fused.py tests/test_gpu_nccl_backend.py tests/test_grad_accumulation.py -m gpu -q --timeout 280 2>&1 | tail -6
echo "=== ncu capture at world size 2: rs_kernel / ag_kernel (rank 0 under ncu)"
MASTER_PORT=29961 timeout 500 bash tools/ncu_multi_gpu.sh 2 rs_kernel gpurun_out/prof_rs_kernel_p2 --sizes-mb 64
MASTER_PORT=29962 timeout 500 bash tools/ncu_multi_gpu.sh 2 ag_kernel gpurun_out/prof_ag_kernel_p2 --sizes-mb 64
ncu --query-metrics 2>/dev/null | grep -i -E "^nvl|nvlrx|nvltx" | head -40 > gpurun_out/nvlink_metric_names.txt
echo "=== BO tuner, BERT-base, 2 GPUs"
bash profiles/r2/scripts/r2_bo.sh 2 2>&1 | grep -E "===|optimal|Total|rror"
echo "=== done"
