#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_4gpu.log) 2>&1
bash profiles/r2/scripts/r2_bo.sh 4 2>&1 | grep -E "===|BO Tuning|Total|rror"
echo "=== done"
