#!/bin/bash
# Race / memory checking of the fused kernels on ONE GPU (the reference has no sanitizer usage at all,
# SURVEY.md §5.2).  memcheck: out-of-bounds / misaligned accesses of the kernels on odd sizes;
# racecheck: shared-memory hazards (segment / hyper tables, grid-arrive flag).
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=dear --print-limit 20 \
    python tools/kernel_bench.py --sizes-mb 0.37,3.1 --iters 2 --nccl 0 2>&1 | grep -E "ERROR SUMMARY|Error|error|RACECHECK SUMMARY|bucket_mb" | head -12
done
