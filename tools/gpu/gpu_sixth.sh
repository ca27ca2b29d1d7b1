#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_sixth.log) 2>&1
echo "=== pytest fused bn"; timeout 900 python -m pytest tests/test_fused_bn.py -m gpu -q --timeout 300 2>&1 | tail -12
echo "=== bench resnet50 fp32 CL fused-bn eager"; timeout 400 python bench.py --steps 30 --warmup 10 --no-e2e 2>&1 | tail -1
echo "=== bench resnet50 fp32 CL fused-bn graph"; timeout 400 python bench.py --steps 30 --warmup 10 --graph 1 2>&1 | tail -1 | tee gpurun_out/bench_fp32_cl_fusedbn_graph.json
echo "=== bench resnet50 fp32 CL unfused graph"; timeout 400 python bench.py --steps 30 --warmup 10 --graph 1 --fused-bn 0 --no-e2e 2>&1 | tail -1
echo "=== bench resnet50 bf16 CL fused-bn graph"; timeout 400 python bench.py --steps 30 --warmup 10 --dtype bf16 --graph 1 2>&1 | tail -1 | tee gpurun_out/bench_bf16_cl_fusedbn_graph.json
echo "=== ncu fused bn kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bn_" -s 900 -c 12 -o gpurun_out/prof_bn_act python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_bn_stdout.log 2>&1
echo "=== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6
echo "=== done"
