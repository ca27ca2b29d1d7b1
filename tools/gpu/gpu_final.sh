#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_final.log) 2>&1
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench default"; timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default_1gpu.json | cut -c1-1200
echo "=== bench reference default"; timeout 400 python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_default_reference_1gpu.json | cut -c1-600
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6
echo "=== ncu fused bn kernels (final)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bn_" -s 900 -c 12 -o gpurun_out/prof_bn_act python bench.py --steps 2 --warmup 3 --no-e2e --graph 0 > gpurun_out/ncu_bn_stdout.log 2>&1
echo "=== done"
