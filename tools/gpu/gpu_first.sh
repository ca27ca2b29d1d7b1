#!/bin/bash
# first contact with the GPU: smoke, single-GPU bench variants, GPU tests
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_first.log) 2>&1
nvidia-smi -L
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 
echo "=== bench fp32 CL eager"; timeout 400 python bench.py --steps 20 --warmup 8 | tail -1 | tee gpurun_out/bench_fp32_cl.json
echo "=== bench fp32 NCHW eager"; timeout 400 python bench.py --steps 20 --warmup 8 --channels-last 0 --no-e2e | tail -1 | tee gpurun_out/bench_fp32_nchw.json
echo "=== bench bf16 CL eager"; timeout 400 python bench.py --steps 20 --warmup 8 --dtype bf16 | tail -1 | tee gpurun_out/bench_bf16_cl.json
echo "=== bench fp32 CL graph"; timeout 400 python bench.py --steps 20 --warmup 8 --graph 1 | tail -1 | tee gpurun_out/bench_fp32_cl_graph.json
echo "=== bench bf16 CL graph"; timeout 400 python bench.py --steps 20 --warmup 8 --dtype bf16 --graph 1 | tail -1 | tee gpurun_out/bench_bf16_cl_graph.json
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -30
echo "=== done"
