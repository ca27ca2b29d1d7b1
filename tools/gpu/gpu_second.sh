#!/bin/bash
# second contact: bf16, CUDA-graph debugging, GPU tests, ncu evidence of the fused kernels
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_second.log) 2>&1
echo "=== bench bf16 CL eager"; timeout 400 python bench.py --steps 20 --warmup 8 --dtype bf16 2>&1 | tail -3 | tee gpurun_out/bench_bf16_cl.json
echo "=== graph debug (resnet18)"; DEAR_GRAPH_DEBUG=1 timeout 300 python -X faulthandler bench.py --model resnet18 --batch-size 16 --steps 5 --warmup 5 --graph 1 --no-e2e 2>&1 | tail -40
echo "=== bench fp32 CL graph"; timeout 400 python -X faulthandler bench.py --steps 20 --warmup 8 --graph 1 2>&1 | tail -5 | tee gpurun_out/bench_fp32_cl_graph.json
echo "=== bench bf16 CL graph"; timeout 400 python -X faulthandler bench.py --steps 20 --warmup 8 --dtype bf16 --graph 1 2>&1 | tail -5 | tee gpurun_out/bench_bf16_cl_graph.json
echo "=== reference arm 1 GPU"; timeout 400 python bench.py --impl reference --steps 20 --warmup 8 2>&1 | tail -2 | tee gpurun_out/bench_reference_1gpu.json
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -15
echo "=== ncu launch list (one step, eager fp32)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_launches_stdout.log 2>&1
echo "=== ncu full capture of the fused kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rs_kernel|ag_kernel" -s 10 -c 6 -o gpurun_out/prof_fused_1gpu python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_full_stdout.log 2>&1
ls -la gpurun_out
echo "=== done"
