#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_fifth.log) 2>&1
echo "=== pytest fused bn + graph + kernels"; timeout 900 python -m pytest tests/test_fused_bn.py tests/test_kernels_direct.py -m gpu -x -q --timeout 300 2>&1 | tail -8
echo "=== bench resnet50 fp32 CL fused-bn"; timeout 400 python bench.py --steps 30 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_fp32_cl_fusedbn.json
echo "=== bench resnet50 fp32 CL unfused"; timeout 400 python bench.py --steps 30 --warmup 10 --fused-bn 0 --no-e2e 2>&1 | tail -1
echo "=== bench resnet50 bf16 CL fused-bn"; timeout 400 python bench.py --steps 30 --warmup 10 --dtype bf16 --no-e2e 2>&1 | tail -1
echo "=== bench resnet50 fp32 CL fused-bn graph"; timeout 400 python bench.py --steps 30 --warmup 10 --graph 1 --no-e2e 2>&1 | tail -1
echo "=== bench BERT-large bf16 graph"; timeout 400 python bench.py --model bert --steps 15 --warmup 6 --graph 1 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_bert_dear_graph_1gpu.json
echo "=== pytest gpu graph test"; timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -x -q --timeout 300 -k "graph" 2>&1 | tail -5
echo "=== ncu fused bn kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bn_" -s 600 -c 12 -o gpurun_out/prof_bn_act python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_bn_stdout.log 2>&1
echo "=== step profile"; timeout 300 python tools/profile_step.py --model resnet50 --steps 3 2>&1 | tail -30
echo "=== done"
