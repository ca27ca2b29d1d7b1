#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_fourth.log) 2>&1
echo "=== bench BERT-large bf16 dear graph"; timeout 500 python bench.py --model bert --steps 15 --warmup 6 --graph 1 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_bert_dear_graph_1gpu.json
echo "=== bench BERT-large bf16 dear eager (grid 96)"; timeout 500 python bench.py --model bert --steps 15 --warmup 6 --no-e2e 2>&1 | tail -1
echo "=== bench resnet50 fp32 CL eager (grid 96)"; timeout 500 python bench.py --steps 30 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_fp32_cl_g96.json
echo "=== bench resnet50 fp32 CL momentum 0.9 wd"; timeout 500 python bench.py --steps 30 --warmup 10 --momentum 0.9 --no-e2e 2>&1 | tail -1
echo "=== sanitizers"; bash tools/sanitize.sh
echo "=== kernel bench P=1 defaults"; timeout 200 python tools/kernel_bench.py --sizes-mb 4,24,64,392 --out gpurun_out/kernel_bench_p1.json 2>&1 | tail -4
echo "=== ncu full capture (defaults, P=1, 24 MB bucket)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rs_kernel|ag_kernel" -s 12 -c 4 -o gpurun_out/prof_fused_p1_v3 python tools/kernel_bench.py --sizes-mb 24 --iters 3 --nccl 0 > gpurun_out/ncu_full_stdout.log 2>&1
echo "=== ncu launch list, steady state"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 700 --csv --log-file gpurun_out/launches_steady.csv python bench.py --steps 4 --warmup 14 --no-e2e > gpurun_out/ncu_launches_stdout.log 2>&1
tail -1 gpurun_out/ncu_launches_stdout.log | cut -c1-200
echo "=== done"
