#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_third.log) 2>&1
for g in 48 96 148; do
  echo "=== kernel bench P=1 grid=$g"; DEAR_RS_GRID=$g DEAR_AG_GRID=$g timeout 200 python tools/kernel_bench.py --sizes-mb 4,24,64,392 --out gpurun_out/kernel_bench_p1_g$g.json 2>&1 | tail -4
done
echo "=== bench BERT-large bf16 dear"; timeout 500 python bench.py --model bert --steps 15 --warmup 5 2>&1 | tail -2 | tee gpurun_out/bench_bert_dear_1gpu.json
echo "=== bench BERT-large reference (fp32)"; timeout 500 python bench.py --impl reference --model bert --steps 15 --warmup 5 2>&1 | tail -2 | tee gpurun_out/bench_bert_reference_1gpu.json
echo "=== bench BERT-large fp32 dear"; timeout 500 python bench.py --model bert --dtype fp32 --steps 15 --warmup 5 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_bert_dear_fp32_1gpu.json
echo "=== bench vgg16 dear"; timeout 500 python bench.py --model vgg16 --steps 15 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_vgg16_dear_1gpu.json
echo "=== bench vgg16 reference"; timeout 500 python bench.py --impl reference --model vgg16 --steps 15 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_vgg16_reference_1gpu.json
echo "=== ncu launch list (one step, eager fp32 CL)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_launches_stdout.log 2>&1
tail -2 gpurun_out/ncu_launches_stdout.log | cut -c1-300
echo "=== ncu full capture of the fused kernels (kernel_bench, P=1)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rs_kernel|ag_kernel" -s 12 -c 4 -o gpurun_out/prof_fused_p1_v2 python tools/kernel_bench.py --sizes-mb 24 --iters 3 > gpurun_out/ncu_full_stdout.log 2>&1
ls -la gpurun_out | tail -12
echo "=== done"
