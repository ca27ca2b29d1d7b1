#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_e2e.log) 2>&1
for i in 1 2; do
echo "=== bench default run $i"; timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default_1gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
done
echo "=== bench default eager"; timeout 400 python bench.py --graph 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
echo "=== done"
