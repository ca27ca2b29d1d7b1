#!/bin/bash
# one ncu --set full capture of the BERT-layer kernels (tcgen05 GEMMs, fused LN, bias+GELU backward)
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_eleventh.log) 2>&1
timeout 170 ncu --set full --clock-control none --import-source on -k regex:'device_kernel|ln_fwd_kernel|ln_bwd_kernel|bias_gelu_bwd' \
  --launch-skip 12 --launch-count 6 -f -o gpurun_out/prof_bert_ops python tools/ncu_bert_ops.py 2>&1 | tail -15
ls -la gpurun_out/prof_bert_ops.ncu-rep
echo "=== done"
