#!/bin/bash
# Multi-node launcher (reference: */horovod_mpi_cj.sh with configs/cluster16..64 hostfiles over mpirun).
# Run ONCE PER NODE with the same arguments except node_rank:
#   cluster=16 node_rank=0 master=10.0.0.1 dnn=resnet50 scripts/launch_multinode.sh
#   cluster=16 node_rank=1 master=10.0.0.1 dnn=resnet50 scripts/launch_multinode.sh
# Across nodes the engine runs on the nccl backend (bucketing, overlap, sharded update unchanged; the fused
# symmetric-memory kernels need one NVSwitch domain).  Same environment "flags" as scripts/launch.sh.
dnn="${dnn:-resnet50}"; bs="${bs:-64}"; method="${method:-dear}"; dtype="${dtype:-fp32}"; threshold="${threshold:-25}"
senlen="${senlen:-64}"; cluster="${cluster:-16}"; node_rank="${node_rank:?node_rank=0..nnodes-1}"; master="${master:?master=<ip of node 0>}"
here="$(cd "$(dirname "$0")/.." && pwd)"
[ -f "$here/configs/envs.conf" ] && source "$here/configs/envs.conf"
source "$here/configs/cluster$cluster" || exit 1
if [[ "$dnn" == bert* ]]; then
  driver="$here/benchmarks/bert_benchmark.py"; extra="--sentence-len $senlen"
else
  driver="$here/benchmarks/imagenet_benchmark.py"; extra=""
fi
exec "${PY:-python}" -m torch.distributed.run --nnodes "$nnodes" --node-rank "$node_rank" --nproc-per-node "$nworkers" \
  --master-addr "$master" --master-port "${MASTER_PORT:-29400}" "$driver" --model "$dnn" --batch-size "$bs" \
  --method "$method" --dtype "$dtype" --threshold "$threshold" $extra "$@"
