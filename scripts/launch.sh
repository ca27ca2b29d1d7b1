#!/bin/bash
# Single-node launcher (replaces the reference's mpirun + hostfile scripts, */horovod_mpi_cj.sh).
# Environment "flags", as in the reference:
#   dnn (resnet50) bs (64) nworkers (8) method (dear) dtype (fp32) threshold (25) exclude_parts ("")
#   senlen (64, BERT only)  nstreams (1)  graph (0)
#   compressor (none; topk | eftopk | gaussian | signum | efsignum | gtopk | gtopkef select the sparse WFBP baseline with
#   density 0.001 and one 64 Mi-element group, as the reference launcher does)  density  mgwfbp (0)  asc (0)  rdma (0)
dnn="${dnn:-resnet50}"; bs="${bs:-64}"; nworkers="${nworkers:-8}"; method="${method:-dear}"
dtype="${dtype:-fp32}"; threshold="${threshold:-25}"; exclude_parts="${exclude_parts:-}"; senlen="${senlen:-64}"
nstreams="${nstreams:-1}"; graph="${graph:-0}"
compressor="${compressor:-none}"; mgwfbp="${mgwfbp:-0}"; asc="${asc:-0}"; rdma="${rdma:-0}"
here="$(cd "$(dirname "$0")/.." && pwd)"
[ -f "$here/configs/envs.conf" ] && source "$here/configs/envs.conf"
# cluster=N picks configs/clusterN (the reference selected an MPI hostfile the same way)
[ -n "$cluster" ] && [ -f "$here/configs/cluster$cluster" ] && source "$here/configs/cluster$cluster"
if [[ "$dnn" == bert* ]]; then
  driver="$here/benchmarks/bert_benchmark.py"; extra="--sentence-len $senlen"
else
  driver="$here/benchmarks/imagenet_benchmark.py"; extra=""
fi
if [ "$compressor" != "none" ]; then
  # reference */horovod_mpi_cj.sh: "--density 0.001 --compressor $compressor ... --threshold 67108864"
  [ "$method" = "dear" ] && method="wfbp"
  threshold=67108864
  extra="$extra --compressor $compressor --density ${density:-0.001}"
fi
[ "$mgwfbp" = "1" ] && extra="$extra --mgwfbp"
[ "$asc" = "1" ] && extra="$extra --asc"
[ "$rdma" = "1" ] && extra="$extra --rdma"
port="${MASTER_PORT:-$((20000 + RANDOM % 20000))}"
exec "${PY:-python}" -m torch.distributed.run --nnodes=1 --nproc-per-node "$nworkers" \
  --master-addr 127.0.0.1 --master-port "$port" "$driver" --model "$dnn" --batch-size "$bs" \
  --method "$method" --dtype "$dtype" --threshold "$threshold" --nstreams "$nstreams" --graph "$graph" \
  ${exclude_parts:+--exclude-parts "$exclude_parts"} $extra "$@"
