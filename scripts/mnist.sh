#!/bin/bash
# MNIST example (reference mnist.sh): nworkers ranks on this node; add --no-cuda for the CPU/gloo path.
here="$(cd "$(dirname "$0")/.." && pwd)"
nworkers="${nworkers:-2}"
exec "${PY:-python}" -m torch.distributed.run --nnodes=1 --nproc-per-node "$nworkers" --master-addr 127.0.0.1 \
  --master-port "${MASTER_PORT:-29511}" "$here/examples/mnist/pytorch_mnist.py" "$@"
