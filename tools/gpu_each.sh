#!/bin/bash
# Run pytest selections in separate processes so a trapped kernel (sticky CUDA error)
# in one selection does not mask the others.  usage: tools/gpu_each.sh file "k-expr" ...
f=$1; shift
for k in "$@"; do
  echo "=== $f -k '$k'"
  timeout 300 python -m pytest "$f" -q -m gpu -k "$k" -x 2>&1 | tail -15
done
