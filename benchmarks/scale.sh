#!/bin/bash
# C3 strong scaling through bench.py, as the driver launches it.  usage: scale.sh <N> [extra bench args]
N=$1; shift
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | grep '^{'
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | grep '^{'
fi
