#!/bin/bash
# usage: multi_gpu.sh <ngpus>   - C3 / C4 / C5 through bench.py under torchrun, one JSON line each
N=$1
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | grep '^{' ; }
run --workload c3
run --workload c4
run --workload c5 --batch 64
run --workload c5 --batch 8
run --workload c5 --batch 1
