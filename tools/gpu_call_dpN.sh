#!/bin/bash
set -u
N=${1:-2}
out=gpurun_out/r02dp${N}c
mkdir -p $out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err
echo "exit $?" >> $out/bench.err
tail -1 $out/bench.json | cut -c1-240; grep -v "Warning\|warn\|run_backward\|^$\|OMP\|\*\*\*" $out/bench.err | tail -5 | cut -c1-200
