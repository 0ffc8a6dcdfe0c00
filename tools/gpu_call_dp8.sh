#!/bin/bash
# N = 8 (and 4): the driver's scaling run, once, to see the captured all-reduce work at full width
set -u
out=gpurun_out/r02dp8
mkdir -p $out
N=${1:-8}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_n$N.json 2> $out/bench_n$N.err
echo "exit $?" >> $out/bench_n$N.err
tail -1 $out/bench_n$N.json | cut -c1-260; grep -v "Warning\|warn\|run_backward\|^$" $out/bench_n$N.err | tail -6 | cut -c1-200
