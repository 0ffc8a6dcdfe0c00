#!/bin/bash
# N = 2: the gradient all-reduce captured into the iteration's CUDA graph (netG part overlapped with netCorr's backward)
set -u
out=gpurun_out/r02dp2b
mkdir -p $out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err
echo "exit $?" >> $out/bench_n2.err
COCOS_OVERLAP_ALLREDUCE=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2_no_overlap.json 2> $out/bench_n2_no_overlap.err
echo "exit $?" >> $out/bench_n2_no_overlap.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $out/bench_ref_n2.json 2> $out/bench_ref_n2.err
for f in bench_n2 bench_n2_no_overlap; do tail -1 $out/$f.json | cut -c1-200; tail -3 $out/$f.err | cut -c1-200; done
tail -1 $out/bench_ref_n2.json | cut -c1-200
