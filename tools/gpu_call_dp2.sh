#!/bin/bash
# N = 2: the gradient all-reduce captured into the iteration's CUDA graph; eager run for comparison
set -u
out=gpurun_out/r02dp2
mkdir -p $out
export COCOS_GRAPH_TRACE=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err
COCOS_CUDA_GRAPH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2_eager.json 2> $out/bench_n2_eager.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
for f in bench_n1 bench_n2 bench_n2_eager; do tail -1 $out/$f.json | cut -c1-420; done
grep -v "Warning\|warn\|^$\|run_backward" $out/bench_n2.err | tail -25
