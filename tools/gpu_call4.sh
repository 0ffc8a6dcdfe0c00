#!/bin/bash
# round 2, call 4: discriminators + VGG on the tape; precision modes (split = strict parity, mixed = generator on single terms)
set -u
out=gpurun_out/r02c4
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_nhwc.py -q -m gpu -k "pack_into or maxpool" -p no:cacheprovider 2>&1 | tail -8 > $out/new_kernels.log
for mode in split mixed; do
  COCOS_CONV_PRECISION=$mode timeout 900 python tools/parity_report.py > $out/parity_$mode.txt 2>&1
  COCOS_CONV_PRECISION=$mode timeout 600 python bench.py --no-cpu-baseline 2>$out/bench_$mode.err | tail -1 > $out/bench_$mode.json
done
COCOS_CUDA_GRAPH=0 timeout 600 python tools/profile_step.py --b 8 --cudnn_benchmark --rows 100 --no_table > $out/profile_step_eager_split.txt 2>&1
tail -3 $out/new_kernels.log
for mode in split mixed; do
  grep -v Warn $out/parity_$mode.txt | cut -c1-330
  cut -c1-200 $out/bench_$mode.json; tail -2 $out/bench_$mode.err
done
head -14 $out/profile_step_eager_split.txt
