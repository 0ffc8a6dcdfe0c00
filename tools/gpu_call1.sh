#!/bin/bash
# round 2, call 1: where does the iteration spend its GPU time now + does K2w flat mode work
set -u
out=gpurun_out/r02c1
mkdir -p $out
COCOS_WGRAD_NARROW=1 timeout 300 python -m pytest tests/test_gpu_corr.py -q -m gpu -k "flat_mode" 2>&1 | tail -15 > $out/pytest_flat_mode.log
COCOS_CUDA_GRAPH=0 timeout 600 python tools/profile_step.py --b 8 --cudnn_benchmark --rows 150 --no_table > $out/profile_step_eager.txt 2>&1
COCOS_WGRAD_NARROW=1 timeout 400 python bench.py --no-cpu-baseline 2>&1 | tail -1 > $out/bench_wgrad_narrow.json
tail -5 $out/pytest_flat_mode.log
head -12 $out/profile_step_eager.txt
cut -c1-200 $out/bench_wgrad_narrow.json
