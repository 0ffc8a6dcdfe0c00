#!/bin/bash
set -u
out=gpurun_out/r02c3
mkdir -p $out
timeout 900 python tools/parity_report.py > $out/parity_default.txt 2>&1
COCOS_NHWC=0 timeout 600 python tools/parity_report.py ade20k_train ade20k_infer_mk3 > $out/parity_old_path.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline 2>$out/bench_default.err | tail -1 > $out/bench_default.json
COCOS_CUDA_GRAPH=0 timeout 600 python tools/profile_step.py --b 8 --cudnn_benchmark --rows 100 --no_table > $out/profile_step_eager.txt 2>&1
cat $out/parity_default.txt | cut -c1-600
cut -c1-400 $out/bench_default.json; tail -3 $out/bench_default.err
head -14 $out/profile_step_eager.txt
