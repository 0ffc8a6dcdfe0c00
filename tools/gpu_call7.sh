#!/bin/bash
# round 2, call 7: pack_w v3, bf16-X backward-weights, remaining ATen ops by shape
set -u
out=gpurun_out/r02c7
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_nhwc.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -8 > $out/nhwc_tests.log
timeout 600 python tools/profile_layers.py --rows 40 > $out/profile_layers.txt 2>&1
timeout 600 python tools/profile_aten_ops.py --rows 90 > $out/profile_aten.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline 2>$out/bench.err | tail -1 > $out/bench.json
timeout 900 python tools/parity_report.py ade20k_train > $out/parity.txt 2>&1
tail -4 $out/nhwc_tests.log
head -36 $out/profile_layers.txt
grep -v Warn $out/profile_aten.txt | head -125
cut -c1-220 $out/bench.json; tail -2 $out/bench.err
grep -v "Warn\|line\[" $out/parity.txt | cut -c1-600
