#!/bin/bash
# round 2, call 5: fused training prologue (fwd + bwd) tests, per-layer profile, bench
set -u
out=gpurun_out/r02c5
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_corr.py -q -m gpu -k "normalize_pack or fused_prologue or tail" -p no:cacheprovider 2>&1 | tail -25 > $out/prologue_tests.log
timeout 600 python tools/profile_layers.py --rows 90 > $out/profile_layers.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline 2>$out/bench.err | tail -1 > $out/bench.json
COCOS_FUSED_PROLOGUE=0 timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > $out/bench_unfused_prologue.json
timeout 600 python tools/parity_report.py ade20k_train ade20k_infer_mk3 > $out/parity.txt 2>&1
tail -12 $out/prologue_tests.log
head -60 $out/profile_layers.txt
cut -c1-220 $out/bench.json; cut -c1-220 $out/bench_unfused_prologue.json; tail -2 $out/bench.err
grep -v "Warn\|line\[" $out/parity.txt | cut -c1-400
