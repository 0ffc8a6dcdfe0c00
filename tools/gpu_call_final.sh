#!/bin/bash
set -u
out=gpurun_out/r02final
mkdir -p $out
timeout 170 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -8 > $out/gpu_suite.log
timeout 100 python bench.py --no-cpu-baseline --no-gpu-baseline 2>$out/bench.err | tail -1 > $out/bench.json
tail -5 $out/gpu_suite.log; cut -c1-220 $out/bench.json; tail -2 $out/bench.err | cut -c1-200
