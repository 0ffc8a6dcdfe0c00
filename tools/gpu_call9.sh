#!/bin/bash
# round 2, call 9: multi-layer spectral norm kernel, fused SPADE epilogue tests, bench, ncu captures of the main kernels
set -u
out=gpurun_out/r02c9
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_nhwc.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 > $out/nhwc_tests.log
timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|warp_out|cosine|Error" | tail -20 > $out/model_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-gpu-baseline 2>$out/bench.err | tail -1 > $out/bench.json
timeout 600 python tools/profile_aten_ops.py --rows 40 > $out/profile_aten.txt 2>&1
bash tools/ncu_kernels.sh > $out/ncu.log 2>&1
tail -4 $out/nhwc_tests.log; cat $out/model_tests.log
cut -c1-200 $out/bench.json; tail -2 $out/bench.err
grep -v Warn $out/profile_aten.txt | head -45
tail -15 $out/ncu.log
