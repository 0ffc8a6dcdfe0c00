#!/bin/bash
# round 2, call 13: pack cache + unpack fast path + smoke + layer profile + bench
set -u
out=gpurun_out/r02c13
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_nhwc.py -q -m gpu -p no:cacheprovider -k "cache or unpack" 2>&1 | tail -6 > $out/new_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|warp_out|cosine|Error" | tail -14 > $out/model_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-gpu-baseline 2>$out/bench.err | tail -1 > $out/bench.json
COCOS_PACK_CACHE=0 timeout 900 python bench.py --no-cpu-baseline --no-gpu-baseline 2>/dev/null | tail -1 > $out/bench_nocache.json
timeout 600 python tools/profile_layers.py --rows 50 > $out/profile_layers.txt 2>&1
tail -3 $out/new_tests.log; tail -2 $out/smoke.log; cat $out/model_tests.log
cut -c1-200 $out/bench.json; cut -c1-200 $out/bench_nocache.json; tail -2 $out/bench.err
head -70 $out/profile_layers.txt
