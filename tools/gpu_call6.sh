#!/bin/bash
# round 2, call 6: rewritten instance-norm / pack_w kernels, wgrad with bf16 X, default-mode parity tests, layer profile
set -u
out=gpurun_out/r02c6
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_nhwc.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15 > $out/nhwc_tests.log
timeout 300 python tools/bench_tapconv.py > $out/bench_tapconv.txt 2>&1
timeout 600 python tools/profile_layers.py --rows 50 > $out/profile_layers.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline 2>$out/bench.err | tail -1 > $out/bench.json
timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -v Warning | tail -40 > $out/model_tests.log
tail -6 $out/nhwc_tests.log
cat $out/bench_tapconv.txt
head -45 $out/profile_layers.txt
cut -c1-220 $out/bench.json; tail -2 $out/bench.err
tail -30 $out/model_tests.log
