#!/bin/bash
# round 2, call 2: bring-up of the NHWC kernels, one pytest process per kernel family (a trap poisons the context)
set -u
out=gpurun_out/r02c2
mkdir -p $out
T="timeout 600 python -m pytest tests/test_gpu_nhwc.py -q -m gpu --timeout 120 -p no:cacheprovider"
$T -k "pack_unpack or spade_mod or inst_act" 2>&1 | tail -40 > $out/ew.log
$T -k "tapconv_forward or residual" 2>&1 | tail -60 > $out/fwd.log
$T -k "backward_data" 2>&1 | tail -40 > $out/dgrad.log
$T -k "tapwgrad" 2>&1 | tail -60 > $out/wgrad.log
if grep -q " passed" $out/fwd.log && ! grep -q "failed" $out/fwd.log; then
  timeout 300 python tools/bench_tapconv.py > $out/bench_tapconv.txt 2>&1
fi
tail -4 $out/ew.log $out/fwd.log $out/dgrad.log $out/wgrad.log
cat $out/bench_tapconv.txt 2>/dev/null | tail -14
