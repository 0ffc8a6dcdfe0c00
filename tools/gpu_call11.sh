#!/bin/bash
# round 2, call 11: contextual-loss / pair-loss kernels, fused losses in the model, where celebahq loses precision
set -u
out=gpurun_out/r02c11
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_nhwc.py tests/test_gpu_corr.py -q -m gpu -p no:cacheprovider -k "pair_loss or contextual" 2>&1 | tail -8 > $out/new_tests.log
timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|warp_out|cosine|Error|assert" | tail -20 > $out/model_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-gpu-baseline 2>$out/bench.err | tail -1 > $out/bench.json
COCOS_FUSED_LOSSES=0 timeout 900 python bench.py --no-cpu-baseline --no-gpu-baseline 2>/dev/null | tail -1 > $out/bench_unfused_losses.json
timeout 600 python tools/parity_trace.py celebahq_train > $out/trace_celebahq.txt 2>&1
timeout 600 python tools/parity_trace.py ade20k_train > $out/trace_ade20k.txt 2>&1
timeout 600 python tools/profile_aten_ops.py --rows 30 > $out/profile_aten.txt 2>&1
tail -4 $out/new_tests.log; cat $out/model_tests.log
cut -c1-200 $out/bench.json; cut -c1-200 $out/bench_unfused_losses.json; tail -2 $out/bench.err
grep "rel L2\|captured\|Error" $out/trace_celebahq.txt $out/trace_ade20k.txt
grep -v Warn $out/profile_aten.txt | head -40
