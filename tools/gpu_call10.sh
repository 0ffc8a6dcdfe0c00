#!/bin/bash
# round 2, call 10: stat-SPADE kernels, all goldens on the tape, ncu summaries (computed on the box), bench per config
set -u
out=gpurun_out/r02c10
mkdir -p $out
rm -rf gpurun_out/ncu
timeout 900 python -m pytest tests/test_gpu_nhwc.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 > $out/nhwc_tests.log
timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|warp_out|cosine|Error" | tail -20 > $out/model_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-gpu-baseline 2>$out/bench.err | tail -1 > $out/bench.json
for c in 2 3 4; do
  timeout 900 python bench.py --config $c --no-cpu-baseline --no-gpu-baseline 2>$out/bench_config$c.err | tail -1 > $out/bench_config$c.json
done
bash tools/ncu_kernels.sh > $out/ncu.log 2>&1
mv gpurun_out/ncu/*_ncu_summary.json $out/ 2>/dev/null
tail -4 $out/nhwc_tests.log; cat $out/model_tests.log
for f in bench bench_config2 bench_config3 bench_config4; do cut -c1-200 $out/$f.json; done
tail -2 $out/bench_config2.err
tail -12 $out/ncu.log
du -sh gpurun_out
