#!/bin/bash
# round 2, call 8: full -m gpu suite, the complete bench line (gpu_baseline + cpu_baseline), kernel table
set -u
out=gpurun_out/r02c8
mkdir -p $out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -25 > $out/gpu_suite.log
timeout 1500 python bench.py 2>$out/bench.err | tail -1 > $out/bench.json
COCOS_CUDA_GRAPH=0 timeout 600 python tools/profile_step.py --b 8 --cudnn_benchmark --rows 60 --no_table > $out/profile_step_eager.txt 2>&1
tail -12 $out/gpu_suite.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c8/bench.json"))
for k in ("value", "ms_per_step", "e2e", "gpu_launches", "step_roofline", "gpu_baseline", "cpu_baseline", "clocks"):
    print(k, d.get(k))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "ms_per_launch")}, {k: d["roofline_k2304"][k] for k in ("achieved", "frac", "ms_per_launch")})
PY
tail -3 $out/bench.err
head -40 $out/profile_step_eager.txt | cut -c1-170
