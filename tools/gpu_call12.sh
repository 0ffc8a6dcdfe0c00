#!/bin/bash
# round 2, call 12: full GPU suite, 2-CTA GEMM testbed, variance conditioning, final-candidate bench line
set -u
out=gpurun_out/r02c12
mkdir -p $out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -12 > $out/gpu_suite.log
timeout 900 python tools/bench_gemm.py > $out/bench_gemm_2cta.txt 2>&1
timeout 600 python tools/parity_trace.py celebahq_train > $out/trace_celebahq.txt 2>&1
timeout 1500 python bench.py 2>$out/bench.err | tail -1 > $out/bench.json
timeout 600 python tools/profile_layers.py --rows 45 > $out/profile_layers.txt 2>&1
tail -6 $out/gpu_suite.log
cat $out/bench_gemm_2cta.txt | cut -c1-300
grep -E "rel L2|cond|InstanceNorm" $out/trace_celebahq.txt | head -30
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c12/bench.json"))
for k in ("value", "ms_per_step", "e2e", "gpu_launches", "step_roofline", "gpu_baseline", "cpu_baseline", "clocks"):
    print(k, d.get(k))
PY
tail -2 $out/bench.err
head -40 $out/profile_layers.txt
