#!/bin/bash
# One gpurun call that answers the open questions of the previous round (run from the repo root on the GPU box):
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# Everything lands under gpurun_out/next/.
set -u
out=gpurun_out/next
mkdir -p $out
# 1. parity: whole GPU suite, then the experimental K2w flat mode on its own (a failure there must not hide the rest)
python -m pytest tests -q -m gpu 2>&1 | tail -8 > $out/pytest_gpu.log
COCOS_WGRAD_NARROW=1 timeout 300 python -m pytest tests/test_gpu_corr.py -q -m gpu -k "flat_mode" 2>&1 | tail -8 > $out/pytest_flat_mode.log
# 2. the bench line (default path) and the eager one
python bench.py 2>&1 | tail -1 > $out/bench_default.json
COCOS_CUDA_GRAPH=0 python bench.py --no-cpu-baseline 2>&1 | tail -1 > $out/bench_eager.json
# 3. is the flat-mode K2w worth routing?  (only meaningful if pytest_flat_mode.log is green)
COCOS_WGRAD_NARROW=1 python bench.py --no-cpu-baseline 2>&1 | tail -1 > $out/bench_wgrad_narrow.json
# 4. where the iteration spends its GPU time now (eager, per-op table; "GPU busy" double counts: halve it)
COCOS_CUDA_GRAPH=0 python tools/profile_step.py --b 8 --cudnn_benchmark --rows 70 > $out/profile_step_table.txt 2>&1
# 5. (opt-in: `bash tools/gpu_first_call.sh ncu`; the last one cost 19 GPU-minutes) launch list of the eager bench
if [ "${1:-}" = "ncu" ]; then
  COCOS_CUDA_GRAPH=0 BENCH_PROFILE=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv \
      --log-file $out/launches.csv python bench.py --steps 1 --warmup 1 > $out/bench_under_ncu.log 2>&1
  python tools/ncu_launch_summary.py $out/launches.csv > $out/launches_summary.txt 2>&1
  # one full capture each of K2 forward and K2w (none exists yet): tensor-pipe share, DRAM bytes vs algorithmic
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel -s 3 -c 1 \
      -o $out/k2_fwd python tools/bench_conv.py > $out/ncu_k2_fwd.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 1 -c 1 \
      -o $out/k2_wgrad python tools/debug_wgrad.py 8 512 512 64 64 3 1 1 > $out/ncu_k2_wgrad.log 2>&1
fi
tail -3 $out/pytest_gpu.log $out/pytest_flat_mode.log
cut -c1-300 $out/bench_default.json
