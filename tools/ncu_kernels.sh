#!/bin/bash
# ncu --set full captures of the step's main kernels on step shapes (one family per process); the .ncu-rep files come
# back in gpurun_out/ncu/ and are summarised with tools/ncu_summary.py into profiles/r02_*_ncu_summary.json
set -u
out=gpurun_out/ncu
mkdir -p $out
cap() {  # family, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $out/$1 \
    python tools/run_kernels_once.py $1 > $out/$1.log 2>&1
  tail -2 $out/$1.log
}
cap tapconv  "tapconv_kernel"                      2 2
cap tapwgrad "tapwgrad_kernel|cast_op_bf16"        2 2
cap spade    "spade_mod_nhwc"                      2 4
cap inst     "inst_act_nhwc|in_stats_nhwc"         4 8
cap corr_bwd "corr_bwd_ds|norm_pack|gemm_f16|corr_fwd4|transpose_f16" 8 16
cap pack     "nhwc_pack|nhwc_unpack|cast_op"       3 6
# summarise on the box and keep the reports small enough to travel back (gpurun_out is capped at 64 MiB)
for f in tapconv tapwgrad spade inst corr_bwd pack; do
  python tools/ncu_summary.py $out/$f.ncu-rep > $out/${f}_ncu_summary.json 2>$out/${f}_summary.err
done
rm -f $out/corr_bwd.ncu-rep $out/inst.ncu-rep $out/pack.ncu-rep $out/spade.ncu-rep
ls -la $out
