#!/bin/bash
# Build libcocos_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libcocos_b200.so
SRCS="api.cu tmap.cu corr_fwd.cu corr_fwd2.cu corr_fwd3.cu corr_fwd4.cu gemm.cu pack.cu $(ls corr_bwd.cu spade_mod.cu inst_act.cu norm_pack.cu conv.cu conv_wgrad.cu tapconv.cu tapwgrad.cu ew_nhwc.cu comm.cu 2>/dev/null || true)"
$NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -shared ${COCOS_NVCC_EXTRA} -o $OUT $SRCS -lcudart
echo "built $(readlink -f $OUT)"
