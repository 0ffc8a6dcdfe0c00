"""Backward of the fused correspondence primitive (what autograd computes
through reference correspondence.py:291-318), on the sm_100a kernels:
stage A (csrc/corr_bwd.cu) emits dS / dS^T / P^T in fp16, three tcgen05 GEMMs
(csrc/gemm.cu) contract them with the channel-major operands."""
import torch

from . import ops


def attend_backward(q, k, v, out, lse, d_out, scale, needs):
    """q [B,Kd,Nq], k [B,Kd,Nk], v [B,Cv,Nk], out/d_out [B,Cv,Nq] fp32 CUDA.
    Returns (dq, dk, dv) channel-major like the inputs (None where not needed)."""
    need_q, need_k, need_v = needs
    b, kd, nq = q.shape
    nk = k.shape[2]
    cv = v.shape[1]
    cvk = ops.round_up(cv, 64)
    q16 = ops.pack_rows(q.contiguous())
    k16 = ops.pack_rows(k.contiguous())
    do16 = ops.pack_rows(d_out, kp=cvk)
    v16 = ops.pack_rows(v.contiguous(), kp=cvk)
    # keep dS inside fp16 range: |dS| <= scale * |P| * |dP - D| ~ scale * 2*|dO|_max*|V|_max*min(cv, 8)
    bound = float(d_out.abs().amax()) * max(float(v.abs().amax()), 1e-30) * 2.0 * min(cv, 8) * scale
    dscale = 1.0 / bound if bound > 0 else 1.0
    ds, dst, pt = ops.corr_warp_bwd_ds(q16, k16, do16, v16, d_out, out, lse, cv, scale, dscale, need_v)
    dq = dk = dv = None
    if need_q:
        dq = ops.gemm_f16(ops.cast_rows_f16(k.contiguous()), ds, alpha=1.0 / dscale)      # [B,Kd,Nq]
    if need_k:
        dk = ops.gemm_f16(ops.cast_rows_f16(q.contiguous()), dst, alpha=1.0 / dscale)     # [B,Kd,Nk]
    if need_v:
        dv = ops.gemm_f16(ops.cast_rows_f16(d_out), pt, alpha=1.0)                         # [B,Cv,Nk]
    return dq, dk, dv
