"""torch-tensor wrappers over the C-ABI (device pointers + current stream).

PyTorch is plumbing here: it owns device memory and the stream; all math on
the hot path runs in the hand-written sm_100a kernels of csrc/.
"""
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _req(t, dtype, name):
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise _lib.CocosError("%s must be a contiguous CUDA %s tensor" % (name, dtype))


def round_up(x, m):
    return (x + m - 1) // m * m


def pack_rows(x, kp=None, split=0):
    """fp32 [B,C,N] -> fp16 [B,N,Kt] (see cocos_pack_rows_f16)."""
    _req(x, torch.float32, "x")
    b, c, n = x.shape
    kp = round_up(c, 64) if kp is None else kp
    kt = kp * (3 if split else 1)
    out = torch.empty((b, n, kt), dtype=torch.float16, device=x.device)
    _lib.check(_lib.lib().cocos_pack_rows_f16(x.data_ptr(), out.data_ptr(), b, c, n, kp, split, _stream()),
               "cocos_pack_rows_f16")
    return out


def pack_v(v):
    """fp32 [B,Cv,Nk] -> fp16 [B,Cvp,Nkp] zero padded (see cocos_pack_v_f16)."""
    _req(v, torch.float32, "v")
    b, cv, nk = v.shape
    cvp, nkp = round_up(cv, 16), round_up(nk, 8)
    out = torch.empty((b, cvp, nkp), dtype=torch.float16, device=v.device)
    _lib.check(_lib.lib().cocos_pack_v_f16(v.data_ptr(), out.data_ptr(), b, cv, nk, cvp, nkp, _stream()),
               "cocos_pack_v_f16")
    return out


def corr_warp_fwd(q16, k16, vt16, cv, nk, scale, want_lse=True, want_corr=False):
    """K1 forward.  q16 [B,Nq,Kd], k16 [B,Nk,Kd], vt16 [B,Cvp,Nkp] fp16.
    Returns (out [B,cv,Nq] fp32, lse [B,Nq] | None, corr [B,Nq,Nk] | None)."""
    for t, nme in ((q16, "q16"), (k16, "k16"), (vt16, "vt16")):
        _req(t, torch.float16, nme)
    b, nq, kd = q16.shape
    if k16.shape[0] != b or k16.shape[2] != kd or k16.shape[1] != nk or vt16.shape[0] != b:
        raise _lib.CocosError("corr_warp_fwd: inconsistent shapes %s %s %s" % (q16.shape, k16.shape, vt16.shape))
    cvp, nkp = vt16.shape[1], vt16.shape[2]
    out = torch.empty((b, cv, nq), dtype=torch.float32, device=q16.device)
    lse = torch.empty((b, nq), dtype=torch.float32, device=q16.device) if want_lse else None
    corr = torch.empty((b, nq, nk), dtype=torch.float32, device=q16.device) if want_corr else None
    _lib.check(_lib.lib().cocos_corr_warp_fwd(q16.data_ptr(), k16.data_ptr(), vt16.data_ptr(), out.data_ptr(),
                                              _ptr(lse), _ptr(corr), b, nq, nk, kd, cv, cvp, nkp, float(scale),
                                              _stream()), "cocos_corr_warp_fwd")
    return out, lse, corr


def gemm_f16(a16, b16, alpha=1.0, out=None, accumulate=False):
    """C[b] = alpha * A[b] @ B[b]^T.  a16 [b,M,K], b16 [b,N,K] fp16 -> fp32 [b,M,N]."""
    _req(a16, torch.float16, "a16")
    _req(b16, torch.float16, "b16")
    bt, m, k = a16.shape
    n = b16.shape[1]
    if b16.shape[0] != bt or b16.shape[2] != k:
        raise _lib.CocosError("gemm_f16: inconsistent shapes %s %s" % (a16.shape, b16.shape))
    if out is None:
        out = torch.empty((bt, m, n), dtype=torch.float32, device=a16.device)
    else:
        _req(out, torch.float32, "out")
    _lib.check(_lib.lib().cocos_gemm_f16(a16.data_ptr(), b16.data_ptr(), out.data_ptr(), bt, m, n, k, k, k, n,
                                         m * k, n * k, m * n, float(alpha), int(bool(accumulate)), _stream()),
               "cocos_gemm_f16")
    return out
