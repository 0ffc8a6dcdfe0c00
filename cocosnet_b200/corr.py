"""Dense correspondence + warp on the fused sm_100a kernel (K1).

Host-side mirror of the tail of NoVGGCorrespondence.forward (reference
models/networks/correspondence.py:272-372): everything after the theta / phi
1x1 convs.  The N x N correlation / softmax matrix is never materialised
(except for `return_corr=True`, which the reference API defines as returning
it, correspondence.py:305-306).
"""
import sys

import torch
import torch.nn.functional as F

from . import ops

_EPS = sys.float_info.epsilon  # correspondence.py:279


def _unfold_center_normalize(x, match_kernel, pono_c):
    """correspondence.py:273-280 (theta) / 283-289 (phi): [B,C,h,w] -> [B,K,N]."""
    b, c = x.shape[:2]
    if match_kernel == 1:
        f = x.reshape(b, c, -1)
    else:
        f = F.unfold(x, kernel_size=match_kernel, padding=match_kernel // 2)
    f = f - f.mean(dim=1 if pono_c else -1, keepdim=True)
    return f / (torch.norm(f, 2, 1, keepdim=True) + _EPS)


class _Attend(torch.autograd.Function):
    """softmax(scale * Q K^T) V with Q,K given channel-major [B,Kd,N] fp32 and
    V channel-major [B,Cv,Nk]; output channel-major [B,Cv,Nq]."""

    @staticmethod
    def forward(ctx, q, k, v, scale, precision):
        split_q, split_k = (1, 2) if precision == "split" else (0, 0)
        q16 = ops.pack_rows(q.contiguous(), split=split_q)
        k16 = ops.pack_rows(k.contiguous(), split=split_k)
        v = v.contiguous()
        if v.shape[1] <= 4 and k.shape[2] % 4 == 0:
            # <= 4 value channels (the RGB exemplar): fp32 values straight into the CUDA-core-PV kernel
            out, lse, _ = ops.corr_warp_fwd(q16, k16, None, v.shape[1], k.shape[2], scale, want_lse=True, v32=v)
        else:
            out, lse, _ = ops.corr_warp_fwd(q16, k16, ops.pack_v(v), v.shape[1], k.shape[2], scale, want_lse=True)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.scale = scale
        ctx.precision = precision
        return out

    @staticmethod
    def backward(ctx, d_out):
        from . import corr_bwd
        q, k, v, out, lse = ctx.saved_tensors
        dq, dk, dv = corr_bwd.attend_backward(q, k, v, out, lse, d_out.contiguous(), ctx.scale,
                                              ctx.needs_input_grad[:3])
        return dq, dk, dv, None, None


class Packed:
    """fp16 K-major operand [B,N,Kd] produced by the fused prologue (no autograd graph behind it)."""

    def __init__(self, t):
        self.t = t


def attend(q, k, v, scale, precision="fp16"):
    """q [B,Kd,Nq], k [B,Kd,Nk], v [B,Cv,Nk] fp32 CUDA -> [B,Cv,Nq].  q / k may also be `Packed` operands
    (inference path): forward only."""
    if isinstance(q, Packed):
        v = v.contiguous()
        nk = k.t.shape[1]
        if v.shape[1] <= 4 and nk % 4 == 0:
            return ops.corr_warp_fwd(q.t, k.t, None, v.shape[1], nk, scale, want_lse=False, v32=v)[0]
        return ops.corr_warp_fwd(q.t, k.t, ops.pack_v(v), v.shape[1], nk, scale, want_lse=False)[0]
    return _Attend.apply(q, k, v, float(scale), precision)


def raw_correlation(q, k, scale):
    """`return_corr=True` path (correspondence.py:305-306): scaled logits [B,Nq,Nk]."""
    q16 = q.t if isinstance(q, Packed) else ops.pack_rows(q.contiguous())
    k16 = k.t if isinstance(k, Packed) else ops.pack_rows(k.contiguous())
    return ops.gemm_f16(q16, k16, alpha=scale)


def _operands(x, match_kernel, pono_c, precision):
    """theta / phi conv output -> normalised correlation operand.  Without autograd (inference) and with
    --PONO_C the whole prologue (unfold, centre, normalise, fp16 pack) is one fused kernel pair."""
    fused = (not (torch.is_grad_enabled() and x.requires_grad)) and pono_c and precision == "fp16" \
        and match_kernel in (1, 3) and x.is_cuda and x.dtype == torch.float32 and x.shape[1] % 4 == 0 \
        and (x.shape[1] * match_kernel * match_kernel) % 64 == 0
    if fused:
        return Packed(ops.normalize_pack(x, match_kernel, _EPS))
    return _unfold_center_normalize(x, match_kernel, pono_c)


def correspondence_tail(theta_conv, phi_conv, ref_img, *, match_kernel=3, pono_c=True, temperature=0.01, down=4,
                        warp_patch=False, ref_seg_map=None, seg_map=None, real_img=None,
                        warp_mask_losstype="none", show_warpmask=False, warp_cycle=False, two_cycle=False,
                        return_corr=False, precision="fp16"):
    """Returns (y, extras): y = warped exemplar before the final upsample,
    [B,3,h,w] (or folded [B,3,256,256] for warp_patch); extras holds
    warp_mask / warp_cycle / warp_i2r / warp_i2r2i when requested."""
    b, _, fh, fw = theta_conv.shape
    theta = _operands(theta_conv, match_kernel, pono_c, precision)
    phi = _operands(phi_conv, match_kernel, pono_c, precision)
    if isinstance(theta, Packed) != isinstance(phi, Packed):  # keep the two operands in the same K order
        theta = _unfold_center_normalize(theta_conv, match_kernel, pono_c)
        phi = _unfold_center_normalize(phi_conv, match_kernel, pono_c)
    scale = 1.0 / temperature
    if return_corr:
        return raw_correlation(theta, phi, scale), {}

    if warp_patch:  # :310-311
        ref = F.unfold(ref_img, down, stride=down)
    else:  # :313-315
        ref = F.avg_pool2d(ref_img, down).reshape(b, ref_img.shape[1], -1)
    channel = ref.shape[1]
    extras = {}
    values = [ref]
    want_direct = warp_mask_losstype == "direct" or show_warpmask
    if want_direct:  # :329-336 rides in the same pass: V = [rgb | seg]
        ref_seg = F.interpolate(ref_seg_map, scale_factor=1 / down, mode="nearest")
        values.append(ref_seg.reshape(b, ref_seg.shape[1], -1))
    y_all = attend(theta, phi, torch.cat(values, 1) if len(values) > 1 else ref, scale, precision)
    y = y_all[:, :channel].contiguous() if len(values) > 1 else y_all
    if want_direct:
        extras["warp_mask"] = y_all[:, channel:].contiguous().reshape(b, -1, fh, fw)

    # column softmax == the same primitive with the operands swapped
    if warp_mask_losstype == "cycle" and not want_direct:  # :337-346
        seg = F.interpolate(seg_map, scale_factor=1 / down, mode="nearest")
        to_ref = attend(phi, theta, seg.reshape(b, seg.shape[1], -1), scale, precision)
        extras["warp_mask"] = attend(theta, phi, to_ref, scale, precision).reshape(b, -1, fh, fw)
    if warp_cycle:  # :350-372
        if warp_patch:
            y_img = F.fold(y, 256, down, stride=down)
            yy = F.unfold(y_img, down, stride=down)
            wc = attend(phi, theta, yy, scale, precision)
            extras["warp_cycle"] = F.fold(wc, 256, down, stride=down)
        else:
            extras["warp_cycle"] = attend(phi, theta, y, scale, precision).reshape(b, channel, fh, fw)
            if two_cycle:
                ri = F.avg_pool2d(real_img, down).reshape(b, channel, -1)
                i2r = attend(phi, theta, ri, scale, precision)
                extras["warp_i2r"] = i2r.reshape(b, channel, fh, fw)
                extras["warp_i2r2i"] = attend(theta, phi, i2r, scale, precision).reshape(b, channel, fh, fw)
    if warp_patch:  # :319-321
        y = F.fold(y, 256, down, stride=down)
    else:  # :323-324
        y = y.reshape(b, channel, fh, fw)
    return y, extras
