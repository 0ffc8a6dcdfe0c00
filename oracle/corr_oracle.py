"""CPU oracle for the CoCosNet correspondence + warp hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may import this module; the product path
(cocosnet_b200/) never does and fails loudly when its CUDA library is missing.

Plain numpy float64 restatement of the reference algorithm.  Every function
cites the reference file:line it follows (paths relative to the reference
root, microsoft/CoCosNet @ de0c1bb).

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference itself, executed on CPU
in the build container by tests/golden/make_golden.py; the resulting vectors
are committed under tests/golden/*.npz and tests/test_oracle_golden.py checks
this file against all of them.
"""
import numpy as np

EPS = 2.220446049250313e-16  # sys.float_info.epsilon, util/util.py:32


# --------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------
def feature_normalize(x):
    """util/util.py:31-34 : x / (||x||_2 over dim 1 + eps)."""
    x = np.asarray(x, np.float64)
    n = np.sqrt((x * x).sum(axis=1, keepdims=True)) + EPS
    return x / n


def unfold(x, k, padding=None, stride=1):
    """torch.nn.functional.unfold on [B,C,H,W] -> [B, C*k*k, L].
    Row order c*k*k + ky*k + kx (correspondence.py:276,286,311)."""
    x = np.asarray(x)
    b, c, h, w = x.shape
    if padding is None:
        padding = 0
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding), (padding, padding)))
    oh = (h + 2 * padding - k) // stride + 1
    ow = (w + 2 * padding - k) // stride + 1
    out = np.empty((b, c, k, k, oh, ow), x.dtype)
    for ky in range(k):
        for kx in range(k):
            out[:, :, ky, kx] = xp[:, :, ky:ky + stride * oh:stride, kx:kx + stride * ow:stride]
    return out.reshape(b, c * k * k, oh * ow)


def fold(cols, out_hw, k, stride):
    """torch.nn.functional.fold for non-overlapping patches
    (correspondence.py:321,356; kernel == stride)."""
    b, ckk, l = cols.shape
    c = ckk // (k * k)
    oh = ow = out_hw // stride
    assert oh * ow == l
    x = cols.reshape(b, c, k, k, oh, ow)
    return x.transpose(0, 1, 4, 2, 5, 3).reshape(b, c, oh * k, ow * k)


def avg_pool(x, k):
    """F.avg_pool2d(x, k) (correspondence.py:313)."""
    b, c, h, w = x.shape
    return np.asarray(x, np.float64).reshape(b, c, h // k, k, w // k, k).mean(axis=(3, 5))


def nearest_down(x, factor):
    """F.interpolate(mode='nearest') to 1/factor size (correspondence.py:258,330)."""
    return x[:, :, ::factor, ::factor]


def upsample_nearest(x, s):
    """nn.Upsample(scale_factor=s) nearest (correspondence.py:188)."""
    return np.repeat(np.repeat(x, s, axis=2), s, axis=3)


def upsample_bilinear(x, s):
    """nn.Upsample(scale_factor=s, mode='bilinear'), align_corners=False
    (correspondence.py:184-186)."""
    b, c, h, w = x.shape

    def axis_weights(n):
        dst = np.arange(n * s)
        src = (dst + 0.5) / s - 0.5
        src = np.maximum(src, 0.0)
        i0 = np.minimum(np.floor(src).astype(np.int64), n - 1)
        i1 = np.minimum(i0 + 1, n - 1)
        l1 = src - i0
        return i0, i1, 1.0 - l1, l1

    y0, y1, wy0, wy1 = axis_weights(h)
    x0, x1, wx0, wx1 = axis_weights(w)
    rows = x[:, :, y0, :] * wy0[None, None, :, None] + x[:, :, y1, :] * wy1[None, None, :, None]
    return rows[:, :, :, x0] * wx0 + rows[:, :, :, x1] * wx1


def center_normalize(f, pono_c):
    """correspondence.py:277-280 / 287-289.  f: [B,K,N].
    mean over dim 1 if PONO_C else over dim -1; then divide by the L2 norm over
    dim 1 plus eps."""
    f = np.asarray(f, np.float64)
    f = f - f.mean(axis=1 if pono_c else -1, keepdims=True)
    n = np.sqrt((f * f).sum(axis=1, keepdims=True)) + EPS
    return f / n


def softmax_rows(z):
    """F.softmax(z, dim=-1) (correspondence.py:307)."""
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


def logsumexp_rows(z):
    m = z.max(axis=-1, keepdims=True)
    return (m + np.log(np.exp(z - m).sum(axis=-1, keepdims=True)))[..., 0]


# --------------------------------------------------------------------------
# the kernel-level primitive:  O = softmax(scale * Q K^T) V
# --------------------------------------------------------------------------
def attend(q, k, v, scale):
    """q:[B,Nq,Kd] k:[B,Nk,Kd] v:[B,Nk,Cv] -> (o:[B,Nq,Cv], lse:[B,Nq]).
    The fused form of correspondence.py:291 (matmul), :304 (/temperature),
    :307 (softmax), :318 (matmul with ref)."""
    q = np.asarray(q, np.float64)
    k = np.asarray(k, np.float64)
    v = np.asarray(v, np.float64)
    z = (q @ k.transpose(0, 2, 1)) * scale
    p = softmax_rows(z)
    return p @ v, logsumexp_rows(z)


def attend_backward(q, k, v, scale, d_o):
    """Analytic gradients of `attend` wrt q, k, v given dO (what autograd does
    through correspondence.py:291-318)."""
    q = np.asarray(q, np.float64)
    k = np.asarray(k, np.float64)
    v = np.asarray(v, np.float64)
    d_o = np.asarray(d_o, np.float64)
    z = (q @ k.transpose(0, 2, 1)) * scale
    p = softmax_rows(z)
    o = np.einsum("bij,bjc->bic", p, v)
    dv = np.einsum("bij,bic->bjc", p, d_o)
    dp = np.einsum("bic,bjc->bij", d_o, v)
    delta = (d_o * o).sum(-1, keepdims=True)
    ds = p * (dp - delta) * scale
    dq = np.einsum("bij,bjk->bik", ds, k)
    dk = np.einsum("bij,bik->bjk", ds, q)
    return dq, dk, dv


# --------------------------------------------------------------------------
# correspondence.py:272-372  (everything after the theta / phi 1x1 convs)
# --------------------------------------------------------------------------
def corr_tail(theta_conv, phi_conv, ref_img, *, match_kernel=3, pono_c=True,
              temperature=0.01, down=4, warp_patch=False, warp_bilinear=False,
              ref_seg_map=None, seg_map=None, real_img=None,
              warp_mask_losstype="none", warp_cycle=False, two_cycle=False,
              return_corr=False):
    """theta_conv/phi_conv: outputs of self.theta / self.phi, [B,C,h,w].
    Returns the dict NoVGGCorrespondence.forward builds (correspondence.py:222-374)."""
    b, c, fh, fw = theta_conv.shape
    out = {}

    def prep(x):  # :272-280 and :282-289
        x = np.asarray(x, np.float64)
        if match_kernel == 1:
            f = x.reshape(b, c, -1)
        else:
            f = unfold(x, match_kernel, padding=match_kernel // 2)
        return center_normalize(f, pono_c)

    theta = prep(theta_conv)
    phi = prep(phi_conv)
    f = theta.transpose(0, 2, 1) @ phi  # :291
    f_wta = f / temperature  # :304
    if return_corr:
        return f_wta  # :305-306
    p = softmax_rows(f_wta)  # :307

    ref_img = np.asarray(ref_img, np.float64)
    if warp_patch:  # :310-316
        ref = unfold(ref_img, down, stride=down)
    else:
        ref = avg_pool(ref_img, down)
        channel = ref.shape[1]
        ref = ref.reshape(b, channel, -1)
    ref = ref.transpose(0, 2, 1)
    y = p @ ref  # :318
    if warp_patch:  # :319-321
        y = fold(y.transpose(0, 2, 1), 256, down, down)
        out["warp_out"] = y
    else:  # :322-327
        y = y.transpose(0, 2, 1).reshape(b, channel, fh, fw)
        out["warp_out"] = upsample_bilinear(y, down) if warp_bilinear else upsample_nearest(y, down)

    p_v = None
    if warp_mask_losstype == "direct":  # :329-336
        rs = nearest_down(np.asarray(ref_seg_map, np.float64), down)
        ch = rs.shape[1]
        rs = rs.reshape(b, ch, -1).transpose(0, 2, 1)
        wm = p @ rs
        out["warp_mask"] = wm.transpose(0, 2, 1).reshape(b, ch, fh, fw)
    elif warp_mask_losstype == "cycle":  # :337-346
        p_v = softmax_rows(f_wta.transpose(0, 2, 1))
        sg = nearest_down(np.asarray(seg_map, np.float64), down)
        ch = sg.shape[1]
        sg = sg.reshape(b, ch, -1).transpose(0, 2, 1)
        to_ref = p_v @ sg
        wm = p @ to_ref
        out["warp_mask"] = wm.transpose(0, 2, 1).reshape(b, ch, fh, fw)

    if warp_cycle:  # :350-372
        if p_v is None:
            p_v = softmax_rows(f_wta.transpose(0, 2, 1))
        if warp_patch:
            yy = unfold(y, down, stride=down).transpose(0, 2, 1)
            wc = (p_v @ yy).transpose(0, 2, 1)
            out["warp_cycle"] = fold(wc, 256, down, down)
        else:
            ch = y.shape[1]
            yy = y.reshape(b, ch, -1).transpose(0, 2, 1)
            wc = (p_v @ yy).transpose(0, 2, 1)
            out["warp_cycle"] = wc.reshape(b, ch, fh, fw)
            if two_cycle:
                ri = avg_pool(np.asarray(real_img, np.float64), down).reshape(b, ch, -1).transpose(0, 2, 1)
                i2r = p_v @ ri
                out["warp_i2r"] = i2r.transpose(0, 2, 1).reshape(b, ch, fh, fw)
                i2r2i = p @ i2r
                out["warp_i2r2i"] = i2r2i.transpose(0, 2, 1).reshape(b, ch, fh, fw)
    return out


# --------------------------------------------------------------------------
# SAGAN attention block, architecture.py:114-127 (weights passed explicitly)
# --------------------------------------------------------------------------
def max_pool2(x):
    b, c, h, w = x.shape
    return x.reshape(b, c, h // 2, 2, w // 2, 2).max(axis=(3, 5))


def sagan_attention(x, w_theta, w_phi, w_g, w_o, gamma):
    """x:[B,ch,H,W]; w_*: 1x1 conv weights as [out,in] matrices (already
    spectrally normalised by the caller)."""
    x = np.asarray(x, np.float64)
    b, ch, h, w = x.shape
    conv = lambda wt, t: np.einsum("oi,bihw->bohw", np.asarray(wt, np.float64), t)
    theta = conv(w_theta, x).reshape(b, ch // 8, h * w)
    phi = max_pool2(conv(w_phi, x)).reshape(b, ch // 8, h * w // 4)
    g = max_pool2(conv(w_g, x)).reshape(b, ch // 2, h * w // 4)
    beta = softmax_rows(np.einsum("bkn,bkm->bnm", theta, phi))
    o = np.einsum("bcm,bnm->bcn", g, beta).reshape(b, ch // 2, h, w)
    return gamma * conv(w_o, o) + x


# --------------------------------------------------------------------------
# PositionalNorm2d + SPADE modulation, normalization.py:63-68,149;
# leaky relu architecture.py:94-95
# --------------------------------------------------------------------------
def positional_norm(x, eps=1e-5):
    x = np.asarray(x, np.float64)
    mean = x.mean(axis=1, keepdims=True)
    var = x.var(axis=1, ddof=1, keepdims=True)
    return (x - mean) / np.sqrt(var + eps)


def spade_modulate(x, gamma, beta, leaky=None):
    out = positional_norm(x) * (1.0 + np.asarray(gamma, np.float64)) + np.asarray(beta, np.float64)
    if leaky is not None:
        out = np.where(out >= 0, out, out * leaky)
    return out


def operand_prologue_backward(x, g, match_kernel=3):
    """Gradient of the --PONO_C operand prologue (unfold -> centre over K -> L2-normalise, correspondence.py:273-289)
    w.r.t. the conv output x [B,C,h,w], given g = dL/d(operand) [B,K,N] in the product's TAP-MAJOR K order
    (k = tap*C + c).  Restates the two passes of cocos_normalize_pack_bwd (per-position a = <fhat, g>, s = sum g;
    then a 9-term gather per pixel) so that the formula -- not just the kernel -- is pinned to autograd of the
    reference expression (tests/test_oracle_golden.py::test_operand_prologue_backward_matches_autograd)."""
    x = np.asarray(x, np.float64)
    g = np.asarray(g, np.float64)
    B, C, h, w = x.shape
    mk, half = match_kernel, match_kernel // 2
    taps = [(r - half, s - half) for r in range(mk) for s in range(mk)]
    K = C * len(taps)
    xp = np.pad(x, ((0, 0), (0, 0), (half, half), (half, half)))
    f = np.stack([xp[:, :, half + di:half + di + h, half + dj:half + dj + w] for di, dj in taps], 1)  # [B,T,C,h,w]
    f = f.reshape(B, K, h * w)
    mean = f.mean(1, keepdims=True)
    inv = 1.0 / (np.sqrt(((f - mean) ** 2).sum(1, keepdims=True)) + EPS)
    fhat = (f - mean) * inv
    a = (fhat * g).sum(1, keepdims=True)
    s = g.sum(1, keepdims=True)
    df = ((g - fhat * a - s / K) * inv).reshape(B, len(taps), C, h, w)
    dx = np.zeros((B, C, h + 2 * half, w + 2 * half))
    for t, (di, dj) in enumerate(taps):  # the fold: position n contributed x[., n + offset(t)]
        dx[:, :, half + di:half + di + h, half + dj:half + dj + w] += df[:, t]
    return dx[:, :, half:half + h, half:half + w]


def contextual_rows(S, h=0.1, eps=1e-3):
    """ContextualLoss.py:117-131 from the correlation matrix S [B,N,N]: cx[b,i] = max_j A_ij, written the way
    cocos_ctx_rows_fwd computes it (the row maximum of A sits at the row minimum of d): 1 / sum_j exp(-a (d - m))."""
    d = 1.0 - np.asarray(S, np.float64)
    m = d.min(-1, keepdims=True)
    a = 1.0 / (h * (m + eps))
    return 1.0 / np.exp(-a * (d - m)).sum(-1)


def contextual_rows_backward(S, g, h=0.1, eps=1e-3):
    """dL/dS given g = dL/dcx [B,N]: the formula of cocos_ctx_rows_bwd (min routed to its first arg-min)."""
    d = 1.0 - np.asarray(S, np.float64)
    B, N, _ = d.shape
    m = d.min(-1, keepdims=True)
    arg = d.argmin(-1)
    a = 1.0 / (h * (m + eps))
    q = np.exp(-a * (d - m))
    sumq = q.sum(-1, keepdims=True)
    cx = 1.0 / sumq
    c = np.asarray(g, np.float64)[..., None] * cx * cx * a
    dd = c * q
    onehot = np.zeros_like(d, dtype=bool)
    np.put_along_axis(onehot, arg[..., None], True, -1)
    sumqd = np.where(onehot, 0.0, q * (d - m)).sum(-1, keepdims=True)
    at_min = c * ((sumq - 1.0) + sumqd / (m + eps))
    dd = np.where(onehot, -at_min, dd)
    return -dd
