"""CPU: pin oracle/corr_oracle.py against the golden vectors minted from the
reference itself (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import corr_oracle as oc
from tests.golden import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    return np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b) + 1e-30)


@pytest.mark.parametrize("name", list(cases.TAIL_CASES))
def test_tail_matches_reference(name):
    spec = cases.TAIL_CASES[name]
    gold = np.load(os.path.join(GOLD, "tail_%s.npz" % name))
    inp = cases.tail_inputs(name)
    out = oc.corr_tail(inp["theta"], inp["phi"], inp["ref_img"], ref_seg_map=inp["ref_seg"],
                       seg_map=inp["seg"], real_img=inp["real_img"], **spec["flags"])
    # reference computes in fp32: a 4096-term softmax at logits up to +-100
    # carries ~1e-5 relative noise; the oracle is fp64.
    for k in ("warp_out", "warp_cycle", "warp_i2r", "warp_i2r2i"):
        if k in gold.files:
            assert out[k].shape == gold[k].shape
            assert _rel(gold[k], out[k]) < 2e-5, k
    if "warp_mask" in gold.files:
        wm = out["warp_mask"]
        assert _rel(gold["warp_mask"], wm[:, ::19]) < 2e-5
        assert _rel(gold["warp_mask_chsum"], wm.sum(axis=1)) < 2e-5
    if "corr" in gold.files:
        corr = oc.corr_tail(inp["theta"], inp["phi"], inp["ref_img"], return_corr=True,
                            **{k: v for k, v in spec["flags"].items() if k in ("match_kernel", "pono_c")})
        assert np.abs(gold["corr"] - corr).max() < 1e-3  # logits in [-100, 100] (1e-5 of range), fp32 reference


def test_attend_backward_matches_autograd():
    import torch
    rng = np.random.default_rng(5)
    q = oc.feature_normalize(rng.standard_normal((2, 24, 16)).transpose(0, 2, 1)).transpose(0, 2, 1)
    k = oc.feature_normalize(rng.standard_normal((2, 40, 16)).transpose(0, 2, 1)).transpose(0, 2, 1)
    v = rng.standard_normal((2, 40, 5))
    d_o = rng.standard_normal((2, 24, 5))
    tq, tk, tv = (torch.tensor(a, requires_grad=True) for a in (q, k, v))
    o = torch.softmax(tq @ tk.transpose(1, 2) * 100.0, -1) @ tv
    o.backward(torch.tensor(d_o))
    o2, lse = oc.attend(q, k, v, 100.0)
    assert np.allclose(o2, o.detach().numpy(), atol=1e-12)
    dq, dk, dv = oc.attend_backward(q, k, v, 100.0, d_o)
    assert np.allclose(dq, tq.grad.numpy(), atol=1e-9)
    assert np.allclose(dk, tk.grad.numpy(), atol=1e-9)
    assert np.allclose(dv, tv.grad.numpy(), atol=1e-9)
    z = torch.tensor(q) @ torch.tensor(k).transpose(1, 2) * 100.0
    assert np.allclose(lse, torch.logsumexp(z, -1).numpy(), atol=1e-10)


def test_bilinear_and_unfold_match_torch():
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(6)
    x = rng.standard_normal((2, 3, 8, 8))
    assert np.allclose(oc.upsample_bilinear(x, 4),
                       F.interpolate(torch.tensor(x), scale_factor=4, mode="bilinear").numpy(), atol=1e-12)
    assert np.allclose(oc.unfold(x, 3, padding=1), F.unfold(torch.tensor(x), 3, padding=1).numpy())
    u = oc.unfold(x, 4, stride=4)
    assert np.allclose(u, F.unfold(torch.tensor(x), 4, stride=4).numpy())
    assert np.allclose(oc.fold(u, 8, 4, 4), x)
    assert np.allclose(oc.positional_norm(x), ((torch.tensor(x) - torch.tensor(x).mean(1, keepdim=True))
                       / torch.tensor(x).var(1, keepdim=True).add(1e-5).sqrt()).numpy(), atol=1e-12)


@pytest.mark.parametrize("mk", [1, 3])
def test_operand_prologue_backward_matches_autograd(mk):
    """The two-pass backward formula of the fused operand prologue (cocos_normalize_pack_bwd) == autograd through the
    reference's unfold / centre / normalise expressions (correspondence.py:273-289, --PONO_C)."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(5)
    B, C, h, w = 2, 8, 6, 5
    x = rng.standard_normal((B, C, h, w))
    g_ref_order = rng.standard_normal((B, C * mk * mk, h * w))  # F.unfold order: k = c*mk*mk + tap
    xt = torch.tensor(x, requires_grad=True)
    f = F.unfold(xt, kernel_size=mk, padding=mk // 2) if mk > 1 else xt.reshape(B, C, -1)
    f = f - f.mean(dim=1, keepdim=True)
    f = f / (torch.norm(f, 2, 1, keepdim=True) + oc.EPS)
    (f * torch.tensor(g_ref_order)).sum().backward()
    g_tap_major = g_ref_order.reshape(B, C, mk * mk, h * w).transpose(0, 2, 1, 3).reshape(B, C * mk * mk, h * w)
    got = oc.operand_prologue_backward(x, g_tap_major, mk)
    assert np.allclose(got, xt.grad.numpy(), rtol=1e-9, atol=1e-12)


def test_contextual_rows_match_reference_expression_and_autograd():
    """The per-row form of the contextual loss (cocos_ctx_rows_fwd / _bwd) == ContextualLoss.py:117-131 and autograd
    through it, on a random correlation matrix."""
    import torch
    rng = np.random.default_rng(3)
    S = rng.uniform(-0.4, 0.9, (2, 37, 37))
    g = rng.standard_normal((2, 37))
    St = torch.tensor(S, requires_grad=True)
    d = 1 - St
    d_norm = d / (torch.min(d, dim=-1, keepdim=True)[0] + 1e-3)
    w = torch.exp((1 - d_norm) / 0.1)
    A = w / torch.sum(w, dim=-1, keepdim=True)
    cx = torch.max(A, dim=-1)[0]
    (cx * torch.tensor(g)).sum().backward()
    assert np.allclose(oc.contextual_rows(S), cx.detach().numpy(), rtol=1e-10)
    assert np.allclose(oc.contextual_rows_backward(S, g), St.grad.numpy(), rtol=1e-8, atol=1e-12)
