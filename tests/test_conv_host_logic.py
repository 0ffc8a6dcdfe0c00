"""CPU: the host-side operand re-layouts of K2 (cocosnet_b200/ops.py) against torch's own convolution.

The CUDA kernel computes, for NHWC x16 [B,Hin,Win,Cp] and wt [Cout, KS*KS*Cp],
    y[b,n,h,w] = sum_{r,s,c} x16[b, h+r-off, w+s-off, c] * wt[n, (r*KS+s)*Cp + c]        (zero outside the image)
(include/cocos_b200.h).  Here that formula is evaluated with plain torch on the CPU from the SAME packed weights and
`off` values the GPU path uses, for the forward and for the backward-data re-use of the kernel (flipped W^T), and for
the [tap][Cin][Cout] -> [Cout][Cin][KS][KS] permutation of the backward-weights workspace."""
import pytest

torch = pytest.importorskip("torch")
import torch.nn.functional as F  # noqa: E402

from cocosnet_b200 import ops  # noqa: E402


def _kernel_formula(x_nhwc, wt, b, h, w, cp, cout, ks, off):
    """Reference evaluation of the documented K2 index formula (fp64, zero fill outside the input)."""
    hin, win = x_nhwc.shape[1:3]
    lo, hi = off, ks - 1 - off + max(h - hin, 0) + max(w - win, 0) + ks  # generous zero halo
    xp = F.pad(x_nhwc, (0, 0, lo, hi, lo, hi))
    y = torch.zeros(b, cout, h, w, dtype=torch.float64)
    for r in range(ks):
        for s in range(ks):
            patch = xp[:, r:r + h, s:s + w, :]                                   # x[b, h+r-off, w+s-off, c]
            wtap = wt[:, (r * ks + s) * cp:(r * ks + s + 1) * cp]               # [Cout, Cp]
            y += torch.einsum("bhwc,nc->bnhw", patch, wtap)
    return y


@pytest.mark.parametrize("cin,cout,ks,pre_padded", [(5, 7, 3, True), (70, 3, 3, False), (6, 4, 1, False)])
def test_forward_weight_layout_and_offset(cin, cout, ks, pre_padded):
    torch.manual_seed(cin)
    b, h, w = 2, 6, 5
    pad = ks // 2
    hin, win = (h + 2 * pad, w + 2 * pad) if pre_padded else (h, w)
    x = torch.randn(b, cin, hin, win, dtype=torch.float64)
    wgt = torch.randn(cout, cin, ks, ks, dtype=torch.float64)
    cp = ops.round_up(cin, 64)
    wt = ops.pack_conv_weight(wgt, dtype=torch.float64)
    assert wt.shape == (cout, ks * ks * cp)
    x_nhwc = F.pad(x.permute(0, 2, 3, 1), (0, cp - cin))
    off = 0 if pre_padded else pad          # conv_fwd_native
    y = _kernel_formula(x_nhwc, wt, b, h, w, cp, cout, ks, off)
    ref = F.conv2d(x, wgt, None, padding=0 if pre_padded else pad)
    assert torch.allclose(y, ref, atol=1e-10)


@pytest.mark.parametrize("cin,cout,ks,pre_padded", [(5, 7, 3, True), (9, 70, 3, False), (6, 4, 1, False)])
def test_backward_data_is_the_forward_kernel_on_flipped_transposed_weights(cin, cout, ks, pre_padded):
    torch.manual_seed(cout)
    b, h, w = 2, 6, 5
    pad = ks // 2
    hin, win = (h + 2 * pad, w + 2 * pad) if pre_padded else (h, w)
    x = torch.randn(b, cin, hin, win, dtype=torch.float64, requires_grad=True)
    wgt = torch.randn(cout, cin, ks, ks, dtype=torch.float64)
    dy = torch.randn(b, cout, h, w, dtype=torch.float64)
    F.conv2d(x, wgt, None, padding=0 if pre_padded else pad).backward(dy)
    # what conv_dgrad_native hands to the kernel
    cp = ops.round_up(cout, 64)
    wt = ops.pack_conv_weight(wgt.flip(2, 3).transpose(0, 1), dtype=torch.float64)   # [Cin, KS*KS*Cout_p]
    off = (ks - 1) if pre_padded else (ks - 1 - ks // 2)
    dy_nhwc = F.pad(dy.permute(0, 2, 3, 1), (0, cp - cout))
    dx = _kernel_formula(dy_nhwc, wt, b, hin, win, cp, cin, ks, off)
    assert torch.allclose(dx, x.grad, atol=1e-10)


def test_wgrad_workspace_permutation():
    """ws[(r*KS+s), c, n] (what cocos_conv_wgrad writes) -> dW[n, c, r, s]."""
    ks, cin, cout = 3, 4, 5
    dw = torch.arange(cout * cin * ks * ks, dtype=torch.float32).view(cout, cin, ks, ks)
    ws = torch.empty(ks * ks, cin, cout)
    for r in range(ks):
        for s in range(ks):
            ws[r * ks + s] = dw[:, :, r, s].t()
    back = ws.view(ks, ks, cin, cout).permute(3, 2, 0, 1).contiguous()   # conv_wgrad_native's last line
    assert torch.equal(back, dw)
