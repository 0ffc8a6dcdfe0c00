"""Network forwards on the 16-bit NHWC tape (cocosnet_b200/tape.py): the SPADE generator, the domain adaptors, the
residual-block stack with the theta / phi convolutions, the PatchGAN discriminators and the VGG19 feature net of the
reference (models/networks/generator.py:60-89,140-160; architecture.py:70-95; normalization.py:129-151;
correspondence.py:13-36,108-146,260-272; discriminator.py:138-177) as chains of hand-written sm_100a kernels: every
convolution (forward, backward-data, backward-weights) on the tap-convolution kernels, every normalisation /
modulation / activation / padding step on the NHWC elementwise kernels, activations fp16 (fp32 + 2-term split
operands on the path that feeds the correlation), gradients bf16.

The nn.Modules keep owning the parameters (state_dict layout, spectral norm, init and the optimisers are untouched);
a forward here gathers the module weights, runs ONE torch.autograd.Function whose inside is the tape, and hands fp32
NCHW tensors back to the torch side (losses, the correspondence kernel K1).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import nhwc
from .. import tape as T
from ..nhwc import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, F16, F32


# ------------------------------------------------------------------------------------------------ parameters
class ParamSet:
    """The tensors that become inputs of the boundary Function (conv weights after their spectral-norm / equal-lr
    pre-hooks ran -- exactly once per module per forward, like nn.Module.__call__ would) and their Params inside."""

    def __init__(self):
        self.tensors, self.slot, self.params = [], {}, None

    def _add(self, key, t):
        if key not in self.slot:
            self.slot[key] = len(self.tensors)
            self.tensors.append(t)

    def conv(self, m):
        if ("w", id(m)) in self.slot:
            return
        for hook in m._forward_pre_hooks.values():
            hook(m, (None,))
        self._add(("w", id(m)), m.weight)
        if m.bias is not None:
            self._add(("b", id(m)), m.bias)

    def spade(self, m):
        """gamma and beta convs share their input: one conv with concatenated filters (gb = [gamma | beta])."""
        self.conv(m.mlp_shared[1])
        self._add(("w", id(m)), torch.cat((m.mlp_gamma.weight, m.mlp_beta.weight), 0))
        self._add(("b", id(m)), torch.cat((m.mlp_gamma.bias, m.mlp_beta.bias), 0))

    def tensor(self, key, t):
        self._add(("t", key), t)

    def bind(self, params):
        self.params = params

    def w(self, m):
        return self.params[self.slot[("w", id(m))]]

    def b(self, m):
        i = self.slot.get(("b", id(m)))
        return None if i is None else self.params[i]

    def t(self, key):
        return self.params[self.slot[("t", key)]]


class Pyramid:
    """The SPADE condition (`segmap`) as fp16 NHWC operands at every resolution a network asks for: nearest
    down-sampling by an integer factor (F.interpolate(mode='nearest'), normalization.py:130) + the reflection halo of
    mlp_shared's ReflectionPad2d, packed straight from the fp32 NCHW input.  Gradients (only the channels that carry
    one: the warped exemplar) are summed back over all levels."""

    def __init__(self, tp, index, seg, grad_ch, split):
        self.tp, self.index, self.seg, self.grad_ch, self.split = tp, index, seg, grad_ch, split
        self.levels, self.sink = {}, {}

    def get(self, h, w, pad):
        key = (h, w, pad)
        if key not in self.levels:
            f = self.seg.shape[2] // h
            assert self.seg.shape[2] == h * f and self.seg.shape[3] == w * f, "SPADE condition: non-integer scale"
            self.levels[key] = T.pack_in(self.tp, self.index, self.seg, F16, pad=pad, split=self.split, f=f, size=(h, w),
                                         grad_ch=self.grad_ch, sink=self.sink)
        return self.levels[key]


# ------------------------------------------------------------------------------------------------ SPADE blocks
def enabled():
    import os
    return os.environ.get("COCOS_NHWC", "1") != "0"


def _dev_ok(t):
    """CUDA tensors on the native backend; anything on the emulation the CPU tests install."""
    return t.dtype == torch.float32 and (t.is_cuda or type(nhwc.backend()).__name__ != "NativeBackend")


def spade_supported(norm):
    return norm.pono and norm.mlp_gamma.kernel_size == (3, 3)


def block_supported(blk):
    norms = [blk.norm_0, blk.norm_1] + ([blk.norm_s] if blk.learned_shortcut else [])
    return blk.dilation == 1 and not blk.use_se and all(spade_supported(n) for n in norms)


def register_block(ps, blk):
    ps.conv(blk.conv_0)
    ps.conv(blk.conv_1)
    ps.spade(blk.norm_0)
    ps.spade(blk.norm_1)
    if blk.learned_shortcut:
        ps.conv(blk.conv_s)
        ps.spade(blk.norm_s)


def spade_norm(tp, ps, norm, x, pyr, mode, slope, pad):
    """SPADE.forward (normalization.py:129-151) + the activation / ReflectionPad2d that follow it in
    SPADEResnetBlock (architecture.py:73-74,94-95): raw x -> conv operand."""
    C = x.v.C
    seg = pyr.get(x.v.H, x.v.W, 1)
    shared = norm.mlp_shared[1]
    actv = T.conv(tp, seg, ps.w(shared), ps.b(shared), act=ACT_RELU, out_kind=F16, out_pad=1, split_out=mode.split,
                  dx_ch=pyr.grad_ch, wsplit=mode.split)
    gb = T.conv(tp, actv, ps.w(norm), ps.b(norm), out_kind=mode.raw)
    return T.spade(tp, x, gb, C, pad, slope, mode.split)


def spade_block(tp, ps, blk, x, pyr, mode, final_act=ACT_NONE, final_slope=0.0):
    """SPADEResnetBlock.forward (architecture.py:70-95); the residual sum (and, for the generator's last block, the
    LeakyReLU in front of conv_img) ride in conv_1's epilogue."""
    if blk.learned_shortcut:
        xs = T.conv(tp, spade_norm(tp, ps, blk.norm_s, x, pyr, mode, 1.0, 0), ps.w(blk.conv_s), None, out_kind=mode.raw)
    else:
        xs = x
    h = T.conv(tp, spade_norm(tp, ps, blk.norm_0, x, pyr, mode, 0.2, 1), ps.w(blk.conv_0), ps.b(blk.conv_0),
               out_kind=mode.raw)
    return T.conv(tp, spade_norm(tp, ps, blk.norm_1, h, pyr, mode, 0.2, 1), ps.w(blk.conv_1), ps.b(blk.conv_1),
                  out_kind=mode.raw, res=xs, act=final_act, slope=final_slope)


# ------------------------------------------------------------------------------------------------ SPADE generator
def generator_supported(net, seg):
    blocks = [net.head_0, net.G_middle_0, net.G_middle_1, net.up_0, net.up_1, net.up_2, net.up_3]
    return enabled() and _dev_ok(seg) and all(block_supported(b) for b in blocks) and seg.shape[2] == seg.shape[3] \
        and seg.shape[2] % 32 == 0 and net.sh == net.sw and isinstance(net.fc, nn.Conv2d)


def generator_forward(net, seg):
    """SPADEGenerator.forward (generator.py:60-89).  seg = cat(warp_out, semantics) fp32 NCHW."""
    opt = net.opt
    grad_ch = (0, 3) if ("warp" in opt.CBN_intype and seg.requires_grad) else None
    first = [net.head_0, net.G_middle_0, net.G_middle_1, net.up_0, net.up_1]
    last = [net.up_2, net.up_3]

    def stage_a(tp, ps, pyr):
        x = T.conv(tp, pyr.get(net.sh, net.sw, 0), ps.w(net.fc), ps.b(net.fc), padding=1, out_kind=F16, dx_ch=grad_ch)
        x = spade_block(tp, ps, net.head_0, x, pyr, T.FAST)
        x = spade_block(tp, ps, net.G_middle_0, T.upsample2(tp, x), pyr, T.FAST)
        x = spade_block(tp, ps, net.G_middle_1, x, pyr, T.FAST)
        x = spade_block(tp, ps, net.up_0, T.upsample2(tp, x), pyr, T.FAST)
        x = spade_block(tp, ps, net.up_1, T.upsample2(tp, x), pyr, T.FAST)
        return T.upsample2(tp, x)

    def stage_b(tp, ps, pyr, x):
        x = spade_block(tp, ps, net.up_2, x, pyr, T.FAST)
        # up_3 + leaky_relu(0.2) (generator.py:87) in its conv_1 epilogue, then conv_img + tanh -> fp32 NCHW
        x = spade_block(tp, ps, net.up_3, T.upsample2(tp, x), pyr, T.FAST, final_act=ACT_LRELU, final_slope=0.2)
        return T.conv(tp, x, ps.w(net.conv_img), ps.b(net.conv_img), padding=1, act=ACT_TANH, nchw=True)

    def run_stage(blocks, extra_convs, fn, inputs):
        ps = ParamSet()
        for m in extra_convs:
            ps.conv(m)
        for blk in blocks:
            register_block(ps, blk)

        def body(tp, ins, params):
            ps.bind(params)
            pyr = Pyramid(tp, 0, ins[0], grad_ch, False)
            return fn(tp, ps, pyr, ins)
        return T.run(body, inputs, ps.tensors)

    if not opt.use_attention:
        return run_stage(first + last, [net.fc, net.conv_img],
                         lambda tp, ps, pyr, ins: [stage_b(tp, ps, pyr, stage_a(tp, ps, pyr))], [seg])[0]
    mid = run_stage(first, [net.fc], lambda tp, ps, pyr, ins: [T.unpack_out(tp, stage_a(tp, ps, pyr))], [seg])[0]
    mid = net.attn(mid)
    return run_stage(last, [net.conv_img],
                     lambda tp, ps, pyr, ins: [stage_b(tp, ps, pyr, T.pack_in(tp, 1, ins[1], F16, grad_ch=(0, ins[1].shape[1])))],
                     [seg, mid])[0]


# ------------------------------------------------------------------------------------------------ domain adaptor
def _plain_in_layer(layer):
    return isinstance(layer, nn.Sequential) and len(layer) == 2 and isinstance(layer[0], nn.Conv2d) \
        and isinstance(layer[1], nn.InstanceNorm2d) and not layer[1].affine and not layer[1].track_running_stats \
        and layer[0].padding_mode == "zeros" and layer[0].dilation == (1, 1) and layer[0].groups == 1


def adaptor_supported(net, x):
    opt = net.opt
    layers = [net.layer1, net.layer2, net.layer3, net.layer4, net.layer5]
    blocks = [net.head_0, net.G_middle_0, net.G_middle_1]
    return enabled() and _dev_ok(x) and not x.requires_grad \
        and all(_plain_in_layer(l) for l in layers) and all(block_supported(b) for b in blocks) \
        and not opt.adaptor_nonlocal and not opt.adaptor_res_deeper and x.shape[2] % 8 == 0 and x.shape[3] % 8 == 0


def adaptor_forward(net, x, precise=True):
    """AdaptiveFeatureGenerator.forward (generator.py:140-160) with seg == input (how NoVGGCorrespondence calls it,
    correspondence.py:245-253).  x fp32 NCHW without gradient (label map or image) -> fp32 NCHW features."""
    mode = T.PRECISE if precise else T.FAST
    layers = [net.layer1, net.layer2, net.layer3, net.layer4, net.layer5]
    blocks = [net.head_0, net.G_middle_0, net.G_middle_1]
    # a one-hot label map is exact in fp16: no lo term for the activations, only for the weights
    is_image = x.shape[1] <= 4
    xsplit = mode.split and is_image
    ps = ParamSet()
    for l in layers:
        ps.conv(l[0])
    for blk in blocks:
        register_block(ps, blk)

    def body(tp, ins, params):
        ps.bind(params)
        src = ins[0]
        h = T.pack_in(tp, 0, src, F16, split=xsplit)
        for i, l in enumerate(layers):
            c = l[0]
            r = T.conv(tp, h, ps.w(c), None, stride=c.stride[0], padding=c.padding[0], out_kind=mode.raw,
                       wsplit=mode.split)
            if i < 4:  # InstanceNorm + LeakyReLU(0.2) -> operand of the next (zero padded) conv
                h, _ = T.inst_act(tp, r, slope=0.2, eps=l[1].eps, out_kind=F16, split_out=mode.split)
            else:      # layer5: InstanceNorm only -> raw input of the SPADE blocks
                h, _ = T.inst_act(tp, r, slope=1.0, eps=l[1].eps, out_kind=mode.raw)
        pyr = Pyramid(tp, 0, src, None, xsplit)
        for blk in blocks:
            h = spade_block(tp, ps, blk, h, pyr, mode)
        return [T.unpack_out(tp, h)]
    return T.run(body, [x], ps.tensors)[0]


# ------------------------------------------------------------------------------------------------ residual stack
def resstack_supported(net, cont):
    return enabled() and _dev_ok(cont) \
        and all(isinstance(b.bn1, nn.InstanceNorm2d) and not b.bn1.affine and b.conv1.kernel_size == (3, 3)
                and b.conv1.stride == (1, 1) and b.prelu.weight.numel() == 1 for b in net.layer) \
        and net.theta.kernel_size == (1, 1)


def resstack_forward(net, cont, ref, precise=True):
    """`self.layer` (4 ResidualBlocks, correspondence.py:13-36,175-179) on both domains as ONE batch of 2B images
    (every op is per sample), then the theta / phi 1x1 convolutions (correspondence.py:272,282) -> fp32 NCHW."""
    mode = T.PRECISE if precise else T.FAST
    B = cont.shape[0]
    ps = ParamSet()
    for blk in net.layer:
        ps.conv(blk.conv1)
        ps.conv(blk.conv2)
        ps.tensor(id(blk.prelu), blk.prelu.weight)
    ps.conv(net.theta)
    ps.conv(net.phi)

    def body(tp, ins, params):
        ps.bind(params)
        both = ins[0]
        C = both.shape[1]
        sink = {}
        x_op = T.pack_in(tp, 0, both, F16, pad=1, split=mode.split, grad_ch=(0, C), sink=sink)
        x_raw = T.pack_in(tp, 0, both, F32, grad_ch=(0, C), sink=sink)
        for blk in net.layer:
            a = ps.t(id(blk.prelu))
            r1 = T.conv(tp, x_op, ps.w(blk.conv1), ps.b(blk.conv1), out_kind=mode.raw, wsplit=mode.split)
            h, _ = T.inst_act(tp, r1, prelu=a, eps=blk.bn1.eps, out_kind=F16, out_pad=1, split_out=mode.split)
            r2 = T.conv(tp, h, ps.w(blk.conv2), ps.b(blk.conv2), out_kind=mode.raw, wsplit=mode.split)
            x_op, x_raw = T.inst_act(tp, r2, prelu=a, res=x_raw, eps=blk.bn2.eps, out_kind=F16, out_pad=1,
                                     split_out=mode.split, want_raw=True)
        # 1x1 convs on the interior of the haloed operand: padding -1 crops the halo
        theta = T.conv(tp, T.slice_batch(tp, x_op, 0, B), ps.w(net.theta), ps.b(net.theta), padding=-1, nchw=True,
                       wsplit=mode.split)
        phi = T.conv(tp, T.slice_batch(tp, x_op, B, 2 * B), ps.w(net.phi), ps.b(net.phi), padding=-1, nchw=True,
                     wsplit=mode.split)
        return [theta, phi]

    return T.run(body, [torch.cat((cont, ref), 0)], ps.tensors)
