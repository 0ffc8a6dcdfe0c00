"""CPU: the 16-bit NHWC tape (cocosnet_b200/tape.py, nets/fast.py) -- generator, domain adaptors, residual stack,
discriminators, VGG -- driven through the torch emulation of the kernels (oracle/nhwc_emul.py, exact mode) must
reproduce the goldens minted from the unmodified reference: same losses, outputs and gradient norms as the plain
host mirror.  This pins the host side of the fast path (tape bookkeeping, hand-written backward chain, weight
gathering incl. spectral norm) to the reference without a GPU; the kernels are pinned to the emulation on the GPU."""
import os

import numpy as np
import pytest
import torch

from cocosnet_b200 import data as cdata
from cocosnet_b200 import nhwc
from cocosnet_b200.options import TrainOptions
from oracle import torch_port
from oracle.nhwc_emul import EmulBackend

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ADE_TRAIN = ["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C",
             "--warp_mask_losstype", "direct", "--weight_mask", "100.0", "--vgg_normal_correct", "--batchSize", "1",
             "--gpu_ids", "-1"]


class CountingEmul(EmulBackend):
    def __init__(self):
        super().__init__(exact=True)
        self.calls = {"tapconv": 0, "tapwgrad": 0, "spade_fwd": 0, "inst_fwd": 0, "maxpool_fwd": 0, "maxpool_bwd": 0,
                      "stride2": 0, "spade_epilogue": 0, "inst_mod": 0}

    def maxpool_fwd(self, *a, **k):
        self.calls["maxpool_fwd"] += 1
        return super().maxpool_fwd(*a, **k)

    def maxpool_bwd(self, *a, **k):
        self.calls["maxpool_bwd"] += 1
        return super().maxpool_bwd(*a, **k)

    def tapconv(self, *a, **k):
        self.calls["tapconv"] += 1
        self.calls["stride2"] += int(a[5]["a_stride"] == 2 and len(a[5]["groups"]) == 16)  # PatchGAN 4x4 / stride 2
        self.calls["spade_epilogue"] += int(a[5].get("mod") is not None)  # SPADE modulation in the conv epilogue
        return super().tapconv(*a, **k)

    def tapwgrad(self, *a, **k):
        self.calls["tapwgrad"] += 1
        return super().tapwgrad(*a, **k)

    def spade_fwd(self, *a, **k):
        self.calls["spade_fwd"] += 1
        return super().spade_fwd(*a, **k)

    def inst_fwd(self, *a, **k):
        self.calls["inst_fwd"] += 1
        self.calls["inst_mod"] += int(k.get("gb") is not None)
        return super().inst_fwd(*a, **k)


CONFIGS = {
    "ade20k_train": ADE_TRAIN,
    # no --PONO: SPADE with (Sync)BatchNorm statistics, adaptor_kernel 4, bilinear warp, cycle term
    "celebahq_train": ["--dataset_mode", "celebahq", "--warp_bilinear", "--adaptor_kernel", "4", "--warp_cycle_w", "1.0",
                       "--batchSize", "1", "--gpu_ids", "-1"],
    # float pose maps as the label input (split operands from the first layer on), folded patch warp
    "deepfashion_train": ["--dataset_mode", "deepfashion", "--warp_patch", "--video_like", "--batchSize", "1",
                          "--gpu_ids", "-1"],
}


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("config", list(CONFIGS))
def test_train_step_on_the_tape_matches_reference_golden(config):
    from cocosnet_b200.pix2pix_model import Pix2PixModel
    be = CountingEmul()
    old = nhwc.set_backend(be)
    try:
        gold = np.load(os.path.join(GOLD, "model_%s.npz" % config))
        opt = TrainOptions().parse(CONFIGS[config], save=False, verbose=False)
        opt.verbose_networks = False
        opt.allow_random_vgg = True
        torch.manual_seed(0)
        model = Pix2PixModel(opt)
        model.vggnet_fix.load_state_dict(torch_port.seeded_vgg_state_dict())
        model.train()
        batch = cdata.synthetic_batch(opt, 1)
        with torch_port.cpu_reference_mode():
            g_losses, out = model(batch, mode="generator")
            sum(g_losses.values()).mean().backward()
            d_losses = model(batch, mode="discriminator", GforD={"fake_image": out["fake_image"]})
            sum(d_losses.values()).mean().backward()
    finally:
        nhwc.set_backend(old)
    # the networks really ran on the tape: 3 adaptor passes x (5 + 24) convs, 7 generator blocks, the residual stack ...
    assert be.calls["tapconv"] > 200 and be.calls["tapwgrad"] > 100 and be.calls["spade_fwd"] == 0, be.calls
    if config == "ade20k_train":
        # --PONO: every SPADE layer modulates in the epilogue of its gamma|beta convolution (no separate pass)
        assert be.calls["spade_epilogue"] > 30, be.calls
    else:
        # batch statistics: gamma|beta convolution + ONE modulating norm kernel per SPADE layer
        assert be.calls["spade_epilogue"] == 0 and be.calls["inst_mod"] > 30, be.calls
    # ... the VGG19 feature net (3 forward passes x 4 poolings, one backward) and both PatchGANs (3 stride-2 4x4
    # convolutions each, G step: fake + real halves, D step: one batch)
    assert be.calls["maxpool_fwd"] == 12 and be.calls["maxpool_bwd"] == 4 and be.calls["stride2"] >= 18, be.calls
    for k, v in g_losses.items():
        assert np.allclose(v.detach().numpy().reshape(-1), gold["g_" + k], rtol=2e-4, atol=1e-6), (k, v, gold["g_" + k])
    for k, v in d_losses.items():
        assert np.allclose(v.detach().numpy().reshape(-1), gold["d_" + k], rtol=2e-4), k
    assert np.allclose(out["fake_image"].detach().numpy()[:, :, ::4, ::4], gold["fake_image_sub"], atol=2e-5)
    assert np.allclose(out["warp_out"].detach().numpy()[:, :, ::4, ::4], gold["warp_out_sub"], atol=2e-5)
    for key in gold.files:
        if key.startswith("gradnorm_"):
            _, netk, pname = key.split("_", 2)
            p = dict(model.net[netk].named_parameters())[pname]
            assert abs(float(p.grad.norm()) - float(gold[key][0])) <= 2e-3 * float(gold[key][0]), key


def _small_train_step(argv, fast, crop=64):
    """One generator + discriminator pass at crop x crop through the tape on the emulation (fast) or through the plain
    host mirror (COCOS_NHWC=0); returns (losses + outputs, gradients)."""
    from cocosnet_b200.pix2pix_model import Pix2PixModel
    saved = os.environ.get("COCOS_NHWC")
    os.environ["COCOS_NHWC"] = "1" if fast else "0"
    old = nhwc.set_backend(EmulBackend(exact=True))
    try:
        opt = TrainOptions().parse(argv + ["--batchSize", "1", "--gpu_ids", "-1", "--crop_size", str(crop), "--load_size",
                                           str(crop)], save=False, verbose=False)
        opt.verbose_networks = False
        opt.allow_random_vgg = True
        torch.manual_seed(0)
        model = Pix2PixModel(opt)
        model.vggnet_fix.load_state_dict(torch_port.seeded_vgg_state_dict())
        model.train()
        batch = cdata.synthetic_batch(opt, 1)
        with torch_port.cpu_reference_mode():
            g_losses, out = model(batch, mode="generator")
            sum(g_losses.values()).mean().backward()
            grads = {k + "/" + n: p.grad.clone() for k in ("netG", "netCorr")
                     for n, p in model.net[k].named_parameters() if p.grad is not None}
            d_losses = model(batch, mode="discriminator", GforD={"fake_image": out["fake_image"]})
            sum(d_losses.values()).mean().backward()
            grads.update({"netD/" + n: p.grad.clone() for n, p in model.net["netD"].named_parameters()
                          if p.grad is not None})
    finally:
        nhwc.set_backend(old)
        if saved is None:
            os.environ.pop("COCOS_NHWC")
        else:
            os.environ["COCOS_NHWC"] = saved
    vals = {"g_" + k: v.detach().reshape(-1) for k, v in g_losses.items()}
    vals.update({"d_" + k: v.detach().reshape(-1) for k, v in d_losses.items()})
    vals.update(fake_image=out["fake_image"].detach(), warp_out=out["warp_out"].detach())
    return vals, grads


SMALL = {
    # 256 + 151 + 2 = 409 channels in the residual stack: C % 8 == 1 (unwritten channel slots, see
    # test_norm_act_outputs_leave_no_uninitialised_channel_slots)
    "coordconv_409": ["--dataset_mode", "ade20k", "--PONO", "--PONO_C", "--maskmix", "--use_coordconv"],
    # 256 + 19 = 275 channels, batch-statistics SPADE, bilinear warp, cycle term, attention
    "celebahq_maskmix_275": ["--dataset_mode", "celebahq", "--maskmix", "--use_attention", "--warp_bilinear",
                             "--warp_cycle_w", "0.1"],
    # column-softmax mask (correspondence.py:337-346)
    "cycle_mask": ["--dataset_mode", "ade20k", "--PONO", "--PONO_C", "--maskmix", "--warp_mask_losstype", "cycle"],
    # 4x4 adaptor kernels + edge maps as labels + the two-cycle term
    "celebahqedge_two_cycle": ["--dataset_mode", "celebahqedge", "--PONO", "--PONO_C", "--adaptor_kernel", "4",
                               "--warp_cycle_w", "1.0", "--two_cycle"],
}


@pytest.mark.timeout(600)
@pytest.mark.parametrize("name", list(SMALL))
def test_tape_equals_plain_mirror_on_other_flag_sets(name):
    """Flag sets without a reference golden, at 64x64 so that they are cheap: the tape (every network on the kernel
    emulation) against the plain host mirror, which test_model_parity_cpu.py pins to the reference.  Losses and outputs
    to fp32 rounding; gradients to 1e-2 (1/T = 100 amplifies rounding into the correspondence gradients), skipping
    parameters whose gradient is analytically zero (biases in front of a normalisation)."""
    want, gw = _small_train_step(SMALL[name], fast=False)
    got, gg = _small_train_step(SMALL[name], fast=True)
    assert set(want) == set(got) and set(gw) == set(gg)
    for k in want:
        d = float((want[k].double() - got[k].double()).norm() / (want[k].double().norm() + 1e-30))
        assert d < 1e-4, (k, d)
    med = {}
    for k, v in gw.items():
        med.setdefault(k.split("/")[0], []).append(float(v.norm()))
    med = {k: float(np.median(v)) for k, v in med.items()}
    checked = 0
    for k, v in gw.items():
        if float(v.norm()) < 1e-2 * med[k.split("/")[0]]:
            continue
        d = float((v.double() - gg[k].double()).norm() / v.double().norm())
        assert d < 1e-2, (k, d)
        checked += 1
    assert checked > 100
