"""CPU: the host-side mirror (options, modules, losses, trainer logic) against
goldens minted from the unmodified reference.  The fused CUDA primitive is
swapped for the reference's unfused torch expressions (oracle/torch_port.py),
so this pins the host logic and the CPU port that bench.py times."""
import json
import os

import numpy as np
import pytest
import torch

from cocosnet_b200 import data as cdata
from cocosnet_b200 import nets
from cocosnet_b200.options import TestOptions as _TestOptions, TrainOptions
from oracle import torch_port

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ADE_TRAIN = ["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C",
             "--warp_mask_losstype", "direct", "--weight_mask", "100.0", "--vgg_normal_correct", "--batchSize", "1",
             "--gpu_ids", "-1"]


def test_option_defaults_match_reference_snapshot():
    snap = json.load(open(os.path.join(GOLD, "options_snapshot.json")))
    for name, (argv, is_train) in snap["cases"].items():
        cls = TrainOptions if is_train else _TestOptions
        opt = cls().parse(argv + ["--gpu_ids", "-1"], save=False, verbose=False)
        mine = {k: v for k, v in vars(opt).items() if k not in ("gpu_ids", "corr_precision", "conv_precision", "channels_last")}
        ref = {k: v for k, v in snap["values"][name].items() if k not in ("gpu_ids", "down")}
        for k, v in ref.items():
            mv = mine[k]
            if isinstance(v, float) and v in (float("inf"),):
                assert mv == v
            else:
                assert mv == v or (isinstance(v, float) and abs(mv - v) < 1e-12), (name, k, mv, v)
        assert set(mine) == set(ref), (name, set(mine) ^ set(ref))


def test_state_dict_keys_match_reference_snapshot():
    snap = json.load(open(os.path.join(GOLD, "state_dict_keys.json")))
    opt = TrainOptions().parse(ADE_TRAIN, save=False, verbose=False)
    opt.verbose_networks = False
    built = {"G": nets.define_G(opt), "D": nets.define_D(opt), "Corr": nets.define_Corr(opt)}
    for name, net in built.items():
        mine = [[k, list(v.shape)] for k, v in net.state_dict().items()]
        assert mine == snap[name], name


MODEL_CONFIGS = {
    "ade20k_train": ADE_TRAIN,
    "celebahq_train": ["--dataset_mode", "celebahq", "--warp_bilinear", "--adaptor_kernel", "4", "--warp_cycle_w", "1.0",
                       "--batchSize", "1", "--gpu_ids", "-1"],
    "deepfashion_train": ["--dataset_mode", "deepfashion", "--warp_patch", "--video_like", "--batchSize", "1",
                          "--gpu_ids", "-1"],
}


@pytest.mark.timeout(900)
@pytest.mark.parametrize("config", list(MODEL_CONFIGS))
def test_full_train_step_matches_reference_golden(config):
    from cocosnet_b200.pix2pix_model import Pix2PixModel
    gold = np.load(os.path.join(GOLD, "model_%s.npz" % config))
    opt = TrainOptions().parse(MODEL_CONFIGS[config], save=False, verbose=False)
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
    for k, v in g_losses.items():
        want = gold["g_" + k]
        assert np.allclose(v.detach().numpy().reshape(-1), want, rtol=2e-4, atol=1e-6), (k, v, want)
    for k, v in d_losses.items():
        assert np.allclose(v.detach().numpy().reshape(-1), gold["d_" + k], rtol=2e-4), k
    assert np.allclose(out["fake_image"].detach().numpy()[:, :, ::4, ::4], gold["fake_image_sub"], atol=2e-5)
    assert np.allclose(out["warp_out"].detach().numpy()[:, :, ::4, ::4], gold["warp_out_sub"], atol=2e-5)
    if "warp_mask_chsum" in gold.files:
        assert np.allclose(out["warp_mask"].detach().numpy().sum(1), gold["warp_mask_chsum"], atol=1e-4)
    if "warp_cycle" in gold.files:
        assert np.allclose(out["warp_cycle"].detach().numpy(), gold["warp_cycle"], atol=2e-5)
    for key in gold.files:
        if key.startswith("gradnorm_"):
            _, netk, pname = key.split("_", 2)
            p = dict(model.net[netk].named_parameters())[pname]
            assert abs(float(p.grad.norm()) - float(gold[key][0])) <= 2e-3 * float(gold[key][0]), key


INFER_CONFIGS = {
    "ade20k_infer_mk3": ["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C",
                         "--batchSize", "1", "--gpu_ids", "-1"],
    "ade20k_infer_mk1": ["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C",
                         "--match_kernel", "1", "--batchSize", "1", "--gpu_ids", "-1"],
}


def build_inference_model(config, gpu=False):
    """Seeded training-mode construction (bit-identical init, no checkpoint offline), then eval() + isTrain=False:
    what the golden script does to the reference (tests/golden/make_golden_model.py:run_inference)."""
    from cocosnet_b200.pix2pix_model import Pix2PixModel
    opt = TrainOptions().parse(INFER_CONFIGS[config], save=False, verbose=False)
    opt.verbose_networks = False
    opt.allow_random_vgg = True
    torch.manual_seed(0)
    model = Pix2PixModel(opt)
    model.eval()
    opt.isTrain = False
    opt.show_corr = False
    if gpu:
        opt.gpu_ids = [0]
        model.cuda()
    return opt, model


@pytest.mark.timeout(900)
@pytest.mark.parametrize("config", list(INFER_CONFIGS))
def test_inference_matches_reference_golden(config):
    """BASELINE configs[0] (the reference's CPU-runnable case): `mode='inference'` of the host mirror on CPU."""
    gold = np.load(os.path.join(GOLD, "model_%s.npz" % config))
    opt, model = build_inference_model(config)
    batch = cdata.synthetic_batch(opt, 1)
    with torch_port.cpu_reference_mode(), torch.no_grad():
        out = model(batch, mode="inference")
    assert np.allclose(out["warp_out"].numpy()[:, :, ::4, ::4], gold["warp_out_sub"], atol=2e-5)
    assert np.allclose(out["fake_image"].numpy()[:, :, ::4, ::4], gold["fake_image_sub"], atol=2e-5)
