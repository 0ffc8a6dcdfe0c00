"""Mint model-level goldens from the UNMODIFIED reference on CPU (build
container only): one full generator step + discriminator step of
Pix2PixModel at B=1 with seeded init / seeded inputs; stores the loss values,
subsampled outputs and a few gradient norms.  python tests/golden/make_golden_model.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_harness as rh  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[1] flags at batch 1 (match_kernel default 3)
    "ade20k_train": ["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C",
                     "--warp_mask_losstype", "direct", "--weight_mask", "100.0", "--vgg_normal_correct",
                     "--batchSize", "1"],
    # BASELINE.json configs[2] flags (no --PONO: SPADE's param-free norm is the (Sync)BatchNorm stub) + the cycle term
    "celebahq_train": ["--dataset_mode", "celebahq", "--warp_bilinear", "--adaptor_kernel", "4",
                       "--warp_cycle_w", "1.0", "--batchSize", "1"],
    # BASELINE.json configs[3] flags: 4x4 patch warp + fold, float pose maps
    "deepfashion_train": ["--dataset_mode", "deepfashion", "--warp_patch", "--video_like", "--batchSize", "1"],
}


def synthetic_batch_for(opt, batch, seed=1234):
    """Same generator as cocosnet_b200.data.synthetic_batch (kept in sync by test)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from cocosnet_b200 import data as cdata
    return cdata.synthetic_batch(opt, batch, seed=seed)


def run(name):
    argv = CONFIGS[name]
    batch = synthetic_batch_for(_opt_stub(argv), 1)
    with rh.reference_imported(), rh.patched_for_cpu_training():
        opt = rh.make_opt(argv, True)
        from models.pix2pix_model import Pix2PixModel
        torch.manual_seed(0)
        model = Pix2PixModel(opt)
        model.train()
        def fresh():
            d = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
            if opt.dataset_mode not in ("deepfashion",):
                d["label"] = d["label"].long()
                d["label_ref"] = d["label_ref"].long()  # reference CPU path never casts it (pix2pix_model.py:171-173)
            return d
        real_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self  # deepfashion / celebahqedge call .cuda() unconditionally
        data = fresh()
        g_losses, out = model(data, mode="generator")
        g_loss = sum(g_losses.values()).mean()
        g_loss.backward()
        res = {"g_" + k: v.detach().numpy().astype(np.float64).reshape(-1) for k, v in g_losses.items()}
        res["fake_image_sub"] = out["fake_image"].detach().numpy()[:, :, ::4, ::4]
        res["warp_out_sub"] = out["warp_out"].detach().numpy()[:, :, ::4, ::4]
        if out.get("warp_mask") is not None:
            res["warp_mask_chsum"] = out["warp_mask"].detach().numpy().sum(1)
        if out.get("warp_cycle") is not None:
            res["warp_cycle"] = out["warp_cycle"].detach().numpy()
        gn = {}
        for key in ("netG", "netCorr"):
            for pname, p in model.net[key].named_parameters():
                if p.grad is not None and (pname.endswith("conv_img.weight") or pname.endswith("theta.weight")
                                           or pname.endswith("phi.weight") or "layer1.0.weight_orig" in pname
                                           or pname.endswith("fc.weight") or pname.endswith("attn.gamma")):
                    gn["gradnorm_%s_%s" % (key, pname)] = np.array([float(p.grad.norm())])
        res.update(gn)
        d_losses = model(fresh(), mode="discriminator", GforD={"fake_image": out["fake_image"]})
        torch.Tensor.cuda = real_cuda
        res.update({"d_" + k: v.detach().numpy().astype(np.float64).reshape(-1) for k, v in d_losses.items()})
    path = os.path.join(HERE, "model_%s.npz" % name)
    np.savez_compressed(path, **res)
    print(name, {k: (v.shape if v.size > 1 else float(v[0])) for k, v in res.items()})
    print(os.path.getsize(path) // 1024, "KiB")


INFER_CONFIGS = {
    # BASELINE.json configs[0]: the reference's CPU-runnable numerics case -- ade20k inference flags at batch 1, with the
    # default match_kernel 3 (K = 2304) and with match_kernel 1 (K = 256, the shape the kernel metric is quoted on)
    "ade20k_infer_mk3": ["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C",
                         "--batchSize", "1"],
    "ade20k_infer_mk1": ["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C",
                         "--match_kernel", "1", "--batchSize", "1"],
}


def run_inference(name):
    """Seeded init (no checkpoint exists offline), modules in eval(), `mode='inference'` (pix2pix_model.py:74-86,
    319-334): what test.py computes per batch."""
    argv = INFER_CONFIGS[name]
    batch = synthetic_batch_for(_opt_stub(argv), 1)
    with rh.reference_imported(), rh.patched_for_cpu_training():
        opt = rh.make_opt(argv, True)
        from models.pix2pix_model import Pix2PixModel
        torch.manual_seed(0)
        model = Pix2PixModel(opt)
        model.eval()
        opt.isTrain = False  # constructed as a training model for the seeded init; forward as test.py would
        opt.show_corr = False  # TestOptions default (test_options.py), read by correspondence.py:325
        d = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
        d["label"] = d["label"].long()
        d["label_ref"] = d["label_ref"].long()
        with torch.no_grad():
            out = model(d, mode="inference")
        res = {"fake_image_sub": out["fake_image"].numpy()[:, :, ::4, ::4],
               "warp_out_sub": out["warp_out"].numpy()[:, :, ::4, ::4]}
    path = os.path.join(HERE, "model_%s.npz" % name)
    np.savez_compressed(path, **res)
    print(name, {k: v.shape for k, v in res.items()}, os.path.getsize(path) // 1024, "KiB")


def _opt_stub(argv):
    from cocosnet_b200.options import TrainOptions
    return TrainOptions().parse(list(argv) + ["--gpu_ids", "-1"], save=False, verbose=False)


if __name__ == "__main__":
    for n in sys.argv[1:] or (list(CONFIGS) + list(INFER_CONFIGS)):
        run_inference(n) if n in INFER_CONFIGS else run(n)
