"""GPU: the full model on the fused sm_100a kernels against (a) the goldens
minted from the unmodified reference and (b) the CPU port on the same weights."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ADE_TRAIN = ["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C",
             "--warp_mask_losstype", "direct", "--weight_mask", "100.0", "--vgg_normal_correct", "--batchSize", "1"]


MODEL_CONFIGS = {
    "ade20k_train": ADE_TRAIN,
    "celebahq_train": ["--dataset_mode", "celebahq", "--warp_bilinear", "--adaptor_kernel", "4", "--warp_cycle_w", "1.0",
                       "--batchSize", "1"],
    "deepfashion_train": ["--dataset_mode", "deepfashion", "--warp_patch", "--video_like", "--batchSize", "1"],
}


def _build(gpu, config="ade20k_train"):
    from cocosnet_b200.options import TrainOptions
    from cocosnet_b200.pix2pix_model import Pix2PixModel
    from oracle import torch_port
    opt = TrainOptions().parse(MODEL_CONFIGS[config] + ["--gpu_ids", "-1"], save=False, verbose=False)
    opt.verbose_networks = False
    opt.allow_random_vgg = True
    torch.manual_seed(0)
    model = Pix2PixModel(opt)  # seeded CPU init == the reference's (tests/test_model_parity_cpu.py)
    model.vggnet_fix.load_state_dict(torch_port.seeded_vgg_state_dict())
    model.train()
    if gpu:
        opt.gpu_ids = [0]
        model.cuda()
    return opt, model


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("native_conv", [False, True])
@pytest.mark.parametrize("config", list(MODEL_CONFIGS))
def test_train_step_matches_reference_golden(config, native_conv):
    """native_conv=False: every convolution in fp32 (cuDNN, TF32 off) -> the 1e-3 bar on the outputs.
    native_conv=True: conv forwards on the tcgen05 kernel with fp16 operands (TF32-class rounding, what stock
    PyTorch does by default on this GPU) -> 5e-3 on the generator output, same bar on the correspondence."""
    from cocosnet_b200 import data as cdata, ops
    gold = np.load(os.path.join(GOLD, "model_%s.npz" % config))
    old = torch.backends.cudnn.allow_tf32
    old_native, old_dgrad = ops.NATIVE_CONV, ops.NATIVE_DGRAD
    torch.backends.cudnn.allow_tf32 = False  # strict numerics for the parity check
    ops.NATIVE_CONV = ops.NATIVE_DGRAD = native_conv  # native: forward, backward-data and backward-weights on K2
    tol = 5.0 if native_conv else 1.0
    try:
        opt, model = _build(gpu=True, config=config)
        batch = cdata.synthetic_batch(opt, 1)
        g_losses, out = model(batch, mode="generator")
        sum(g_losses.values()).mean().backward()
        d_losses = model(batch, mode="discriminator", GforD={"fake_image": out["fake_image"]})
    finally:
        torch.backends.cudnn.allow_tf32 = old
        ops.NATIVE_CONV, ops.NATIVE_DGRAD = old_native, old_dgrad
    # outputs: north-star tolerance 1e-3 relative
    # (TF32-class conv rounding upstream of the correlation is amplified by 1/temperature = 100: measured
    #  3e-3..6e-3 on warp_out for cuDNN-TF32 and for the native kernels alike, profiles/r01_precision_modes.txt)
    assert _rel(out["warp_out"].detach().cpu().numpy()[:, :, ::4, ::4], gold["warp_out_sub"]) < (1e-2 if native_conv else 1e-3)
    assert _rel(out["fake_image"].detach().cpu().numpy()[:, :, ::4, ::4], gold["fake_image_sub"]) < 1e-3 * tol
    if "warp_mask_chsum" in gold.files:
        assert np.abs(out["warp_mask"].detach().cpu().numpy().sum(1) - gold["warp_mask_chsum"]).max() < 2e-3
    if "warp_cycle" in gold.files:
        assert _rel(out["warp_cycle"].detach().cpu().numpy(), gold["warp_cycle"]) < 2e-3 * tol
    for k, v in g_losses.items():
        want = float(gold["g_" + k][0])
        assert abs(float(v.mean()) - want) <= 2e-3 * tol * max(abs(want), 1.0), (k, float(v.mean()), want)
    for k, v in d_losses.items():
        want = float(gold["d_" + k][0])
        assert abs(float(v.mean()) - want) <= 2e-3 * tol * abs(want), k
    # gradients through the fused backward (fp16 dS): looser
    for key in gold.files:
        if key.startswith("gradnorm_"):
            _, netk, pname = key.split("_", 2)
            p = dict(model.net[netk].named_parameters())[pname]
            assert abs(float(p.grad.norm()) - float(gold[key][0])) <= 2e-2 * float(gold[key][0]), \
                (key, float(p.grad.norm()), float(gold[key][0]))


def test_inference_mode_runs_and_is_deterministic():
    from cocosnet_b200 import data as cdata
    from cocosnet_b200.options import TestOptions as TOpt
    from cocosnet_b200.pix2pix_model import Pix2PixModel
    opt = TOpt().parse(["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C",
                        "--gpu_ids", "-1", "--name", "nonexistent_ckpt", "--batchSize", "2"], save=False, verbose=False)
    opt.verbose_networks = False
    torch.manual_seed(0)
    model = Pix2PixModel(opt)
    opt.gpu_ids = [0]
    model.cuda().eval()
    batch = cdata.synthetic_batch(opt, 2)
    o1 = model(batch, mode="inference")
    o2 = model(batch, mode="inference")
    assert o1["fake_image"].shape == (2, 3, 256, 256) and o1["warp_out"].shape == (2, 3, 256, 256)
    assert torch.equal(o1["fake_image"], o2["fake_image"])
    assert torch.isfinite(o1["fake_image"]).all()


@pytest.mark.timeout(900)
def test_graphed_train_step_tracks_eager(monkeypatch):
    """trainer.run_step: eager for GRAPH_WARMUP calls, then ONE CUDA graph per iteration.  Same seed, same batches:
    the losses after 6 iterations agree with a trainer that never captures (split-K atomics and cuDNN algorithm
    choices make the two runs differ in the last bits, hence a tolerance, not equality)."""
    from cocosnet_b200 import data as cdata
    from cocosnet_b200.options import TrainOptions
    from cocosnet_b200.trainer import Pix2PixTrainer
    from oracle import torch_port

    def run(use_graph):
        opt = TrainOptions().parse(ADE_TRAIN[:-1] + ["2", "--gpu_ids", "0"], save=False, verbose=False)
        opt.verbose_networks = False
        opt.allow_random_vgg = True
        torch.manual_seed(0)
        trainer = Pix2PixTrainer(opt)
        trainer.pix2pix_model.vggnet_fix.load_state_dict(torch_port.seeded_vgg_state_dict())
        if not use_graph:
            trainer.graph_error = "disabled for the comparison"
        hist = []
        for it in range(6):
            batch = cdata.synthetic_batch(opt, 2, seed=100 + it)
            trainer.run_step(batch)
            hist.append({k: float(v.mean()) for k, v in trainer.get_latest_losses().items()})
        return trainer, hist

    if os.environ.get("COCOS_CUDA_GRAPH", "1") != "1":
        pytest.skip("CUDA-graph step disabled by COCOS_CUDA_GRAPH=0")
    tg, hg = run(True)
    assert tg._graph is not None, tg.graph_error
    assert tg.graph_native_launches > 100
    te, he = run(False)
    assert te._graph is None
    print("graphed:", hg)
    print("eager:  ", he)
    # iteration 3 is the first replay.  Everything but the mask loss agrees to 3e-2 there (measured: <= 6e-3; the
    # mask term, a log of tiny probabilities weighted by 100, already differs by 2e-4 between two EAGER runs at
    # iteration 1 because of split-K / atomics ordering, and by 15 % at iteration 3); two replays later the slow
    # reconstruction losses still agree while the adversarial ones have diverged chaotically (batch 2, random data).
    for k in he[3]:
        tol = 0.5 if k == "mask" else 3e-2
        assert abs(hg[3][k] - he[3][k]) <= tol * max(abs(he[3][k]), 1.0), (3, k, hg[3][k], he[3][k])
    for k in ("perc", "contextual", "fm"):
        assert abs(hg[5][k] - he[5][k]) <= 3e-2 * abs(he[5][k]), (5, k, hg[5][k], he[5][k])
    # the graph really trains: losses move between replays and the weights differ from the start
    assert any(abs(hg[-1][k] - hg[-2][k]) > 0 for k in hg[-1])


@pytest.mark.timeout(900)
@pytest.mark.parametrize("native_conv", [False, True])
@pytest.mark.parametrize("config", ["ade20k_infer_mk3", "ade20k_infer_mk1"])
def test_inference_matches_reference_golden(config, native_conv):
    """BASELINE configs[0]: `mode='inference'` on the GPU (fused normalise+pack prologue, K1, K2) against what the
    unmodified reference produced on the CPU for the same seeded weights and batch.  Bounds as in the train-step
    test: fp32 convs -> the fp16-operand correlation is the only rounding (measured <= 9e-4 at K = 256); K2 convs ->
    TF32-class error amplified by 1/T = 100 on warp_out."""
    from cocosnet_b200 import data as cdata, ops
    from tests.test_model_parity_cpu import build_inference_model
    gold = np.load(os.path.join(GOLD, "model_%s.npz" % config))
    old_tf32, old_native = torch.backends.cudnn.allow_tf32, ops.NATIVE_CONV
    torch.backends.cudnn.allow_tf32 = False
    ops.NATIVE_CONV = native_conv
    try:
        opt, model = build_inference_model(config, gpu=True)
        batch = cdata.synthetic_batch(opt, 1)
        with torch.no_grad():
            out = model(batch, mode="inference")
    finally:
        torch.backends.cudnn.allow_tf32, ops.NATIVE_CONV = old_tf32, old_native
    warp = _rel(out["warp_out"].cpu().numpy()[:, :, ::4, ::4], gold["warp_out_sub"])
    fake = _rel(out["fake_image"].cpu().numpy()[:, :, ::4, ::4], gold["fake_image_sub"])
    assert warp < (2e-2 if native_conv else 2e-3), warp
    assert fake < (8e-3 if native_conv else 2e-3), fake
