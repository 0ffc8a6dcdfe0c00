"""Test infrastructure (NOT product code): import the UNMODIFIED reference from
/root/reference on CPU, with stubs for the pieces that are not vendored there.

Only `tests/golden/make_golden.py` (run in the build container, where
/root/reference exists) uses this.  Nothing here travels to the GPU box at run
time: the golden vectors it produces are committed under tests/golden/.

Stubs / patches:
  * models.networks.sync_batchnorm  -> torch BatchNorm / DataParallel aliases
    (un-vendored third party package, README.md:28-34 of the reference)
  * matplotlib, skimage              -> empty modules (imported, unused on path)
  * cwd must be /root/reference      (util/util.py:22 loads ./util/color150.mat)
  * torch.load('models/vgg19_conv.pth') -> seeded random VGG state_dict
  * torch.optim.Adam betas=(0, 0.9)  -> float cast (pix2pix_model.py:101)
"""
import argparse
import contextlib
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("COCOS_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models", "networks"))


def _install_stubs():
    if "models.networks.sync_batchnorm" in sys.modules:
        return
    sb = types.ModuleType("models.networks.sync_batchnorm")
    sb.SynchronizedBatchNorm2d = nn.BatchNorm2d
    sb.SynchronizedBatchNorm1d = nn.BatchNorm1d
    sb.DataParallelWithCallback = nn.DataParallel
    sys.modules["models.networks.sync_batchnorm"] = sb
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.feature"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                if name == "skimage":
                    m.feature = types.ModuleType("skimage.feature")
                sys.modules[name] = m


@contextlib.contextmanager
def reference_imported():
    """Context: cwd=/root/reference, reference on sys.path, stubs installed."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    old_cwd = os.getcwd()
    old_argv = sys.argv
    os.chdir(REF_ROOT)
    sys.path.insert(0, REF_ROOT)
    sys.argv = [old_argv[0]]
    _install_stubs()
    try:
        yield
    finally:
        os.chdir(old_cwd)
        sys.argv = old_argv
        try:
            sys.path.remove(REF_ROOT)
        except ValueError:
            pass


class _CpuIds(list):
    """gpu_ids with len()==0 whose [0] is 'cpu' (pix2pix_model.py:35)."""

    def __getitem__(self, i):
        return "cpu"


def make_opt(argv, is_train):
    """Build the reference `opt` Namespace the way BaseOptions.parse() does
    (options/base_options.py:173-202) but without writing checkpoints."""
    from options.train_options import TrainOptions
    from options.test_options import TestOptions

    o = TrainOptions() if is_train else TestOptions()
    old = sys.argv
    sys.argv = [old[0]] + list(argv)
    try:
        opt = o.gather_options()
    finally:
        sys.argv = old
    opt.isTrain = is_train
    opt.semantic_nc = opt.label_nc + (1 if opt.contain_dontcare_label else 0)
    opt.gpu_ids = _CpuIds()
    return opt


def seeded_vgg_state_dict(seed=7):
    from models.networks.correspondence import VGG19_feature_color_torchversion

    g = torch.Generator().manual_seed(seed)
    net = VGG19_feature_color_torchversion()
    sd = net.state_dict()
    for k, v in sd.items():
        if k.endswith("weight"):
            fan_in = v[0].numel()
            v.copy_(torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5)
        else:
            v.zero_()
    return sd


@contextlib.contextmanager
def patched_for_cpu_training():
    real_load = torch.load
    real_adam = torch.optim.Adam

    def fake_load(path, *a, **k):
        if isinstance(path, str) and path.endswith("vgg19_conv.pth"):
            return seeded_vgg_state_dict()
        return real_load(path, *a, **k)

    class Adam(real_adam):
        def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), **kw):
            super().__init__(params, lr=lr, betas=(float(betas[0]), float(betas[1])), **kw)

    torch.load = fake_load
    torch.optim.Adam = Adam
    try:
        yield
    finally:
        torch.load = real_load
        torch.optim.Adam = real_adam
