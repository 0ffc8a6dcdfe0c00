"""CPU port of the full reference train step.  TEST INFRASTRUCTURE ONLY
(tests/, smoke(), bench.py cpu_baseline / --impl reference).

The reference is a Python program that cannot travel to the GPU box, so the
"reference arm" there is this port: the host-side mirror modules of
cocosnet_b200 (bit-identical seeded init and state_dict keys to the reference,
tests/test_model_parity_cpu.py) with the fused CUDA primitive swapped for the
reference's own unfused torch expressions:

    f = matmul(theta^T, phi) / T; P = softmax(f); y = matmul(P, ref)
    (reference correspondence.py:291, 304, 307, 318; architecture.py:122-125)

Pinned against the real reference in the build container: the golden loss
values in tests/golden/model_*.npz were produced by the unmodified reference
(tests/golden/make_golden_model.py) and this port reproduces them.
"""
import contextlib

import torch


def attend_unfused(q, k, v, scale, precision="fp16"):
    """q [B,Kd,Nq], k [B,Kd,Nk], v [B,Cv,Nk] -> [B,Cv,Nq]; plain fp32 torch ops."""
    f = torch.matmul(q.permute(0, 2, 1), k) * scale
    p = torch.softmax(f, dim=-1)
    return torch.matmul(p, v.permute(0, 2, 1)).permute(0, 2, 1)


def raw_correlation_unfused(q, k, scale):
    return torch.matmul(q.permute(0, 2, 1), k) * scale


@contextlib.contextmanager
def cpu_reference_mode():
    """Route the mirror modules through the unfused torch expressions (CPU)."""
    from cocosnet_b200 import corr
    saved = corr.attend, corr.raw_correlation
    corr.attend, corr.raw_correlation = attend_unfused, raw_correlation_unfused
    try:
        yield
    finally:
        corr.attend, corr.raw_correlation = saved


def seeded_vgg_state_dict(seed=7):
    """The seeded stand-in for models/vgg19_conv.pth lives in the product package (cocosnet_b200/data.py)."""
    from cocosnet_b200.data import seeded_vgg_state_dict as f
    return f(seed)
