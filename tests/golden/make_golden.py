"""Generate golden vectors by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/tail_<case>.npz.  Each file holds the outputs of
NoVGGCorrespondence.forward (reference correspondence.py:222-374) when the
outputs of its theta / phi 1x1 convs are replaced (forward hooks) by the seeded
tensors of tests/golden/cases.py, i.e. the reference's own tail code run on
known inputs.  Big outputs are stored subsampled (see `_store`).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_harness as rh  # noqa: E402
from tests.golden import cases  # noqa: E402


def _store(out):
    res = {}
    for k, v in out.items():
        if not torch.is_tensor(v):
            continue
        a = v.detach().numpy().astype(np.float32)
        if k == "warp_mask" and a.shape[1] > 32:
            res[k + "_chsum"] = a.sum(axis=1)
            a = a[:, ::19]
        res[k] = a
    return res


def run_case(name):
    spec = cases.TAIL_CASES[name]
    inp = cases.tail_inputs(name)
    with rh.reference_imported():
        opt = rh.make_opt(spec["argv"], bool(spec.get("train")))
        if spec.get("train"):
            opt.novgg_featpair = 0
        import models.networks as networks
        torch.manual_seed(0)
        net = networks.define_Corr(opt)
        net.eval() if not spec.get("train") else net.train()
        th = torch.from_numpy(inp["theta"])
        ph = torch.from_numpy(inp["phi"])
        h1 = net.theta.register_forward_hook(lambda m, i, o: th)
        h2 = net.phi.register_forward_hook(lambda m, i, o: ph)
        args = (torch.from_numpy(inp["ref_img"]), torch.from_numpy(inp["real_img"]),
                torch.from_numpy(inp["seg"]), torch.from_numpy(inp["ref_seg"]))
        with torch.no_grad():
            out = net(*args)
            res = _store(out)
            if inp["theta"].shape[-1] <= 24:
                corr = net(*args, return_corr=True)
                res["corr"] = corr.numpy().astype(np.float32)
        h1.remove(); h2.remove()
    path = os.path.join(HERE, "tail_%s.npz" % name)
    np.savez_compressed(path, **res)
    print(name, {k: v.shape for k, v in res.items()}, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    names = sys.argv[1:] or list(cases.TAIL_CASES)
    for n in names:
        run_case(n)
