"""Where does the GPU default path leave the fp32 reference expressions?  The same seeded model and batch run (a) on
the CPU through the mirror's plain torch modules and (b) on the GPU through the kernels; relative L2 of the
intermediate tensors the model exposes.  python tools/parity_trace.py celebahq_train"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cocosnet_b200 import data as cdata  # noqa: E402
import test_gpu_model as G  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else "celebahq_train"
    outs = {}
    for gpu in (False, True):
        opt, model = G._build(gpu=gpu, config=config)
        batch = cdata.synthetic_batch(opt, 1)
        grab = {}
        corr = model.net["netCorr"]
        hooks = [corr.theta.register_forward_hook(lambda m, i, o: grab.__setitem__("theta_in", i[0].detach())),
                 corr.adaptive_model_seg.register_forward_hook(lambda m, i, o: grab.__setitem__("adapt_seg", o.detach())),
                 corr.adaptive_model_img.register_forward_hook(lambda m, i, o: grab.setdefault("adapt_img", o.detach()))]
        with torch.no_grad():
            if gpu:
                g_losses, out = model(batch, mode="generator")
            else:
                from oracle import torch_port
                with torch_port.cpu_reference_mode():
                    g_losses, out = model(batch, mode="generator")
        for h in hooks:
            h.remove()
        grab.update({k: out[k].detach() for k in ("warp_out", "fake_image") if out.get(k) is not None})
        outs[gpu] = grab
    for k in outs[False]:
        if k in outs[True]:
            print("%-12s rel L2 %.3e" % (k, rel(outs[True][k], outs[False][k])))
        else:
            print("%-12s not captured on the GPU path (fused away)" % k)


if __name__ == "__main__":
    main()
