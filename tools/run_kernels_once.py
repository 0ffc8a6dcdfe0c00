"""Launch the step's main kernels a few times each on the shapes of the ade20k batch-8 iteration, for `ncu -k regex:...`
captures (tools/ncu_kernels.sh).  python tools/run_kernels_once.py <family> ; families: tapconv tapwgrad spade inst
corr_bwd prologue pack"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocosnet_b200 import nhwc, ops  # noqa: E402


def rnd(*shape, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g)


def main():
    fam = sys.argv[1]
    reps = 3
    if fam == "tapconv":
        # adaptor SPADE-block conv 512->512 @64x64, B=8, split operands (27 groups) and its single-term backward-data
        x = nhwc.pack(rnd(8, 512, 64, 64), nhwc.F16, pad=1, split=True)
        w = rnd(512, 512, 3, 3, seed=1) * 0.02
        dy = nhwc.pack(rnd(8, 512, 64, 64, seed=2), nhwc.BF16)
        for _ in range(reps):
            nhwc.conv(x, w, None, out_kind=nhwc.F32)
            nhwc.conv_dgrad(dy, w, (66, 66), in_pad=1)
    elif fam == "tapwgrad":
        x = nhwc.pack(rnd(8, 512, 64, 64), nhwc.F16, pad=1, split=True)
        dy = nhwc.pack(rnd(8, 512, 64, 64, seed=2), nhwc.BF16)
        for _ in range(reps):
            nhwc.conv_wgrad(dy, x, 3)
    elif fam == "spade":
        x = nhwc.pack(rnd(8, 512, 64, 64), nhwc.F32)
        gb = nhwc.pack(rnd(8, 1024, 64, 64, seed=1), nhwc.F32)
        dy = nhwc.pack(rnd(8, 512, 64, 64, seed=2), nhwc.BF16, pad=1)
        for _ in range(reps):
            y, mean, rstd = nhwc.spade_mod_fwd(x, gb, 512, pad=1, slope=0.2, split_out=True)
            nhwc.spade_mod_bwd(dy, x, gb, mean, rstd, 512, 1, 0.2)
    elif fam == "inst":
        x = nhwc.pack(rnd(16, 408, 64, 64), nhwc.F32)
        res = nhwc.pack(rnd(16, 408, 64, 64, seed=1), nhwc.F32)
        dy = nhwc.pack(rnd(16, 408, 64, 64, seed=2), nhwc.BF16, pad=1)
        a = torch.full((1,), 0.25, device="cuda")
        for _ in range(reps):
            st = nhwc.in_stats(x)
            nhwc.inst_act_fwd(x, st, slope_ptr=a, res=res, out_pad=1, split_out=True, want_raw=True)
            nhwc.inst_act_bwd(dy, x, st, slope_ptr=a, res=res, want_dres=True, dslope=torch.zeros((), device="cuda"))
    elif fam == "corr_bwd":
        from cocosnet_b200 import corr
        theta, phi = rnd(8, 256, 64, 64).requires_grad_(True), rnd(8, 256, 64, 64, seed=1).requires_grad_(True)
        ref = torch.rand(8, 3, 256, 256, device="cuda")
        for _ in range(reps):
            y, _ = corr.correspondence_tail(theta, phi, ref, match_kernel=3, pono_c=True)
            y.square().mean().backward()
    elif fam == "pack":
        sem = torch.zeros(8, 151, 256, 256, device="cuda")
        for _ in range(reps):
            nt = nhwc.pack(sem, nhwc.F16, pad=1)
            nhwc.unpack(nhwc.pack(rnd(8, 64, 256, 256), nhwc.F16))
            nhwc.as_bf16(nt)
            nt.bf = None
    else:
        raise SystemExit("unknown family " + fam)
    torch.cuda.synchronize()
    print("ok", fam)


if __name__ == "__main__":
    main()
