"""Microbenchmark of the NHWC tap-convolution kernels (forward / backward-data / backward-weights) on layer shapes of
the ade20k step at batch 8.  CUDA events, L2 flushed between launches.  python tools/bench_tapconv.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocosnet_b200 import nhwc  # noqa: E402

SHAPES = [  # name, ks, stride, padding, cin, cout, b, h, w, halo
    ("spade conv 512->512 @64", 3, 1, 0, 512, 512, 8, 64, 64, 1),
    ("gamma/beta 128->512 @64", 3, 1, 0, 128, 512, 8, 64, 64, 1),
    ("gamma/beta 128->128 @256", 3, 1, 0, 128, 128, 8, 256, 256, 1),
    ("mlp_shared 154->128 @256", 3, 1, 0, 154, 128, 8, 256, 256, 1),
    ("up_3 conv 64->64 @256", 3, 1, 0, 64, 64, 8, 256, 256, 1),
    ("G_middle 1024->1024 @16", 3, 1, 0, 1024, 1024, 8, 16, 16, 1),
    ("resblock 407->407 @64 split", 3, 1, 0, 407, 407, 8, 64, 64, 1),
    ("patchgan 154->64 s2 @256", 4, 2, 1, 154, 64, 16, 256, 256, 0),
    ("patchgan 128->256 s2 @64", 4, 2, 1, 128, 256, 16, 64, 64, 0),
    ("vgg 256->256 @64", 3, 1, 1, 256, 256, 8, 64, 64, 0),
]


def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    print("%-30s %10s %10s %10s %10s  (ms, TFLOP/s algorithmic; weight packing included in every call; wgrad_bf16x = "
          "backward-weights with X already bf16: no in-kernel fp16->bf16 conversion)" % ("layer", "fwd", "dgrad", "wgrad", "wgrad_bf16x"))
    for name, ks, stride, padding, cin, cout, b, h, w, halo in SHAPES:
        split = "split" in name
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(b, cin, h, w, device="cuda", generator=g)
        wt = torch.randn(cout, cin, ks, ks, device="cuda", generator=g) * 0.05
        xin = nhwc.pack(x, nhwc.F16, pad=halo, split=split)
        hin, win = h + 2 * halo, w + 2 * halo
        ho, wo = nhwc.conv_out_size(hin, ks, padding, stride), nhwc.conv_out_size(win, ks, padding, stride)
        dy = nhwc.pack(torch.randn(b, cout, ho, wo, device="cuda", generator=g), nhwc.BF16)
        flops = 2.0 * b * ho * wo * cout * cin * ks * ks
        t_f = timeit(lambda: nhwc.conv(xin, wt, None, stride=stride, padding=padding, out_kind=nhwc.F32 if split else nhwc.F16))
        t_d = timeit(lambda: nhwc.conv_dgrad(dy, wt, (hin, win), stride=stride, padding=padding, in_pad=halo))
        t_w = timeit(lambda: nhwc.conv_wgrad(dy, xin, ks, stride=stride, padding=padding))
        xb = nhwc.pack(x, nhwc.BF16, pad=halo)
        t_wb = timeit(lambda: nhwc.conv_wgrad(dy, xb, ks, stride=stride, padding=padding))
        print("%-30s %5.3f/%4.0f %5.3f/%4.0f %5.3f/%4.0f %5.3f/%4.0f" % (name, t_f, flops / t_f / 1e9, t_d, flops / t_d / 1e9,
                                                                        t_w, flops / t_w / 1e9, t_wb, flops / t_wb / 1e9))


if __name__ == "__main__":
    main()
