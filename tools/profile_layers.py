"""Per-layer view of the train step: every call into the native NHWC backend (tap convolutions, backward-weights,
weight packs, SPADE / instance-norm / pack kernels) is bracketed by CUDA events during ONE eager iteration and the
calls are aggregated by shape signature: milliseconds, algorithmic TFLOP/s (tensor kernels) or GB/s (HBM-bound
kernels).  python tools/profile_layers.py [--b 8] [--rows 60]"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cocosnet_b200 import data as cdata  # noqa: E402
from cocosnet_b200 import nhwc  # noqa: E402
from cocosnet_b200.trainer import Pix2PixTrainer  # noqa: E402

RECORDS = []
_BYTES = {nhwc.F16: 2, nhwc.BF16: 2, nhwc.F32: 4}


def _nt_bytes(nt, C=None):
    return nt.t.numel() // nt.Cs * (C if C is not None else nt.Cs) * _BYTES[nt.kind]


def _sig_conv(x, w, bias, res, y, d):
    work = 2.0 * d["B"] * d["H"] * d["W"] * d["Cout"] * len(d["groups"]) * d["kchunks"] * 64
    if d.get("mod") is not None:
        return ("tapconv_spade %dx%d Ca%d->2x%d g%d kc%d B%d" % (d["H"], d["W"], d["Ca"], d["Cout"] // 2, len(d["groups"]),
                                                                 d["kchunks"], d["B"]), work, "flop")
    return ("tapconv%s s%d %dx%d Ca%d->%d g%d kc%d B%d%s" % ("_bf16" if d["bf16"] else "", d["a_stride"], d["H"], d["W"],
                                                          d["Ca"], d["Cout"], len(d["groups"]), d["kchunks"], d["B"],
                                                          " ->nchw" if d["y_kind"] == 0 else ""), work, "flop")


def _sig_wgrad(dy, x, ws, d):
    work = 2.0 * d["B"] * d["H"] * d["W"] * d["Cout"] * d["Cin"] * len(d["groups"])
    return ("tapwgrad s%d %dx%d Cin%d Cout%d g%d B%d" % (d["a_stride"], d["H"], d["W"], d["Cin"], d["Cout"],
                                                        len(d["groups"]), d["B"]), work, "flop")


def _sig_packw(w, dst, rows, rows_alloc, kc, groups, transposed, bf16):
    return ("pack_w %s%s" % (tuple(w.shape), " T" if transposed else ""), w.numel() * 4.0 + dst.numel() * 2.0, "byte")


def _sig_spade_fwd(x, gb, y, mean, rstd, C, pad, slope, eps):
    return ("spade_fwd %dx%d C%d xk%d" % (x.H, x.W, C, x.kind), _nt_bytes(x, C) + _nt_bytes(gb, 2 * C) + y.t.numel() * 2.0, "byte")


def _sig_spade_bwd(dy, x, gb, mean, rstd, dx, dx_acc, dgb, C, pad, slope, gb_W=0):
    return ("spade_bwd %dx%d C%d xk%d" % (x.H, x.W, C, x.kind),
            dy.t.numel() * 2.0 + _nt_bytes(x, C) + _nt_bytes(gb, 2 * C) + (2 if dx_acc else 1) * dx.t.numel() * 2.0 + dgb.t.numel() * 2.0,
            "byte")


def _sig_in_stats(x, stats, C):
    return ("in_stats %dx%d C%d xk%d" % (x.H, x.W, C, x.kind), _nt_bytes(x), "byte")


def _sig_inst_fwd(x, stats, res, slope_ptr, slope, y, y2, eps, C, gb=None, batch_stats=False):
    b = _nt_bytes(x) + y.t.numel() * _BYTES[y.kind] + (res.t.numel() * _BYTES[res.kind] if res else 0) + \
        (y2.t.numel() * 4 if y2 else 0)
    return ("inst_fwd %dx%d C%d xk%d%s" % (x.H, x.W, C, x.kind, " +res" if res else ""), b, "byte")


def _sig_inst_bwd(dy, dy2, x, stats, res, slope_ptr, slope, bstats, dslope, dx, dx_acc, dres, dres_acc, eps, C, gb=None,
                  dgb=None, batch_stats=False, const_stats=False, phase=0):
    b = 2 * (dy.t.numel() * 2.0 + _nt_bytes(x) + (res.t.numel() * _BYTES[res.kind] if res else 0)
             + (dy2.t.numel() * 2.0 if dy2 else 0)) + dx.t.numel() * 2.0 + (dres.t.numel() * 2.0 if dres else 0)
    return ("inst_bwd %dx%d C%d xk%d%s" % (x.H, x.W, C, x.kind, " +res" if res else ""), b, "byte")


def _sig_act_bwd(dy, y, dz, C, act, slope):
    return ("act_bwd %dx%d C%d" % (y.H, y.W, C), dy.t.numel() * 2.0 + _nt_bytes(y) + dz.t.numel() * 2.0, "byte")


def _sig_pack(src, dst, C, f, c_lo=0, c_span=0):
    return ("pack %s -> %dx%d k%d f%d" % (tuple(src.shape[1:]), dst.H, dst.W, dst.kind, f),
            src.numel() * 4.0 / (f * f) + dst.t.numel() // dst.Cs * max(C, c_span) * _BYTES[dst.kind] * (2 if dst.lo else 1), "byte")


def _sig_unpack(src, c_lo, C, dst, cd_lo, f, acc):
    return ("unpack %dx%d C%d k%d f%d%s" % (src.H, src.W, C, src.kind, f, " acc" if acc else ""),
            src.t.numel() // src.Cs * C * _BYTES[src.kind] + src.B * C * src.H * src.W * 4.0 * (2 if acc else 1), "byte")


def _sig_colsum(x2d, kind, Cs, C, rows, out):
    return ("colsum rows%d Cs%d" % (rows, Cs), rows * Cs * _BYTES[kind], "byte")


def _sig_maxpool_fwd(x, y):
    return ("maxpool_fwd %dx%d Cs%d" % (x.H, x.W, x.Cs), x.t.numel() * 2.0 * 1.25, "byte")


def _sig_maxpool_bwd(dy, x, dx):
    return ("maxpool_bwd %dx%d Cs%d" % (x.H, x.W, x.Cs), x.t.numel() * 2.0 * 2.25, "byte")


def _sig_pair_fwd(x, y, w, scale, mode, out):
    return ("pair_loss_fwd %dx%d C%d" % (x.H, x.W, x.C), 2.0 * x.t.numel() * 2, "byte")


def _sig_pair_bwd(x, y, w, scale, mode, g, dx, acc):
    return ("pair_loss_bwd %dx%d C%d" % (x.H, x.W, x.C), (3.0 + (1 if acc else 0)) * x.t.numel() * 2, "byte")


def _sig_cast(x, dst):
    return ("cast_bf16 %dx%d Cs%d" % (x.H, x.W, x.Cs), x.t.numel() * 2.0 + dst.t.numel() * 2.0, "byte")


def _sig_pono(x, C, eps, mean, rstd):
    return ("pono_stats %dx%d C%d xk%d" % (x.H, x.W, C, x.kind), _nt_bytes(x, C), "byte")


SIGS = {"pair_loss_fwd": _sig_pair_fwd, "pair_loss_bwd": _sig_pair_bwd, "cast_bf16": _sig_cast, "pono_stats": _sig_pono,
        "tapconv": _sig_conv, "tapwgrad": _sig_wgrad, "pack_w": _sig_packw, "spade_fwd": _sig_spade_fwd,
        "spade_bwd": _sig_spade_bwd, "in_stats": _sig_in_stats, "inst_fwd": _sig_inst_fwd, "inst_bwd": _sig_inst_bwd,
        "act_bwd": _sig_act_bwd, "pack": _sig_pack, "unpack": _sig_unpack, "colsum": _sig_colsum,
        "maxpool_fwd": _sig_maxpool_fwd, "maxpool_bwd": _sig_maxpool_bwd}


def instrument(be):
    for name, sig in SIGS.items():
        orig = getattr(be, name)

        def wrapped(*a, _orig=orig, _sig=sig, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = _orig(*a, **k)
            e.record()
            RECORDS.append((_sig(*a, **k), s, e))
            return r
        setattr(be, name, wrapped)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=8)
    ap.add_argument("--rows", type=int, default=70)
    args = ap.parse_args()
    opt = bench.make_opt(args.b, gpu=True)
    torch.manual_seed(0)
    trainer = Pix2PixTrainer(opt)
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in cdata.synthetic_batch(opt, args.b).items()}

    def step():
        trainer.run_generator_one_step(batch)
        trainer.run_discriminator_one_step(batch)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    instrument(nhwc.backend())
    step()
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for (name, work, unit), s, e in RECORDS:
        a = agg.setdefault(name, [0, 0.0, 0.0, unit])
        a[0] += 1
        a[1] += s.elapsed_time(e)
        a[2] += work
    total = sum(a[1] for a in agg.values())
    fam = collections.defaultdict(float)
    for name, a in agg.items():
        fam[name.split()[0]] += a[1]
    print("native NHWC backend calls in one iteration (B=%d): %d calls, %.2f ms (event-bracketed, includes launch gaps)"
          % (args.b, len(RECORDS), total))
    print("--- by kernel family ---")
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
        print("%8.3f ms  %s" % (v, k))
    print("--- by shape signature: ms, calls, algorithmic TFLOP/s or GB/s ---")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.rows]:
        rate = a[2] / a[1] / 1e9 if a[3] == "flop" else a[2] / a[1] / 1e6
        print("%8.3f ms x%-3d %7.0f %s  %s" % (a[1], a[0], rate, "TF/s" if a[3] == "flop" else "GB/s", name))


if __name__ == "__main__":
    main()
