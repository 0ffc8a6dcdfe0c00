"""Which torch (ATen) ops are left in the train step, and on which tensors?  One eager iteration under torch.profiler
with record_shapes: CUDA time per (op, input shapes), hand-written kernels excluded.
python tools/profile_aten_ops.py [--b 8] [--rows 70]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cocosnet_b200 import data as cdata  # noqa: E402
from cocosnet_b200.trainer import Pix2PixTrainer  # noqa: E402


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
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
        step()
        torch.cuda.synchronize()
    rows = []
    for k in prof.key_averages(group_by_input_shape=True):
        t = k.self_device_time_total / 1e3
        if t <= 0 or not k.key.startswith("aten::"):
            continue
        rows.append((t, k.count, k.key, str(k.input_shapes)[:150]))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    print("ATen ops with device time in one iteration (B=%d): %.2f ms over %d calls" % (args.b, total, sum(r[1] for r in rows)))
    byop = {}
    for t, n, key, _ in rows:
        a = byop.setdefault(key, [0.0, 0])
        a[0] += t
        a[1] += n
    print("--- by op ---")
    for key, (t, n) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:25]:
        print("%8.3f ms x%-4d %s" % (t, n, key))
    print("--- by op and input shapes ---")
    for t, n, key, shapes in rows[:args.rows]:
        print("%8.3f ms x%-3d %-28s %s" % (t, n, key, shapes))


if __name__ == "__main__":
    main()
