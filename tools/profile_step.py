"""Where does the train step spend its GPU time?  torch.profiler kernel table
for one G+D step at batch B (after warm-up).  python tools/profile_step.py --b 8"""
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
    ap.add_argument("--rows", type=int, default=45)
    ap.add_argument("--channels_last", action="store_true")
    ap.add_argument("--no_table", action="store_true")
    ap.add_argument("--cudnn_benchmark", action="store_true")
    args = ap.parse_args()
    opt = bench.make_opt(args.b, gpu=True)
    opt.no_cudnn_benchmark = not args.cudnn_benchmark
    opt.channels_last = args.channels_last
    torch.manual_seed(0)
    trainer = Pix2PixTrainer(opt)
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in cdata.synthetic_batch(opt, args.b).items()}

    def step():
        trainer.run_generator_one_step(batch)
        trainer.run_discriminator_one_step(batch)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        step()
    e.record()
    torch.cuda.synchronize()
    print("step ms: %.2f  (%.2f img/s at B=%d)  peak mem %.1f GB" % (s.elapsed_time(e) / 3, args.b * 3e3 / s.elapsed_time(e),
                                                              args.b, torch.cuda.max_memory_allocated() / 2 ** 30))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    ka = prof.key_averages()
    busy = sum(k.self_device_time_total for k in ka) / 1e3
    print("GPU busy (sum of kernel durations) %.2f ms; launches %d" % (busy, sum(k.count for k in ka if k.self_device_time_total > 0)))
    if args.no_table:
        return
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=args.rows, max_name_column_width=70))


if __name__ == "__main__":
    main()
