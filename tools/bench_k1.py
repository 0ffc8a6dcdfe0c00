"""Micro-benchmark of the hot kernels (CUDA events, L2 flushed between runs).
python tools/bench_k1.py [--b 8]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocosnet_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=8)
    ap.add_argument("--n", type=int, default=4096)
    args = ap.parse_args()
    b, n = args.b, args.n
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    res = {}
    for kd, cv in ((256, 3), (2304, 3), (256, 154), (768, 3)):
        q = torch.randn(b, kd, n, device="cuda")
        q = q / q.norm(dim=1, keepdim=True)
        k = torch.randn(b, kd, n, device="cuda")
        k = k / k.norm(dim=1, keepdim=True)
        v = torch.rand(b, cv, n, device="cuda")
        q16, k16, vt = ops.pack_rows(q), ops.pack_rows(k), ops.pack_v(v)
        med, best = timeit(lambda: ops.corr_warp_fwd(q16, k16, vt, cv, n, 100.0, v32=(v if cv <= 4 else None)), flush=flush)
        fl = b * (2.0 * n * n * kd + 2.0 * n * n * cv)
        res["k1_fwd_kd%d_cv%d" % (kd, cv)] = dict(ms=med, best_ms=best, tflops=fl / med / 1e9,
                                                 us_per_img=med * 1e3 / b)
        med, best = timeit(lambda: ops.pack_rows(q), flush=flush)
        res["pack_rows_kd%d" % kd] = dict(ms=med, gbs=b * kd * n * 6 / med / 1e6)
        if cv == 3:
            # unfused torch path for context (cuBLAS fp32/TF32 matmul + softmax + matmul)
            def unfused():
                f = torch.matmul(q.transpose(1, 2), k) * 100.0
                p = torch.softmax(f, -1)
                return torch.matmul(p, v.transpose(1, 2))
            med_u, _ = timeit(unfused, iters=5, flush=flush)
            res["torch_unfused_kd%d" % kd] = dict(ms=med_u, tflops=fl / med_u / 1e9)
        del q, k, v, q16, k16, vt
    a = torch.randn(1, 8192, 8192, device="cuda").half()
    bb = torch.randn(1, 8192, 8192, device="cuda").half()
    med, best = timeit(lambda: ops.gemm_f16(a, bb), iters=5)
    res["gemm_8192"] = dict(ms=med, tflops=2 * 8192 ** 3 / med / 1e9)
    med, _ = timeit(lambda: torch.matmul(a[0], bb[0].t()), iters=5)
    res["cublas_fp16_8192"] = dict(ms=med, tflops=2 * 8192 ** 3 / med / 1e9)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
