"""Launch the fused correspondence kernel a few times (for ncu captures).
python tools/run_k1_once.py [--b 8] [--kd 256] [--cv 3] [--bwd]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocosnet_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=8)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--kd", type=int, default=256)
ap.add_argument("--cv", type=int, default=3)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
q = torch.randn(a.b, a.kd, a.n, device="cuda")
q = q / q.norm(dim=1, keepdim=True)
k = torch.randn(a.b, a.kd, a.n, device="cuda")
k = k / k.norm(dim=1, keepdim=True)
v = torch.rand(a.b, a.cv, a.n, device="cuda")
q16, k16, vt = ops.pack_rows(q), ops.pack_rows(k), ops.pack_v(v)
for _ in range(a.reps):
    out, lse, _ = ops.corr_warp_fwd(q16, k16, vt, a.cv, a.n, 100.0, v32=(v if a.cv <= 4 else None))
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
