"""Scaling of the barrier skeleton with the number of key tiles / CTAs (timing only)."""
import os, subprocess, sys
CODE = """
import sys, torch
sys.path.insert(0, '.')
from cocosnet_b200 import ops
b, nq, nk, kd, cv = %d, %d, %d, 256, 3
q = torch.randn(b, kd, nq, device='cuda'); q = q / q.norm(dim=1, keepdim=True)
k = torch.randn(b, kd, nk, device='cuda'); k = k / k.norm(dim=1, keepdim=True)
v = torch.rand(b, cv, nk, device='cuda')
q16, k16, vt = ops.pack_rows(q), ops.pack_rows(k), ops.pack_v(v)
for _ in range(3): ops.corr_warp_fwd(q16, k16, vt, cv, nk, 100.0, v32=v)
ts = []
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.corr_warp_fwd(q16, k16, vt, cv, nk, 100.0, v32=v); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ts.sort(); print('%%.4f' %% ts[len(ts)//2])
"""
for dbg in (159, 0):
    for b, nq, nk in ((4, 4096, 4096), (8, 4096, 4096), (4, 4096, 512), (4, 4096, 1024), (4, 4096, 2048), (4, 4096, 8192), (1, 4096, 4096), (1, 128, 4096), (1, 128, 16384)):
        env = dict(os.environ, COCOS_K1_DBG=str(dbg))
        r = subprocess.run([sys.executable, "-c", CODE % (b, nq, nk)], env=env, capture_output=True, text=True)
        print("dbg=%3d b=%d nq=%5d nk=%5d ctas=%4d tiles=%3d ms=%s %s" % (dbg, b, nq, nk, b * nq // 128, nk // 128, r.stdout.strip(), r.stderr.strip()[-200:] if r.returncode else ""), flush=True)
