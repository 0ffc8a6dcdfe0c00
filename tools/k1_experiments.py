"""Timing experiments on the fused kernel (COCOS_K1_DBG variants are NOT numerically valid)."""
import os, subprocess, sys
for dbg, promo in ((0, 256), (0, 128), (0, 0), (31, 256), (31, 0), (31 | 128, 256), (15 | 128, 256), (128, 256)):
    env = dict(os.environ, COCOS_K1_DBG=str(dbg), COCOS_TMA_L2PROMO=str(promo))
    r = subprocess.run([sys.executable, "-c", """
import sys, torch
sys.path.insert(0, '.')
from cocosnet_b200 import ops
b, n, kd, cv = 8, 4096, 256, 3
q = torch.randn(b, kd, n, device='cuda'); q = q / q.norm(dim=1, keepdim=True)
k = torch.randn(b, kd, n, device='cuda'); k = k / k.norm(dim=1, keepdim=True)
v = torch.rand(b, cv, n, device='cuda')
q16, k16, vt = ops.pack_rows(q), ops.pack_rows(k), ops.pack_v(v)
for _ in range(3): ops.corr_warp_fwd(q16, k16, vt, cv, n, 100.0, v32=(v if cv <= 4 else None))
ts = []
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.corr_warp_fwd(q16, k16, vt, cv, n, 100.0, v32=(v if cv <= 4 else None)); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ts.sort(); print('%.4f' % ts[len(ts)//2])
"""], env=env, capture_output=True, text=True)
    print("dbg=%3d promo=%3d ms=%s %s" % (dbg, promo, r.stdout.strip(), r.stderr.strip()[-200:] if r.returncode else ""))
