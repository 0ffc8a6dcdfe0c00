import os, subprocess, sys
CODE = """
import sys, torch
sys.path.insert(0, '.')
from cocosnet_b200 import ops
a = torch.randn(1, 8192, 8192, device='cuda').half(); b = torch.randn(1, 8192, 8192, device='cuda').half()
c = ops.gemm_f16(a, b); ref = a[0].float() @ b[0].float().t()
print('rel %.2e' % float((c[0]-ref).norm()/ref.norm()))
for _ in range(3): ops.gemm_f16(a, b)
ts=[]
for _ in range(8):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.gemm_f16(a, b); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ts.sort(); ms = ts[len(ts)//2]; print('ms %.3f tflops %.0f' % (ms, 2*8192**3/ms/1e9))
ts=[]
for _ in range(8):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); torch.matmul(a[0], b[0].t()); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ts.sort(); ms = ts[len(ts)//2]; print('cublas ms %.3f tflops %.0f' % (ms, 2*8192**3/ms/1e9))
"""
for bn, two in ((128, 0), (256, 0), (256, 1), (256, 0), (256, 1)):
    r = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, COCOS_GEMM_BN=str(bn), COCOS_GEMM_2CTA=str(two)),
                       capture_output=True, text=True, timeout=300)
    print("BN=%d 2cta=%d:" % (bn, two), r.stdout.replace("\n", " | "), r.stderr[-600:] if r.returncode else "")
