"""K2 forward vs cuDNN (TF32) on the SPADE-block shapes of the ade20k generator, B=8."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocosnet_b200 import ops

def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it

torch.backends.cudnn.benchmark = True
for (cin, cout, hw) in ((1024, 1024, 8), (1024, 1024, 16), (512, 512, 32), (256, 256, 64), (128, 128, 128), (64, 64, 256),
                        (154, 128, 256), (128, 128, 256), (128, 2048, 16), (512, 512, 64)):
    b = 8
    x = torch.randn(b, cin, hw + 2, hw + 2, device="cuda")
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.02
    bias = torch.zeros(cout, device="cuda")
    fl = 2.0 * b * hw * hw * cin * cout * 9
    m_cudnn = t(lambda: F.conv2d(x, w, bias))
    m_mine = t(lambda: ops.conv_fwd_native(x, w, bias, True))
    cp = ops.round_up(cin, 64)
    x16 = ops.pack_rows(x.view(b, cin, -1), kp=cp); wt = ops.pack_conv_weight(w)
    y = torch.empty(b, cout, hw, hw, device="cuda")
    from cocosnet_b200 import _lib
    def kern():
        _lib.check(_lib.lib().cocos_conv_fwd(x16.data_ptr(), wt.data_ptr(), bias.data_ptr(), y.data_ptr(), b, hw, hw, hw + 2, hw + 2, cp, cout, 3, 0, 0,
                                             torch.cuda.current_stream().cuda_stream), "conv")
    m_kern = t(kern)
    print("cin %4d cout %4d %3dx%-3d  cudnn tf32 %.3f ms (%.0f TF)  native total %.3f ms  kernel only %.3f ms (%.0f TF)" % (
        cin, cout, hw, hw, m_cudnn, fl / m_cudnn / 1e9, m_mine, m_kern, fl / m_kern / 1e9), flush=True)
