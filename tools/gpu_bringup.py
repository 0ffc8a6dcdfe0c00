"""Stage-by-stage GPU bring-up diagnostics (each stage in its own process so a
trapped kernel does not poison the next).  Usage: python tools/gpu_bringup.py [stage ...]
Not a pytest test, but test infrastructure all the same: it uses the CPU oracle as the checker and prints numbers
that localise descriptor / pipeline bugs."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stage_pack():
    import torch
    from cocosnet_b200 import ops
    x = torch.randn(2, 70, 100, device="cuda")
    y = ops.pack_rows(x)
    ref = torch.zeros(2, 100, 128, device="cuda", dtype=torch.float16)
    ref[:, :, :70] = x.permute(0, 2, 1).half()
    print("pack_rows exact:", torch.equal(y, ref))
    y3 = ops.pack_rows(x, split=1)
    hi = x.half()
    lo = (x - hi.float()).half()
    print("pack_rows split:", torch.equal(y3[:, :, :70], hi.permute(0, 2, 1)),
          torch.equal(y3[:, :, 128:198], lo.permute(0, 2, 1)), torch.equal(y3[:, :, 256:326], hi.permute(0, 2, 1)))
    v = torch.randn(2, 3, 100, device="cuda")
    pv = ops.pack_v(v)
    print("pack_v:", pv.shape, torch.equal(pv[:, :3, :100], v.half()), float(pv[:, 3:].abs().max()),
          float(pv[:, :, 100:].abs().max()))


def _gemm_case(b, m, n, k):
    import torch
    from cocosnet_b200 import ops
    a = (torch.randn(b, m, k, device="cuda") / k ** 0.5).half()
    bb = torch.randn(b, n, k, device="cuda").half()
    c = ops.gemm_f16(a, bb)
    torch.cuda.synchronize()
    ref = a.float() @ bb.float().transpose(1, 2)
    err = (c - ref).abs()
    rel = float(err.norm() / ref.norm())
    print("gemm b%d %dx%dx%d rel %.3e max %.3e" % (b, m, n, k, rel, float(err.max())))
    if rel > 1e-3:
        e = err[0]
        rows = (e.max(dim=1).values > 1e-2).nonzero().flatten()[:16].tolist()
        cols = (e.max(dim=0).values > 1e-2).nonzero().flatten()[:16].tolist()
        print("  bad rows (first 16):", rows, " bad cols:", cols)
        print("  c[0,:4,:8]\n", c[0, :4, :8], "\n  ref\n", ref[0, :4, :8])


def stage_gemm():
    _gemm_case(1, 128, 128, 64)
    _gemm_case(1, 128, 128, 256)
    _gemm_case(2, 256, 384, 512)
    _gemm_case(1, 200, 72, 96)


def _fwd_case(b, nq, nk, kd, cv, scale, peaky=False, dump=False):
    import numpy as np
    import torch
    from cocosnet_b200 import ops
    from oracle import corr_oracle as oc
    g = torch.Generator(device="cpu").manual_seed(nq * 7 + nk)
    q = torch.randn(b, kd, nq, generator=g)
    k = torch.randn(b, kd, nk, generator=g)
    if peaky and nq == nk:
        perm = torch.randperm(nk, generator=g)
        k = q[:, :, perm] + 0.05 * torch.randn(b, kd, nk, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    k = k / k.norm(dim=1, keepdim=True)
    v = torch.rand(b, cv, nk, generator=g) * 2 - 1
    q16 = ops.pack_rows(q.cuda().contiguous())
    k16 = ops.pack_rows(k.cuda().contiguous())
    vt = ops.pack_v(v.cuda().contiguous())
    out, lse, corr = ops.corr_warp_fwd(q16, k16, vt, cv, nk, scale, want_lse=True, want_corr=dump)
    torch.cuda.synchronize()
    # oracle on the fp16-rounded operands (isolates kernel error from rounding)
    qr = q16.float().cpu().numpy()[:, :, :kd]
    kr = k16.float().cpu().numpy()[:, :, :kd]
    vr = vt.float().cpu().numpy()[:, :cv, :nk].transpose(0, 2, 1)
    o_ref, lse_ref = oc.attend(qr, kr, vr, scale)
    o_ref = o_ref.transpose(0, 2, 1)
    o = out.cpu().numpy()
    rel = np.linalg.norm(o - o_ref) / np.linalg.norm(o_ref)
    o_true, _ = oc.attend(q.numpy().transpose(0, 2, 1), k.numpy().transpose(0, 2, 1), v.numpy().transpose(0, 2, 1), scale)
    rel_true = np.linalg.norm(o - o_true.transpose(0, 2, 1)) / np.linalg.norm(o_true)
    print("fwd b%d nq%d nk%d kd%d cv%d scale%g peaky%d: rel(vs fp16-operand oracle) %.3e  rel(vs fp64 inputs) %.3e  lse maxerr %.3e  nan %d"
          % (b, nq, nk, kd, cv, scale, peaky, rel, rel_true, np.abs(lse.cpu().numpy() - lse_ref).max(), int(np.isnan(o).sum())))
    if dump:
        z = (qr @ kr.transpose(0, 2, 1)) * scale
        ce = np.abs(corr.cpu().numpy() - z)
        print("   corr dump maxerr %.3e (logit range %.1f)" % (ce.max(), np.abs(z).max()))
        if ce.max() > 1e-2:
            bad = np.argwhere(ce[0] > 1e-2)
            print("   first bad (row,col):", bad[:10].tolist(), " n_bad", len(bad))
            print(corr[0, :4, :8].cpu().numpy(), "\n", z[0, :4, :8])
    if rel > 2e-3:
        print("   out[0,:, :6]", o[0, :, :6], "\n   ref", o_ref[0, :, :6])


def stage_fwd_small():
    _fwd_case(1, 128, 128, 64, 3, 1.0, dump=True)
    _fwd_case(1, 128, 128, 64, 3, 100.0, dump=True)
    _fwd_case(1, 256, 384, 128, 3, 100.0, dump=True)


def stage_fwd_ragged():
    _fwd_case(2, 200, 300, 64, 5, 100.0, dump=True)
    _fwd_case(1, 576, 576, 256, 3, 100.0, peaky=True)
    _fwd_case(1, 512, 512, 320, 20, 100.0)   # streamed-Q path (Kd > 256)
    _fwd_case(1, 256, 256, 64, 154, 100.0)   # wide V (direct mask)


def stage_fwd_full():
    _fwd_case(1, 4096, 4096, 256, 3, 100.0)
    _fwd_case(1, 4096, 4096, 256, 3, 100.0, peaky=True)
    _fwd_case(1, 4096, 4096, 2304, 3, 100.0)
    _fwd_case(2, 4096, 4096, 256, 154, 100.0)


STAGES = dict(pack=stage_pack, gemm=stage_gemm, fwd_small=stage_fwd_small, fwd_ragged=stage_fwd_ragged,
              fwd_full=stage_fwd_full)

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    for nme in names:
        print("==== stage", nme, flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", nme], timeout=300,
                               capture_output=True, text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print("  [exit %d] stderr tail:\n%s" % (r.returncode, r.stderr[-3000:]))
        except subprocess.TimeoutExpired:
            print("  [TIMEOUT]")
