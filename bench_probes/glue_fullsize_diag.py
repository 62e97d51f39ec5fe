"""Diagnostic: every glue kernel of the 32-channel block at a given voxel count against fp64 torch autograd; prints
max err / tol per output (rtol 1e-4, atol 2e-5) instead of asserting."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from factorizer_b200 import _lib as L

dev = torch.device("cuda:0")
lib = L.lib()


def call(fn, *args):
    L.check(fn(*[a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args]))


def ratio(a, b, scale=1.0):
    a = a.detach().double() / scale; b = b.detach().double() / scale
    return float(((a - b).abs() / (2e-5 + 1e-4 * b.abs())).max())


def run(B, vox, HID=64, mode=None):
    if mode is not None:
        lib.fz_set_glue_mode(mode)
    torch.manual_seed(3)
    C = 32
    st = torch.cuda.current_stream().cuda_stream
    r = lambda *s: torch.randn(*s, device=dev)
    x, m, gout = 2 * r(B, C, vox) + 0.5, r(B, C, vox), r(B, C, vox)
    g1, b1n, g2, b2n = 1 + 0.3 * r(C), 0.3 * r(C), 1 + 0.3 * r(C), 0.3 * r(C)
    w_in, w_out, b_out = r(C, C) / 6, r(C, C) / 6, 0.2 * r(C)
    w1, bb1, w2, bb2 = r(HID, C) / 6, 0.2 * r(HID), r(C, HID) / 7, 0.2 * r(C)
    eps = 1e-5
    D = lambda t: t.detach().double().requires_grad_(True)
    F = torch.nn.functional
    ln = lambda t, g, b: F.layer_norm(t.movedim(1, -1), (C,), g, b, eps).movedim(-1, 1)
    lin = lambda t, w, b=None: torch.einsum("oc,bcv->bov", w, t) + (0 if b is None else b[None, :, None])
    res = {}
    z = torch.empty_like(x)
    call(lib.fz_ln_linear_forward, x, g1, b1n, w_in, z, B, C, vox, eps, st)
    xd, g1d, b1d, wind = D(x), D(g1), D(b1n), D(w_in)
    zd = lin(ln(xd, g1d, b1d), wind)
    res["z"] = ratio(z, zd)
    dz, resid = r(B, C, vox), r(B, C, vox)
    ref = torch.autograd.grad((zd * dz.double()).sum() + (xd * resid.double()).sum(), [xd, wind, g1d, b1d])
    dx, dw, dg, dbt = torch.empty_like(x), torch.empty_like(w_in), torch.empty_like(g1), torch.empty_like(b1n)
    call(lib.fz_linear_backward, dz, x, g1, b1n, w_in, resid, dx, dw, None, dg, dbt, B, C, vox, eps, 1, st)
    res["dx(in_proj+ln1)"] = ratio(dx, ref[0])
    for got, want, what in ((dw, ref[1], "dW_in"), (dg, ref[2], "dgamma1"), (dbt, ref[3], "dbeta1")):
        res[what] = ratio(got, want, max(1.0, float(want.abs().max())))
    del zd, xd, ref
    md, woutd, boutd = D(m), D(w_out), D(b_out)
    yd = lin(md, woutd, boutd)
    ref = torch.autograd.grad((yd * dz.double()).sum(), [md, woutd, boutd])
    dm, dwo, dbo = torch.empty_like(x), torch.empty_like(w_out), torch.empty_like(b_out)
    call(lib.fz_linear_backward, dz, m, None, None, w_out, None, dm, dwo, dbo, None, None, B, C, vox, 0.0, 0, st)
    res["dm"] = ratio(dm, ref[0])
    for got, want, what in ((dwo, ref[1], "dW_out"), (dbo, ref[2], "db_out")):
        res[what] = ratio(got, want, max(1.0, float(want.abs().max())))
    del yd, md, ref
    x1, out = torch.empty_like(x), torch.empty_like(x)
    call(lib.fz_mixer_mlp_forward, x, m, w_out, b_out, g2, b2n, w1, bb1, w2, bb2, x1, out, B, C, HID, vox, eps, st)
    x1d = (x.double() + lin(m.double(), w_out.double(), b_out.double())).detach().requires_grad_(True)
    g2d, b2d, w1d, bb1d, w2d, bb2d = D(g2), D(b2n), D(w1), D(bb1), D(w2), D(bb2)
    outd = x1d + lin(F.gelu(lin(ln(x1d, g2d, b2d), w1d, bb1d)), w2d, bb2d)
    res["x1"] = ratio(x1, x1d); res["out"] = ratio(out, outd)
    ref = torch.autograd.grad((outd * gout.double()).sum(), [x1d, g2d, b2d, w1d, bb1d, w2d, bb2d])
    dx1 = torch.empty_like(x)
    got = [dx1] + [torch.empty_like(t) for t in (g2, b2n, w1, bb1, w2, bb2)]
    call(lib.fz_mlp_backward, x1, gout, g2, b2n, w1, bb1, w2, *got, B, C, HID, vox, eps, st)
    res["dx1"] = ratio(got[0], ref[0])
    for a, b, what in zip(got[1:], ref[1:], ("dgamma2", "dbeta2", "dW1", "db1", "dW2", "db2")):
        res[what] = ratio(a, b, max(1.0, float(b.abs().max())))
    torch.cuda.synchronize()
    print(f"B={B} vox={vox} HID={HID} mode={lib.fz_get_glue_mode()}: " + "  ".join(f"{k}={v:.3g}" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    for vox in (64 ** 3, 96 ** 3, 128 ** 3):
        run(1, vox)
    run(2, 64 ** 3)
    run(1, 128 ** 3, mode=0)
