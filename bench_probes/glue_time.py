"""Times every glue kernel of the 32-channel block alone through the C ABI at (1, 32, 128^3) in the given glue modes, and
prints the mlp_backward outputs' error ratios against the FP32-pipe kernel (mode 0) at full size."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from factorizer_b200 import _lib as L

dev = torch.device("cuda:0")
lib = L.lib()
modes = [int(a) for a in sys.argv[1:]] or [0, 5]


def call(fn, *args):
    L.check(fn(*[a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args]))


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


B, C, HID, vox, eps = 1, 32, 64, 128 ** 3, 1e-5
torch.manual_seed(3)
st = torch.cuda.current_stream().cuda_stream
r = lambda *s: torch.randn(*s, device=dev)
x, m, gout = 2 * r(B, C, vox) + 0.5, r(B, C, vox), r(B, C, vox)
g1, b1n, g2, b2n = 1 + 0.3 * r(C), 0.3 * r(C), 1 + 0.3 * r(C), 0.3 * r(C)
w_in, w_out, b_out = r(C, C) / 6, r(C, C) / 6, 0.2 * r(C)
w1, bb1, w2, bb2 = r(HID, C) / 6, 0.2 * r(HID), r(C, HID) / 7, 0.2 * r(C)
z, x1, out, dx1, dm, dx = (torch.empty_like(x) for _ in range(6))
gr = [torch.empty_like(t) for t in (g2, b2n, w1, bb1, w2, bb2)]
dwo, dbo, dwi, dg1, db1n = torch.empty_like(w_out), torch.empty_like(b_out), torch.empty_like(w_in), torch.empty_like(g1), torch.empty_like(b1n)
ref = None
for mode in modes:
    lib.fz_set_glue_mode(mode)
    t = {}
    t["ln_linear_fwd"] = timed(lambda: call(lib.fz_ln_linear_forward, x, g1, b1n, w_in, z, B, C, vox, eps, st))
    t["mixer_mlp_fwd"] = timed(lambda: call(lib.fz_mixer_mlp_forward, x, m, w_out, b_out, g2, b2n, w1, bb1, w2, bb2, x1, out, B, C, HID, vox, eps, st))
    t["mlp_bwd"] = timed(lambda: call(lib.fz_mlp_backward, x1, gout, g2, b2n, w1, bb1, w2, dx1, *gr, B, C, HID, vox, eps, st))
    t["linear_bwd(out)"] = timed(lambda: call(lib.fz_linear_backward, dx1, m, None, None, w_out, None, dm, dwo, dbo, None, None, B, C, vox, 0.0, 0, st))
    t["linear_bwd(in+ln)"] = timed(lambda: call(lib.fz_linear_backward, dm, x, g1, b1n, w_in, dx1, dx, dwi, None, dg1, db1n, B, C, vox, eps, 1, st))
    print(f"mode {mode}: " + "  ".join(f"{k} {v:.1f} us" for k, v in t.items()) + f"  sum {sum(t.values()):.1f} us", flush=True)
    cur = [dx1.clone()] + [g.clone() for g in gr]
    if ref is None:
        ref = cur
    else:
        names = ("dx1", "dgamma2", "dbeta2", "dW1", "db1", "dW2", "db2")
        out_s = []
        for a, b, nme in zip(cur, ref, names):
            sc = 1.0 if nme == "dx1" else max(1.0, float(b.abs().max()))
            out_s.append(f"{nme}={float(((a - b).abs() / sc / (2e-5 + 1e-4 * b.abs() / sc)).max()):.3g}")
        print(f"   mode {mode} vs mode {modes[0]} (err / tol): " + "  ".join(out_s), flush=True)
