"""Experiment build only (FZ_TUNING=1): clock stamps of CTA 0 of mlp_bwd_tc, per tile."""
import sys, os, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from factorizer_b200 import _lib as L
dev = torch.device("cuda:0")
lib = L.lib()
raw = ctypes.CDLL(lib._name)
B, C, HID, vox, eps = 1, 32, 64, 128 ** 3, 1e-5
torch.manual_seed(3)
st = torch.cuda.current_stream().cuda_stream
r = lambda *s: torch.randn(*s, device=dev)
x1, gout = r(B, C, vox), r(B, C, vox)
g2, b2n = 1 + 0.3 * r(C), 0.3 * r(C)
w1, bb1, w2 = r(HID, C) / 6, 0.2 * r(HID), r(C, HID) / 7
dx1 = torch.empty_like(x1)
gr = [torch.empty_like(t) for t in (g2, b2n, w1, bb1, w2, torch.empty(C, device=dev))]
buf = torch.zeros(16 * 16, dtype=torch.int64, device=dev)
raw.fz_debug_mlp_bwd_trace.argtypes = [ctypes.c_void_p]
assert raw.fz_debug_mlp_bwd_trace(buf.data_ptr()) == 0
args = [a.data_ptr() for a in (x1, gout, g2, b2n, w1, bb1, w2, dx1, *gr)]
for _ in range(3):
    L.check(lib.fz_mlp_backward(*args, B, C, HID, vox, eps, st))
torch.cuda.synchronize()
t = buf.cpu().view(16, 16)
names = ["P1 start", "wg2(prev) seen", "wg1(prev) seen", "P1 staged", "g1 seen", "P2 done", "g3 seen", "P3 done",
         "MMA0: p1a seen", "MMA0: g1 issued", "MMA0: p2 seen", "MMA0: g3a issued", "MMA2: wg1 issued", "fetch issued", "MMA1: p1b seen", "MMA1: g2 issued"]
for it in range(2, 8):
    t0 = int(t[it, 0])
    print(f"tile {it} (+{t0 - int(t[it - 1, 0])}): " + "  ".join(f"{names[k]} {int(t[it, k]) - t0}" for k in range(1, 16)))
