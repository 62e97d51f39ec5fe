"""fz_linear_forward_ex: time of each fused epilogue (none / + residual / GELU with both outputs / * gelu'(aux)) on channel-map
shapes of the Swin Factorizer, against the plain kernel + the ATen elementwise pass it replaces."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from factorizer_b200 import _lib as L
dev = torch.device("cuda:0")
lib = L.lib()
st = torch.cuda.current_stream().cuda_stream
shapes = [(64, 64, 64 ** 3), (128, 64, 64 ** 3), (64, 128, 64 ** 3), (128, 128, 32 ** 3), (256, 128, 32 ** 3), (256, 256, 16 ** 3),
          (512, 256, 16 ** 3), (512, 512, 8 ** 3), (1024, 512, 8 ** 3)]


def timed(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for cout, cin, vox in shapes:
    B = 1
    x, W, b = torch.randn(B, cin, vox, device=dev), torch.randn(cout, cin, device=dev) / cin ** 0.5, torch.randn(cout, device=dev)
    y, y2, aux = (torch.empty(B, cout, vox, device=dev) for _ in range(3))
    aux.normal_()
    run = lambda epi: L.check(lib.fz_linear_forward_ex(x.data_ptr(), W.data_ptr(), b.data_ptr(), y.data_ptr(), B, cin, cout, vox, epi, 0,
                                                       aux.data_ptr(), y2.data_ptr(), st))
    t = [timed(lambda e=e: run(e)) for e in range(4)]
    t_add = timed(lambda: torch.add(y, aux, out=y2))
    t_gelu = timed(lambda: torch.nn.functional.gelu(y))
    print(f"({cout:4d} x {cin:4d}) x {vox:7d}: none {t[0]:6.1f}  +res {t[1]:6.1f}  gelu {t[2]:6.1f}  gelu' {t[3]:6.1f} us | aten add {t_add:5.1f}  gelu {t_gelu:5.1f} us",
          flush=True)
