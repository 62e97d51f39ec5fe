"""fz_linear_forward (tcgen05, 3xTF32) against the library's batched fp32 GEMM on the channel-map shapes of the Swin Factorizer."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from factorizer_b200 import _lib as L
dev = torch.device("cuda:0")
lib = L.lib()
torch.backends.cuda.matmul.allow_tf32 = False
st = torch.cuda.current_stream().cuda_stream
shapes = [(64, 64, 64 ** 3), (128, 64, 64 ** 3), (64, 128, 64 ** 3), (64, 256, 64 ** 3), (128, 128, 32 ** 3), (256, 128, 32 ** 3),
          (128, 256, 32 ** 3), (256, 256, 16 ** 3), (512, 256, 16 ** 3), (256, 512, 16 ** 3), (512, 512, 8 ** 3), (40, 96, 32 ** 3 + 36),
          (3, 32, 128 ** 3)]


def timed(f, n=5):
    for _ in range(2):
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
    B = 2 if vox < 40000 else 1
    x, W, b = torch.randn(B, cin, vox, device=dev), torch.randn(cout, cin, device=dev) / cin ** 0.5, torch.randn(cout, device=dev)
    y = torch.empty(B, cout, vox, device=dev)
    f_tc = lambda: L.check(lib.fz_linear_forward(x.data_ptr(), W.data_ptr(), b.data_ptr(), y.data_ptr(), B, cin, cout, vox, st))
    f_lib = lambda: torch.baddbmm(b[None, :, None], W.unsqueeze(0).expand(B, -1, -1), x)
    f_tc()
    torch.cuda.synchronize()
    ref = torch.baddbmm(b.double()[None, :, None], W.double().unsqueeze(0).expand(B, -1, -1), x.double())
    err = float(((y.double() - ref).abs() / (1e-5 + 1e-4 * ref.abs())).max())
    err_lib = float(((f_lib().double() - ref).abs() / (1e-5 + 1e-4 * ref.abs())).max())
    print(f"({cout:4d} x {cin:4d}) x {B} x {vox:8d}: library {timed(f_lib):7.1f} us  tcgen05 {timed(f_tc):7.1f} us   err / tol: library {err_lib:.3g}  tcgen05 {err:.3g}", flush=True)
