"""Time the pointwise-linear weight-gradient kernel (csrc/fz_linear.cu) against the library SGEMM autograd would use."""
import sys
sys.path.insert(0, '.')
import torch
from factorizer_b200 import _lib as L

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda:0')
lib = L.lib()
for cout, cin, n in [(32, 108, 128), (32, 64, 128), (3, 32, 128), (64, 256, 64), (128, 512, 32), (64, 64, 64), (128, 64, 64), (64, 128, 64), (128, 128, 32), (256, 128, 32), (256, 256, 16), (512, 256, 16), (512, 512, 8), (1024, 512, 8)]:
    vox = n ** 3
    gy = torch.randn(1, cout, vox, device=dev)
    x = torch.randn(1, cin, vox, device=dev)
    gw = torch.empty(cout, cin, device=dev)
    gb = torch.empty(cout, device=dev)
    def ours():
        rc = lib.fz_linear_wgrad(L.ptr(gy), L.ptr(x), L.ptr(gw), L.ptr(gb), 1, cout, cin, vox, L.stream_ptr(dev))
        assert rc == 0
    def cublas():
        return torch.bmm(gy, x.transpose(1, 2)), gy.sum((0, 2))
    out = []
    for f in (ours, cublas):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            f()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) * 100)
    ref = cublas()[0][0]
    err = float((gw - ref).abs().max() / ref.abs().max())
    print(f"{cout:5d} x {cin:4d} x {n}^3: kernel {out[0]:8.1f} us   bmm+sum {out[1]:8.1f} us   rel diff {err:.1e}   "
          f"FMA/clk/SM {cout * cin * vox / (out[0] * 1e-6) / 148 / 1.9e9:6.1f}")
