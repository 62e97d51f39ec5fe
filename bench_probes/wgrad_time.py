"""Times fz_linear_wgrad (FP32 pipe against tcgen05) on the weight-gradient shapes of the Swin Factorizer at 128^3."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from factorizer_b200 import _lib as L
dev = torch.device("cuda:0")
lib = L.lib()
st = torch.cuda.current_stream().cuda_stream
shapes = [(32, 108, 128 ** 3), (64, 256, 64 ** 3), (64, 64, 64 ** 3), (128, 64, 64 ** 3), (64, 128, 64 ** 3), (128, 512, 32 ** 3),
          (128, 128, 32 ** 3), (256, 128, 32 ** 3), (256, 1024, 16 ** 3), (256, 256, 16 ** 3), (512, 256, 16 ** 3), (512, 2048, 8 ** 3),
          (512, 512, 8 ** 3), (1024, 512, 8 ** 3), (3, 32, 128 ** 3)]
for cout, cin, vox in shapes:
    dy, x = torch.randn(1, cout, vox, device=dev), torch.randn(1, cin, vox, device=dev)
    res = []
    outs = []
    for mode in (0, 7):
        lib.fz_set_glue_mode(mode)
        dW, db = torch.empty(cout, cin, device=dev), torch.empty(cout, device=dev)
        f = lambda: L.check(lib.fz_linear_wgrad(dy.data_ptr(), x.data_ptr(), dW.data_ptr(), db.data_ptr(), 1, cout, cin, vox, st))
        for _ in range(2):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            f()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 5 * 1e3)
        outs.append((dW.clone(), db.clone()))
    ref = torch.einsum("bov,bcv->oc", dy.double(), x.double())
    sc = float(ref.abs().max())
    errs = [float((o[0].double() - ref).abs().max()) / sc for o in outs]
    print(f"({cout:4d} x {cin:4d}) x {vox:8d} voxels: fp32 pipe {res[0]:7.1f} us  tcgen05 {res[1]:7.1f} us   max err / max |dW|: {errs[0]:.2e} {errs[1]:.2e}", flush=True)
