"""One fz_linear_wgrad launch per shape (for ncu): the (64 x 64) x 64^3 and (128 x 64) x 64^3 weight gradients."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from factorizer_b200 import _lib as L
dev = torch.device("cuda:0")
lib = L.lib()
st = torch.cuda.current_stream().cuda_stream
for cout, cin, vox in [(64, 64, 64 ** 3), (128, 64, 64 ** 3)]:
    dy, x = torch.randn(1, cout, vox, device=dev), torch.randn(1, cin, vox, device=dev)
    dW, db = torch.empty(cout, cin, device=dev), torch.empty(cout, device=dev)
    for _ in range(3):
        L.check(lib.fz_linear_wgrad(dy.data_ptr(), x.data_ptr(), dW.data_ptr(), db.data_ptr(), 1, cout, cin, vox, st))
    torch.cuda.synchronize()
