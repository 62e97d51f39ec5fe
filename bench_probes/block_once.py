"""One FactorizerBlock (config 3) fwd+bwd, a few times: the command ncu wraps for the glue kernels."""
import sys
sys.path.insert(0, '.')
import torch
import factorizer_b200 as ft

dev = torch.device('cuda:0')
C, N = 32, 128
blk = ft.FactorizerBlock(channels=C, spatial_size=(N, N, N), norm=ft.LayerNorm,
                         reshape=(ft.SWMatricize, {"head_dim": 8, "patch_size": 8}), act=torch.nn.ReLU,
                         factorize=ft.NMF, rank=1, num_iters=5, init="uniform", solver="hals", mlp_ratio=2,
                         dropout=0.0).to(dev)
x = torch.rand(1, C, N, N, N, device=dev, requires_grad=True)
gy = torch.randn(1, C, N, N, N, device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(n):
    blk(x).backward(gy)
torch.cuda.synchronize()
