"""Where does the FactorizerBlock (config 3) spend its time?  torch.profiler kernel table."""
import sys
sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
import factorizer_b200 as ft

dev = torch.device('cuda:0')
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
C, N = 32, 128
blk = ft.FactorizerBlock(channels=C, spatial_size=(N, N, N), norm=ft.LayerNorm,
                         reshape=(ft.SWMatricize, {"head_dim": 8, "patch_size": 8}), act=torch.nn.ReLU,
                         factorize=ft.NMF, rank=1, num_iters=5, init="uniform", solver="hals", mlp_ratio=2,
                         dropout=0.0).to(dev)
x = torch.rand(1, C, N, N, N, device=dev, requires_grad=True)
gy = torch.randn(1, C, N, N, N, device=dev)
for _ in range(2):
    blk(x).backward(gy)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        blk(x).backward(gy)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
