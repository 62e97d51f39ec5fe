"""torch.profiler kernel table of one Swin Factorizer training step (config 5) / inference pass (config 4)."""
import sys
sys.path.insert(0, '.')
import torch
from torch import nn
from torch.profiler import profile, ProfilerActivity
import factorizer_b200 as ft

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
n = 128
net = ft.Factorizer(in_channels=4, out_channels=3, spatial_size=(n, n, n), norm=ft.LayerNorm,
                    reshape=(ft.SWMatricize, {"head_dim": 8, "patch_size": 8}), act=nn.ReLU, factorize=ft.NMF, rank=1,
                    num_iters=5, init="uniform", solver="hals", mlp_ratio=2, dropout=0.1).to(dev)
x = torch.rand(1, 4, n, n, n, device=dev)
target = torch.randint(0, 2, (1, 3, n, n, n), device=dev).float()


def step():
    if mode == "infer":
        with torch.no_grad():
            return net(x)
    net.zero_grad(set_to_none=True)
    logits = net(x)
    loss = nn.functional.binary_cross_entropy_with_logits(logits, target)
    loss.backward()
    return loss


if mode == "infer":
    net.eval()
for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=80))
