"""Which ATen operators are left in a Swin Factorizer training step: operator table (CUDA time attributed to the calling op)."""
import sys
sys.path.insert(0, '.')
import torch
from torch import nn
from torch.profiler import profile, ProfilerActivity
import factorizer_b200 as ft

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
    net.zero_grad(set_to_none=True)
    logits = net(x)
    loss = nn.functional.binary_cross_entropy_with_logits(logits, target)
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith("aten::") and e.self_device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:40]:
    print(f"{e.key:32s} calls {e.count:4d}  self cuda {e.self_device_time_total:9.1f} us  shapes {str(e.input_shapes)[:150]}")
prof.export_chrome_trace("gpurun_out/r02y_trace.json")
