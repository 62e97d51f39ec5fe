"""The Swin Factorizer training step as ONE CUDA graph (bench.py's `model` leg): replay time with the tcgen05 channel map and
with the library GEMM in its place, and the kernel table of two replays (torch.profiler sees the graph's kernels)."""
import sys
sys.path.insert(0, '.')
import torch
from torch import nn
from torch.profiler import profile, ProfilerActivity
import factorizer_b200 as ft
from factorizer_b200 import _ops

dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
n = 128
x = torch.rand(1, 4, n, n, n, device=dev)
target = torch.randint(0, 2, (1, 3, n, n, n), device=dev).float()


def build():
    torch.manual_seed(1234)
    net = ft.Factorizer(in_channels=4, out_channels=3, spatial_size=(n, n, n), norm=ft.LayerNorm,
                        reshape=(ft.SWMatricize, {"head_dim": 8, "patch_size": 8}), act=nn.ReLU, factorize=ft.NMF, rank=1,
                        num_iters=5, init="uniform", solver="hals", mlp_ratio=2, dropout=0.1).to(dev)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-5, capturable=True, fused=True)

    def body():
        opt.zero_grad(set_to_none=True)
        logits = net(x)
        p = torch.sigmoid(logits)
        dice = 1 - (2 * (p * target).sum((2, 3, 4)) + 1e-5) / (p.sum((2, 3, 4)) + target.sum((2, 3, 4)) + 1e-5)
        loss = nn.functional.binary_cross_entropy_with_logits(logits, target) + dice.mean()
        loss.backward()
        opt.step()

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            body()
        opt.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            body()
    torch.cuda.current_stream(dev).wait_stream(side)
    return g, net, opt


def timed(g, reps=5):
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


g, net, opt = build()
print(f"graph step, tcgen05 channel maps: {timed(g):.3f} ms", flush=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=60, max_name_column_width=90))
del g, net, opt
if len(sys.argv) > 1 and sys.argv[1] == "ab":
    real = _ops.linear_forward

    def lib_forward(x, weight, bias):
        wb = weight.unsqueeze(0).expand(x.shape[0], -1, -1)
        return torch.bmm(wb, x) if bias is None else torch.baddbmm(bias[None, :, None], wb, x)

    _ops.linear_forward = lib_forward
    g, net, opt = build()
    print(f"graph step, library GEMM channel maps: {timed(g):.3f} ms", flush=True)
