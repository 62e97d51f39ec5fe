import sys, os, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import factorizer_b200 as ft
dev = torch.device('cuda:0')
C, N = 32, 32
blk = ft.FactorizerBlock(channels=C, spatial_size=(N, N, N), norm=ft.LayerNorm, reshape=(ft.SWMatricize, {"head_dim": 8, "patch_size": 8}),
                         act=torch.nn.ReLU, factorize=ft.NMF, rank=1, num_iters=5, init="uniform", solver="hals", mlp_ratio=2, dropout=0.0).to(dev)
xb = torch.randn(1, C, N, N, N, device=dev, requires_grad=True)
gyb = torch.randn(1, C, N, N, N, device=dev)
stream = torch.cuda.current_stream()
for mode in ("fwd", "fwd+bwd"):
    try:
        cap = torch.cuda.Stream(dev)
        cap.wait_stream(stream)
        with torch.cuda.stream(cap):
            for _ in range(2):
                out = blk(xb)
                if mode != "fwd": out.backward(gyb)
            xb.grad = None
            for p in blk.parameters(): p.grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cap):
                out = blk(xb)
                if mode != "fwd": out.backward(gyb)
        stream.wait_stream(cap)
        g.replay(); torch.cuda.synchronize()
        print(mode, "capture ok")
    except Exception as e:
        traceback.print_exc()
        print(mode, "FAILED", type(e).__name__)
        torch.cuda.synchronize()
