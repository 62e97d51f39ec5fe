"""Fused matricize+NMF core (fz_swnmf_forward / _backward) over the window geometries the reference uses anywhere
(SURVEY.md section 8a 'parameter envelope'): which kernel path serves each, microseconds, fraction of the HBM roofline
(algorithmic bytes 2 N_el e forward, 3 N_el e backward).  Markdown table on stdout."""
import ctypes, json, os, sys
sys.path.insert(0, '.')
import torch
import factorizer_b200 as ft
from factorizer_b200 import _lib

dev = torch.device('cuda:0')
lib = _lib.lib()
peak = 6549.1
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])

CASES = [
    ("Swin stage 1 / cfg 2 (README.md:43-96)", 1, 32, (128,) * 3, dict(head_dim=8, patch_size=8)),
    ("Swin stage 2", 1, 64, (64,) * 3, dict(head_dim=8, patch_size=8)),
    ("Swin stage 3", 1, 128, (32,) * 3, dict(head_dim=8, patch_size=8)),
    ("Swin stage 4", 1, 256, (16,) * 3, dict(head_dim=8, patch_size=8)),
    ("Swin stage 5 (one window per head, pure wrap)", 1, 512, (8,) * 3, dict(head_dim=8, patch_size=8)),
    ("brats23 bundle: S=4 shifts [0,2,4,6], batch 2", 2, 32, (128,) * 3, dict(head_dim=8, patch_size=8, shifts=[None, 2, 4, 6])),
    ("isles22 bundle: patch 4 (8x64), S=4, batch 8, 64^3", 8, 32, (64,) * 3, dict(head_dim=8, patch_size=4, shifts=[None, 1, 2, 3])),
    ("tests/test_factorizer.py: num_heads=8, patch 4 (4x64)", 3, 32, (64,) * 3, dict(num_heads=8, patch_size=4)),
    ("tests/test_factorizer.py stage 2 (8x64)", 3, 64, (32,) * 3, dict(num_heads=8, patch_size=4)),
    ("tests/test_factorizer.py stage 4 (32x64)", 3, 256, (8,) * 3, dict(num_heads=8, patch_size=4)),
    ("default FactMixer reshape Matricize(num_heads=1, grid_size=1), MU (tests/test_factorizer.py:14-110)", 1, 16, (64,) * 3,
     dict(cls="Matricize", num_heads=1, grid_size=1, solver="mu")),
    ("same, 32 channels, batch 2, HALS", 2, 32, (64,) * 3, dict(cls="Matricize", num_heads=1, grid_size=1)),
]
PATH = {0: "generic smem", 1: "window-at-a-time TMA", 2: "three-pass octant", 3: "sub-warp register", 4: "octant x pairs (rolled)", 5: "grid-wide pass per sweep"}
print("| geometry | x shape | matrix | windows | path | fwd us | bwd us | fwd+bwd % of HBM roofline |")
print("|---|---|---|---|---|---|---|---|")
for name, B, C, size, kw in CASES:
    kw = dict(kw)
    cls, solver = kw.pop("cls", "SWMatricize"), kw.pop("solver", "hals")
    sw = getattr(ft, cls)((None, C, *size), **kw)
    M, N = sw.output_size[2:]
    nmf = ft.NMF((M, N), rank=1, num_iters=5, init='uniform', solver=solver).to(dev)
    g, s = sw._geom.c_geom(B), nmf.solver_spec().c_solver()
    x = torch.rand(B, C, *size, device=dev); gy = torch.randn_like(x)
    y = torch.empty_like(x); gx = torch.empty_like(x)
    nsaved = lib.fz_swnmf_saved_bytes(ctypes.byref(g), ctypes.byref(s))
    saved = torch.empty(max(nsaved, 1), dtype=torch.uint8, device=dev)
    ws = torch.zeros(max(lib.fz_swnmf_workspace_bytes(ctypes.byref(g), ctypes.byref(s)), 1), dtype=torch.uint8, device=dev)
    u0, v0 = nmf.init.u0, nmf.init.v0
    sp = torch.cuda.current_stream(dev).cuda_stream
    sv = saved.data_ptr() if nsaved else None
    def fwd():
        _lib.check(lib.fz_swnmf_forward(x.data_ptr(), u0.data_ptr(), v0.data_ptr(), y.data_ptr(), sv, ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, sp))
    def bwd():
        _lib.check(lib.fz_swnmf_backward(x.data_ptr(), gy.data_ptr(), u0.data_ptr(), v0.data_ptr(), sv, gx.data_ptr(), ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, sp))
    def timeit(fn, reps=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    tf = timeit(fwd); path = lib.fz_last_path(); tb = timeit(bwd)
    nel = x.numel()
    frac = 5 * nel * 4 / ((tf + tb) * 1e-6) / 1e9 / peak
    nwin = sw.output_size[1] * len(sw._geom.shifts) * B * (C // M)
    print(f"| {name} | {tuple(x.shape)} | {M}x{N} | {nwin} | {PATH[path]} | {tf:.1f} | {tb:.1f} | {100 * frac:.1f} |")
    del x, gy, y, gx, saved, ws
