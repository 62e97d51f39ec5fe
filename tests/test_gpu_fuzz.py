"""Seeded sweep of window geometries through the fused core (C ABI) against the numpy oracle in fp64: every kernel
family (octant, paired sets, window-at-a-time, sub-warp, grid-wide, generic), wrap-around windows on every axis, batches,
truncated unrolls, both solvers.  The oracle is pinned against the reference's goldens (tests/test_oracle.py); this sweep
extends the pinned parity to geometries the goldens do not contain."""
import numpy as np
import pytest
import torch
from torch import nn

from conftest import assert_close
from oracle import factorizer_oracle as O

pytestmark = pytest.mark.gpu

# (x shape, reshape class, reshape kwargs, solver, num_iters, num_grad_steps, relu, expected kernel path)
CASES = [
    # octant kernels: non-cubic volume, batch 2, T = 3
    ((2, 16, 16, 24, 8), "SWMatricize", dict(head_dim=8, patch_size=8), "hals", 3, None, True, 2),
    # paired sets: base shifts (2, 6) only (no unshifted set), and three pairs
    ((1, 8, 16, 16, 16), "SWMatricize", dict(head_dim=8, patch_size=8, shifts=[2, 6]), "hals", 5, None, True, 4),
    ((1, 16, 16, 8, 16), "SWMatricize", dict(head_dim=8, patch_size=8, shifts=[None, 2, 4, 6, (0, 2, 0), (4, 6, 4)]), "hals", 4, 2, True, 4),
    # window-at-a-time: shifts that do not pair up (odd, and three sets)
    ((1, 8, 16, 16, 16), "SWMatricize", dict(head_dim=8, patch_size=8, shifts=[None, 3, 5]), "hals", 5, None, True, 1),
    # sub-warp kernels: every M, wrap on all axes, MU, truncation, no ReLU, 2-D and 1-D inputs
    ((2, 8, 8, 8, 8), "SWMatricize", dict(head_dim=4, patch_size=4, shifts=[None, 1, 3]), "hals", 5, None, True, 3),
    ((1, 32, 8, 4, 12), "SWMatricize", dict(head_dim=16, patch_size=4, shifts=[(1, 2, 3), 2]), "mu", 4, 2, False, 3),
    ((3, 64, 4, 8, 4), "SWMatricize", dict(head_dim=32, patch_size=4), "hals", 2, 1, True, 3),
    ((2, 16, 24, 16), "SWMatricize", dict(head_dim=8, patch_size=8, shifts=[None, (3, 5)]), "mu", 5, None, True, 3),
    ((2, 8, 48), "SWMatricize", dict(head_dim=8, patch_size=16, shifts=[None, 5, 11]), "hals", 5, None, True, 3),
    ((1, 8, 16, 16, 16), "SWMatricize", dict(head_dim=8, patch_size=(4, 8, 8), shifts=[None, 1]), "hals", 5, None, True, 3),
    # grid-wide passes: odd column counts, heads, M = 1, truncation
    ((2, 6, 5, 9, 167), "Matricize", dict(num_heads=2, grid_size=1), "mu", 5, None, True, 5),
    ((1, 24, 7, 11, 13), "Matricize", dict(num_heads=1, grid_size=1), "hals", 3, 2, True, 5),
    ((3, 4, 16500), "Matricize", dict(num_heads=4, grid_size=1), "hals", 5, None, False, 5),
    # generic: rank-1 windows with no specialised kernel (patch 6)
    ((1, 6, 12, 12), "SWMatricize", dict(head_dim=3, patch_size=6), "hals", 5, None, True, 0),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_fused_core_against_oracle(case):
    import factorizer_b200 as ft
    from factorizer_b200 import _lib, _ops
    xs, cls, kw, solver, T, K, relu, path = CASES[case]
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(100 + case)
    n = len(xs) - 2
    reshape = getattr(ft, cls)((None, *xs[1:]), **kw)
    M, N = reshape.output_size[2:]
    nmf = ft.NMF((M, N), rank=1, num_iters=T, num_grad_steps=K, init="uniform", solver=solver).to(dev)
    x_np = rng.standard_normal(xs).astype(np.float32)
    if solver == "mu" or not relu:
        # MU needs non-negative input; HALS on signed data without the ReLU is ill-conditioned (a row of X v can
        # change sign, u collapses to 0 and the next sweep divides by eps: SURVEY App. C), which no fp32
        # implementation reproduces bit-stably -- the reference's FactMixer always applies the ReLU
        x_np = np.abs(x_np) + np.float32(0.05)
    gy_np = rng.standard_normal(xs).astype(np.float32)
    x = torch.from_numpy(x_np).to(dev).requires_grad_(True)
    y = _ops.SWNMF.apply(x, nmf.init.u0, nmf.init.v0, reshape._geom, nmf.solver_spec(), relu)
    assert _lib.lib().fz_last_path() == path
    (gx,) = torch.autograd.grad((y * torch.from_numpy(gy_np).to(dev)).sum(), x)
    # oracle, fp64
    kwo = dict(kw)
    shifts = kwo.pop("shifts", None)
    H, d, grid, patch = O.resolve_geometry((None, *xs[1:]), **kwo)
    if cls == "SWMatricize":
        shifts = O.default_shifts(patch) if shifts is None else shifts
    else:
        shifts = [shifts]
    shifts = O.normalise_shifts(shifts, n)
    u0, v0 = nmf.init.u0.cpu().numpy().astype(np.float64), nmf.init.v0.cpu().numpy().astype(np.float64)
    x64, g64 = x_np.astype(np.float64), gy_np.astype(np.float64)
    y_ref = O.swnmf_forward(x64, u0, v0, H, d, grid, patch, shifts, relu=relu, solver=solver, num_iters=T)
    gx_ref = O.swnmf_backward(x64, g64, u0, v0, H, d, grid, patch, shifts, relu=relu, solver=solver, num_iters=T,
                              num_grad_steps=K)
    assert_close(y.detach().cpu().numpy(), y_ref, what="y")
    assert_close(gx.cpu().numpy(), gx_ref, what="gx")
