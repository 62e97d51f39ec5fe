"""TEST INFRASTRUCTURE ONLY -- FactorizerBlock reference at full size: torch fp64 glue around the pinned C oracle core.

Used by tests/test_gpu_fullsize.py and by bench.py's `parity_checked` (the timed buffers are compared with this after the
timing).  The glue (LayerNorm, k=1 projections, GELU MLP, residuals: reference factorizer/factorizer.py:74-77, 34-57) is
plain torch in float64 on whatever device the inputs live on, differentiated by torch autograd; the FactMixer core
(reshape -> ReLU -> NMF -> inverse, factorizer.py:41-50) is oracle/nmf_oracle.c (fp32, CPU), which tests/test_oracle.py
pins on the reference-generated golden vectors.
"""
from __future__ import annotations

import numpy as np
import torch

from . import c_oracle as CO

SHIFTS = [(0, 0, 0), (4, 4, 4)]


class OracleCore(torch.autograd.Function):
    """The FactMixer core as the C oracle computes it, inside a float64 torch graph."""

    @staticmethod
    def forward(ctx, z, v0):
        z32 = z.detach().to(torch.float32).cpu().numpy()
        ctx.z32, ctx.v0 = z32, v0
        return torch.from_numpy(CO.swnmf_forward(z32, v0, 8, (8, 8, 8), SHIFTS)).to(z.device, z.dtype)

    @staticmethod
    def backward(ctx, g):
        g32 = g.to(torch.float32).cpu().numpy()
        return torch.from_numpy(CO.swnmf_backward(ctx.z32, g32, ctx.v0, 8, (8, 8, 8), SHIFTS)).to(g.device, g.dtype), None


def block_reference(state_dict, x, gy, z_product=None, z_tolerance=None):
    """Returns (out, dx, {parameter name: gradient}, z) in float64.

    The ReLU in front of the factorization (factorizer.py:44) is a kink: where the mixer's pre-activation
    z = in_proj(LN(x)) is within fp32 rounding of zero, an fp32 and an fp64 evaluation of z sit on different sides, and
    the gradient of that element then differs by its whole masked term -- at 128^3 x 32 channels a handful of the 67 M
    pre-activations do (measured: input-gradient error 1.8e3 x tolerance at those voxels, 0.05 x everywhere else).  When
    `z_product` (the product's fp32 z) is given, it is first checked against the fp64 z (callback `z_tolerance(got, ref)`),
    and the reference then continues from it (value of the product, derivative of the fp64 graph), so that both sides
    see the same ReLU mask."""
    CO.use_all_cores()
    sd = {k: v.detach().to(torch.float64) for k, v in state_dict.items()}
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.endswith(("u0", "v0"))}
    C = x.shape[1]
    xd = x.detach().to(torch.float64).requires_grad_(True)

    def ln(t, w, b):
        return torch.nn.functional.layer_norm(t.movedim(1, -1), (C,), w, b, 1e-5).movedim(-1, 1)

    def lin(t, w, b=None):
        out = torch.einsum("oi,bi...->bo...", w.squeeze(-1), t)
        return out if b is None else out + b.view(1, -1, *([1] * (t.dim() - 2)))

    v0 = sd["fact.factorize.init.v0"].cpu().numpy().astype(np.float32)
    h = ln(xd, p["norm1.norm.weight"], p["norm1.norm.bias"])
    z = lin(h, p["fact.in_proj.linear.weight"])
    z_ref = z.detach()
    if z_product is not None:
        if z_tolerance is not None:
            z_tolerance(z_product, z_ref)
        z = z + (z_product.to(torch.float64) - z).detach()
    m = OracleCore.apply(z, v0)
    x1 = xd + lin(m, p["fact.out_proj.linear.weight"], p["fact.out_proj.linear.bias"])
    h2 = ln(x1, p["norm2.norm.weight"], p["norm2.norm.bias"])
    a = torch.nn.functional.gelu(lin(h2, p["mlp.block.0.linear.weight"], p["mlp.block.0.linear.bias"]))
    out = x1 + lin(a, p["mlp.block.3.linear.weight"], p["mlp.block.3.linear.bias"])
    names = list(p)
    grads = torch.autograd.grad((out * gy.to(torch.float64)).sum(), [xd] + [p[k] for k in names])
    return out.detach(), grads[0], dict(zip(names, grads[1:])), z_ref
