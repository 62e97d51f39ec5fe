"""CPU BASELINE / TEST INFRASTRUCTURE ONLY -- plain-PyTorch restatement of the reference's eager path.

The reference (pashtari/factorizer) is pure PyTorch; it is absent on the GPU box, so the CPU leg of bench.py times this
restatement of the same ATen operator sequence instead (torch.roll + permute-copy + cat, bmm / add / div / clamp_min per
half-step, the window sets summed in shift order, k=1 Conv1d projections, nn.LayerNorm on the channels-last view, exact
GELU), differentiated by torch autograd exactly as the reference is.  Each function cites what it restates
(file:line relative to the reference checkout).  Nothing under ``factorizer_b200/`` imports this module.

Parity status: PINNED -- ``tests/test_oracle.py::test_torch_port_matches_golden`` checks outputs, input gradients and
parameter gradients against the reference-generated golden vectors (``tests/golden/fused.npz``, ``block.npz``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-16  # factorizer/factorization/matrix_factorization.py:200,236


def matricize(x, head_dim, patch, shift):
    """Reshape.forward (operations.py:266-272) with Matricize's pattern (:321-325):
    b (h d) (g0 p0) (g1 p1) (g2 p2) -> (b h) (g0 g1 g2) d (p0 p1 p2), after torch.roll by `shift`."""
    B, C, D, H, W = x.shape
    if any(shift):
        x = torch.roll(x, tuple(shift), (2, 3, 4))
    h = C // head_dim
    p0, p1, p2 = patch
    x = x.reshape(B, h, head_dim, D // p0, p0, H // p1, p1, W // p2, p2)
    x = x.permute(0, 1, 3, 5, 7, 2, 4, 6, 8)
    return x.reshape(B * h, (D // p0) * (H // p1) * (W // p2), head_dim, p0 * p1 * p2)


def unmatricize(y, B, C, size, head_dim, patch, shift):
    """Reshape.inverse_forward (operations.py:274-280)."""
    D, H, W = size
    h = C // head_dim
    p0, p1, p2 = patch
    y = y.reshape(B, h, D // p0, H // p1, W // p2, head_dim, p0, p1, p2)
    y = y.permute(0, 1, 5, 2, 6, 3, 7, 4, 8).reshape(B, C, D, H, W)
    if any(shift):
        y = torch.roll(y, tuple(-s for s in shift), (2, 3, 4))
    return y


def sw_matricize(x, head_dim, patch, shifts):
    """SWMatricize.forward (operations.py:417-421): cat over the window sets."""
    return torch.cat([matricize(x, head_dim, patch, s) for s in shifts], dim=0)


def sw_inverse(y, B, C, size, head_dim, patch, shifts):
    """SWMatricize.inverse_forward (operations.py:423-434): 0.0 + inv_0 + inv_1 + ..., then / S."""
    parts = torch.chunk(y, len(shifts), dim=0)
    out = 0.0
    for part, s in zip(parts, shifts):
        out = out + unmatricize(part, B, C, size, head_dim, patch, s)
    return out / len(shifts)


def hals_rank1(x, u0, v0, num_iters=5):
    """MatrixFactorization.decompose + reconstruct with CoordinateDescent.update_u, R = 1
    (matrix_factorization.py:122-136, 224-227, 514-533): u = relu((x v + eps) / (v.v + eps)), then the same for v."""
    u = u0.expand(*x.shape[:-2], *u0.shape)
    v = v0.expand(*x.shape[:-2], *v0.shape)
    xt = x.mT
    for _ in range(num_iters):
        a, b = x @ v, v.mT @ v
        u = torch.relu((a + EPS) / (b + EPS))
        a, b = xt @ u, u.mT @ u
        v = torch.relu((a + EPS) / (b + EPS))
    return u @ v.mT


def fact_core(z, u0, v0, head_dim=8, patch=(8, 8, 8), shifts=((0, 0, 0), (4, 4, 4)), num_iters=5):
    """reshape -> act -> factorize -> reshape.inverse_forward of FactMixer.forward (factorizer.py:41-50)."""
    B, C = z.shape[:2]
    m = torch.relu(sw_matricize(z, head_dim, patch, shifts))
    return sw_inverse(hals_rank1(m, u0, v0, num_iters), B, C, tuple(z.shape[2:]), head_dim, patch, shifts)


def layernorm_cf(x, weight, bias, eps=1e-5):
    """layers/norm.py:29-34: channels to the back, nn.LayerNorm, channels to the front."""
    return F.layer_norm(x.movedim(1, -1), (x.shape[1],), weight, bias, eps).movedim(-1, 1)


def linear_cf(x, weight, bias=None):
    """layers/linear.py:53-58: flatten the spatial dims, k=1 Conv1d, view back."""
    return F.conv1d(x.flatten(2), weight, bias).view(x.shape[0], -1, *x.shape[2:])


def block_forward(x, sd, head_dim=8, patch=(8, 8, 8), shifts=((0, 0, 0), (4, 4, 4)), num_iters=5):
    """FactorizerBlock.forward (factorizer.py:74-77) with FactMixer.forward (:34-57) and MLP (layers/mlp.py:54-60),
    dropout 0; `sd` = the block's state_dict (tensors, possibly requiring grad)."""
    h = layernorm_cf(x, sd["norm1.norm.weight"], sd["norm1.norm.bias"])
    z = linear_cf(h, sd["fact.in_proj.linear.weight"])
    m = fact_core(z, sd["fact.factorize.init.u0"], sd["fact.factorize.init.v0"], head_dim, patch, shifts, num_iters)
    x = x + linear_cf(m, sd["fact.out_proj.linear.weight"], sd["fact.out_proj.linear.bias"])
    h = layernorm_cf(x, sd["norm2.norm.weight"], sd["norm2.norm.bias"])
    a = F.gelu(linear_cf(h, sd["mlp.block.0.linear.weight"], sd["mlp.block.0.linear.bias"]))
    return x + linear_cf(a, sd["mlp.block.3.linear.weight"], sd["mlp.block.3.linear.bias"])
