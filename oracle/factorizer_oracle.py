"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's (pashtari/factorizer) context-modelling hot path:
shifted-window matricize, the differentiable NMF layer ('mu' and 'hals' solvers) and the
FactMixer core that chains them.  Nothing under ``factorizer_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs use it, and only as
the checker / CPU baseline, never as the product path.

Parity status: PINNED.  The reference ships no golden vectors of its own (its tests only assert
shapes / finiteness / non-negativity), so the oracle is pinned against outputs of the reference
itself: ``tests/golden/make_golden.py`` imports ``/root/reference`` in the build container, runs
the reference modules (forward and torch-autograd backward) on seeded inputs and commits the results
under ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks every function here against them.

All ``file:line`` citations are relative to the reference checkout (``/root/reference``).

Conventions: ``x`` volumes are ``(B, C, *spatial)`` C-contiguous; matricised tensors are
``(S*B*H, G, d, P)``; NMF operands are ``(..., M, N)``, ``u (..., M, R)``, ``v (..., N, R)``.
The arithmetic dtype follows the input (float32 to mirror the reference, float64 to measure
conditioning).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

EPS = 1e-16  # factorizer/factorization/matrix_factorization.py:200,236


# ----------------------------------------------------------------------------------------------
# geometry helpers  (factorizer/factorization/operations.py:299-355, 381-415)
# ----------------------------------------------------------------------------------------------
def _ntuple(v, n):
    if isinstance(v, (tuple, list)):
        assert len(v) == n
        return tuple(v)
    return (v,) * n


def resolve_geometry(input_size, num_heads=None, head_dim=None, grid_size=None, patch_size=None):
    """Infer (H, d, grid, patch) the way Reshape.infer_dims does (operations.py:196-235)."""
    C = input_size[1]
    spatial = tuple(input_size[2:])
    n = len(spatial)
    assert (num_heads, head_dim) != (None, None)
    if num_heads is not None and head_dim is not None:
        H, d = max(num_heads, 1), max(head_dim, 1)
    elif num_heads is not None:
        H = max(num_heads, 1)
        d = C // H
    else:
        d = max(head_dim, 1)
        H = C // d
    grid_size = _ntuple(grid_size, n)
    patch_size = _ntuple(patch_size, n)
    grid, patch = [], []
    for s, g, p in zip(spatial, grid_size, patch_size):
        assert (g, p) != (None, None)
        if g is not None and p is not None:
            g, p = max(g, 1), max(p, 1)
        elif g is not None:
            g = max(g, 1)
            p = s // g
        else:
            p = max(p, 1)
            g = s // p
        grid.append(g)
        patch.append(p)
    return H, d, tuple(grid), tuple(patch)


def default_shifts(patch):
    """SWMatricize default: [None, patch//2] (operations.py:397-398)."""
    return [None, tuple(p // 2 for p in patch)]


def normalise_shifts(shifts, n):
    """None -> zeros; int -> n-tuple (operations.py:342-346)."""
    out = []
    for s in shifts:
        if s is None:
            out.append((0,) * n)
        else:
            out.append(tuple(int(v) for v in _ntuple(s, n)))
    return out


# ----------------------------------------------------------------------------------------------
# Matricize / SWMatricize  (operations.py:266-280, 321-325, 417-434)
# ----------------------------------------------------------------------------------------------
def matricize(x: np.ndarray, H: int, d: int, grid, patch, shift=None) -> np.ndarray:
    """Reshape.forward: torch.roll (operations.py:268-269) then the einops pattern
    ``b (h d) (g0 p0).. -> (b h) (g0..) d (p0..)`` (operations.py:321-325)."""
    B = x.shape[0]
    n = x.ndim - 2
    if shift is not None and any(shift):
        x = np.roll(x, shift, axis=tuple(range(2, 2 + n)))
    shape = [B, H, d]
    for g, p in zip(grid, patch):
        shape += [g, p]
    x = x.reshape(shape)
    g_axes = [3 + 2 * i for i in range(n)]
    p_axes = [4 + 2 * i for i in range(n)]
    x = x.transpose([0, 1] + g_axes + [2] + p_axes)
    return np.ascontiguousarray(x).reshape(B * H, math.prod(grid), d, math.prod(patch))


def unmatricize(y: np.ndarray, B: int, H: int, d: int, grid, patch, shift=None) -> np.ndarray:
    """Reshape.inverse_forward: inverse pattern then roll by -shift (operations.py:274-280)."""
    n = len(grid)
    y = y.reshape([B, H] + list(grid) + [d] + list(patch))
    # axes now: b h g0..g{n-1} d p0..p{n-1}  ->  b h d g0 p0 g1 p1 ...
    perm = [0, 1, 2 + n]
    for i in range(n):
        perm += [2 + i, 3 + n + i]
    y = y.transpose(perm)
    spatial = [g * p for g, p in zip(grid, patch)]
    y = np.ascontiguousarray(y).reshape([B, H * d] + spatial)
    if shift is not None and any(shift):
        y = np.roll(y, tuple(-s for s in shift), axis=tuple(range(2, 2 + n)))
    return y


def swmat_forward(x, H, d, grid, patch, shifts) -> np.ndarray:
    """SWMatricize.forward: concatenate the shifted window sets on dim 0 (operations.py:417-421)."""
    return np.concatenate([matricize(x, H, d, grid, patch, s) for s in shifts], axis=0)


def swmat_inverse(y, B, H, d, grid, patch, shifts) -> np.ndarray:
    """SWMatricize.inverse_forward: ``out = 0.0 + inv_0 + inv_1 ...; out / S`` in that order
    (operations.py:423-434)."""
    S = len(shifts)
    n = y.shape[0] // S
    out = 0.0
    for j, s in enumerate(shifts):
        out = out + unmatricize(y[j * n:(j + 1) * n], B, H, d, grid, patch, s)
    return (out / y.dtype.type(S)).astype(y.dtype)


# ----------------------------------------------------------------------------------------------
# NMF half-steps  (matrix_factorization.py:122-136, 210-229, 241-247)
# ----------------------------------------------------------------------------------------------
def _mT(a):
    return np.swapaxes(a, -1, -2)


def _half_fwd(x, u, v, kind: str, eps):
    """One ``update_u(x, u, v)``; returns (u_new, cache).  The v update is the same function on
    ``x.mT`` with the roles swapped (matrix_factorization.py:122-124)."""
    a = x @ v                       # (..., M, R)
    b = _mT(v) @ v                  # (..., R, R)
    if kind == "mu":                # matrix_factorization.py:241-247
        num = u * a + eps
        den = u @ b + eps
        u_new = num / den
        return u_new, (x, u, v, a, b, den, u_new)
    if kind == "hals":              # matrix_factorization.py:210-229 with project = ReLU (:609)
        R = u.shape[-1]
        if R == 1:
            pre = (a + eps) / (b + eps)
            u_new = np.maximum(pre, 0)
            return u_new, (x, u, v, a, b, pre, u_new)
        u_new = u.copy()
        pre = np.empty_like(u)
        for r in range(R):
            idx = [j for j in range(R) if j != r]
            num = a[..., r:r + 1] - u_new[..., idx] @ b[..., idx, r:r + 1] + eps
            den = b[..., r:r + 1, r:r + 1] + eps
            pre[..., r:r + 1] = num / den
            u_new[..., r:r + 1] = np.maximum(pre[..., r:r + 1], 0)
        return u_new, (x, u, v, a, b, pre, u_new)
    raise ValueError(kind)


def _half_bwd(cache, g_new, kind: str, eps):
    """Adjoint of ``_half_fwd``: given dL/du_new returns (dL/dx, dL/du, dL/dv)."""
    x, u, v, a, b, aux, u_new = cache
    R = u.shape[-1]
    if kind == "mu":
        den = aux
        gn = g_new / den                       # d/d num
        gd = -g_new * u_new / den              # d/d den
        gu = gn * a + gd @ _mT(b)
        ga = gn * u
        gb = _mT(u) @ gd
    else:
        pre = aux
        ga = np.zeros_like(a)
        gb = np.zeros_like(b)
        gu = np.zeros_like(u)
        gun = g_new.copy()                     # grads wrt the *new* columns, updated in place
        for r in reversed(range(R)):
            g = gun[..., r:r + 1] * (pre[..., r:r + 1] > 0)
            den = b[..., r:r + 1, r:r + 1] + eps
            acc = g / den                      # d/d numerator
            gb[..., r:r + 1, r:r + 1] -= (g * pre[..., r:r + 1]).sum(-2, keepdims=True) / den
            ga[..., r:r + 1] += acc
            for j in range(R):
                if j == r:
                    continue
                # column j seen by step r: new if j < r, old if j > r
                uj = u_new[..., j:j + 1] if j < r else u[..., j:j + 1]
                gb[..., j:j + 1, r:r + 1] -= (acc * uj).sum(-2, keepdims=True)
                contrib = -acc * b[..., j:j + 1, r:r + 1]
                if j < r:
                    gun[..., j:j + 1] += contrib
                else:
                    gu[..., j:j + 1] += contrib
        # for R == 1 the old u does not feed u_new at all (matrix_factorization.py:224-227)
    gx = ga @ _mT(v)
    gv = _mT(x) @ ga + v @ (gb + _mT(gb))
    return gx, gu, gv


def nmf_decompose(x, u0, v0, solver="hals", num_iters=5, eps=EPS, keep=False):
    """MatrixFactorization.decompose (matrix_factorization.py:514-530) with RandomInit
    (:52-58: the same u0/v0 broadcast to every matrix) and BCDSolver.forward (:126-136)."""
    dt = x.dtype
    eps = dt.type(eps)
    batch = x.shape[:-2]
    u = np.broadcast_to(u0.astype(dt), batch + u0.shape).copy()
    v = np.broadcast_to(v0.astype(dt), batch + v0.shape).copy()
    caches = []
    for _ in range(num_iters):
        u, cu = _half_fwd(x, u, v, solver, eps)
        v, cv = _half_fwd(_mT(x), v, u, solver, eps)
        if keep:
            caches.append((cu, cv))
    return (u, v, caches) if keep else (u, v)


def nmf_forward(x, u0, v0, solver="hals", num_iters=5, eps=EPS):
    """MatrixFactorization.forward: reconstruct(decompose(x)) = u @ v.mT
    (matrix_factorization.py:532-533, 544-546)."""
    u, v = nmf_decompose(x, u0, v0, solver, num_iters, eps)
    return u @ _mT(v)


def nmf_backward(x, u0, v0, gy=None, gu=None, gv=None, solver="hals", num_iters=5,
                 num_grad_steps=None, eps=EPS):
    """dL/dx of the unrolled solver.  ``gy`` is dL/d(u v^T); ``gu``/``gv`` are optional direct
    gradients on the factors (decompose() users).  ``num_grad_steps=k`` differentiates only the
    last k iterations (MatrixFactorization.context, matrix_factorization.py:506-512)."""
    dt = x.dtype
    eps_t = dt.type(eps)
    k = num_iters if num_grad_steps is None else min(num_grad_steps, num_iters)
    u, v, caches = nmf_decompose(x, u0, v0, solver, num_iters, eps, keep=True)
    gu_t = np.zeros_like(u) if gu is None else gu.astype(dt).copy()
    gv_t = np.zeros_like(v) if gv is None else gv.astype(dt).copy()
    if gy is not None:
        gu_t = gu_t + gy @ v
        gv_t = gv_t + _mT(gy) @ u
    gx = np.zeros_like(x)
    for t in reversed(range(num_iters - k, num_iters)):
        cu, cv = caches[t]
        # v_t = update_u(x.mT, v_{t-1}, u_t)
        gxT, gv_prev, gu_from_v = _half_bwd(cv, gv_t, solver, eps_t)
        gx += _mT(gxT)
        gu_t = gu_t + gu_from_v
        # u_t = update_u(x, u_{t-1}, v_{t-1})
        gx_u, gu_prev, gv_from_u = _half_bwd(cu, gu_t, solver, eps_t)
        gx += gx_u
        gu_t = gu_prev
        gv_t = gv_prev + gv_from_u
    return gx


# ----------------------------------------------------------------------------------------------
# FactMixer core: reshape -> act -> factorize -> inverse  (factorizer/factorizer.py:41-50)
# ----------------------------------------------------------------------------------------------
def swnmf_forward(x, u0, v0, H, d, grid, patch, shifts, relu=True, solver="hals", num_iters=5,
                  eps=EPS):
    B = x.shape[0]
    m = swmat_forward(x, H, d, grid, patch, shifts)
    if relu:
        m = np.maximum(m, 0)                                   # factorizer.py:44
    y = nmf_forward(m, u0, v0, solver, num_iters, eps)         # factorizer.py:47
    return swmat_inverse(y, B, H, d, grid, patch, shifts)      # factorizer.py:50


def swnmf_backward(x, gy_vol, u0, v0, H, d, grid, patch, shifts, relu=True, solver="hals",
                   num_iters=5, num_grad_steps=None, eps=EPS):
    """Adjoint of swnmf_forward w.r.t. x: gather dY/S into every shift's windows, NMF backward,
    ReLU mask, scatter-add over the shifts (no 1/S)."""
    B = x.shape[0]
    S = len(shifts)
    m = swmat_forward(x, H, d, grid, patch, shifts)
    mask = (m > 0) if relu else None
    if relu:
        m = np.maximum(m, 0)
    g = swmat_forward(gy_vol, H, d, grid, patch, shifts) / x.dtype.type(S)
    gm = nmf_backward(m, u0, v0, gy=g, solver=solver, num_iters=num_iters,
                      num_grad_steps=num_grad_steps, eps=eps)
    if relu:
        gm = gm * mask
    n = gm.shape[0] // S
    out = np.zeros_like(x)
    for j, s in enumerate(shifts):
        out = out + unmatricize(gm[j * n:(j + 1) * n], B, H, d, grid, patch, s)
    return out


# ----------------------------------------------------------------------------------------------
# Gram-matrix form of rank-1 HALS on non-negative windows (what the CUDA kernels evaluate when
# act = ReLU).  Algebraically identical to nmf_forward / nmf_backward above whenever no ReLU of the
# solver clips after the first half-step, which X >= 0 guarantees
# (factorizer/factorization/matrix_factorization.py:224-227: a = x @ v >= 0).  Kept here so the
# conditioning of the reformulation is pinned against the reference's golden vectors on CPU.
# ----------------------------------------------------------------------------------------------
def hals_r1_gram(x, v0, gy, num_iters=5, num_grad_steps=None, eps=EPS):
    """x (n, M, N) >= 0, v0 (N,), gy (n, M, N) -> (y, dx), every op in x.dtype.

    Forward: Gam = X X^T, r = X 1, a_1 = X v0; u_{t+1} = relu((rd_t (Gam u_t + eps r) + eps) /
    (rd_t^2 (u_t Gam u_t + 2 eps u_t.r + N eps^2) + eps)), v_T = rd_T (X^T u_T + eps), y = u_T v_T^T.
    Backward: SURVEY App. A.3 with vbar_t = X^T z_t + kappa_t 1 (t < T):
    dX = rd_T u_T gv^T + M X + m 1^T + abar_L v_{L-1}^T."""
    dt = x.dtype.type
    x = np.ascontiguousarray(x)
    n, M, N = x.shape
    T = num_iters
    K = T if num_grad_steps is None else max(1, min(T, num_grad_steps))
    eps = dt(eps)
    v0 = v0.astype(x.dtype)
    Gm = np.einsum("nik,njk->nij", x, x).astype(x.dtype)
    r = x.sum(-1).astype(x.dtype)
    a = np.einsum("nik,k->ni", x, v0).astype(x.dtype)
    b = np.full(n, (v0 * v0).sum(), dtype=x.dtype)
    us, bs, rds = [], [], []
    for t in range(T):
        u = np.maximum((a + eps) / (b + eps)[:, None], 0).astype(x.dtype)
        rd = (dt(1) / ((u * u).sum(-1) + eps)).astype(x.dtype)
        us.append(u); bs.append(b.copy()); rds.append(rd)
        if t < T - 1:
            Gu = np.einsum("nij,nj->ni", Gm, u).astype(x.dtype)
            a = ((Gu + eps * r) * rd[:, None]).astype(x.dtype)
            b = (((u * Gu).sum(-1) + 2 * eps * (u * r).sum(-1) + dt(N) * eps * eps) * rd * rd).astype(x.dtype)
    uT, rdT = us[-1], rds[-1]
    vT = np.maximum((np.einsum("nik,ni->nk", x, uT) + eps) * rdT[:, None], 0).astype(x.dtype)
    y = (uT[:, :, None] * vT[:, None, :]).astype(x.dtype)
    if gy is None:
        return y, None
    gy = gy.astype(x.dtype)
    gu = np.einsum("nik,nk->ni", gy, vT)
    gv = np.einsum("nik,ni->nk", gy, uT).astype(x.dtype)
    w = (gu + np.einsum("nik,nk->ni", x, gv * rdT[:, None])).astype(x.dtype)
    e = (gv * vT).sum(-1).astype(x.dtype)
    Mm = np.zeros((n, M, M), x.dtype)
    m = np.zeros((n, M), x.dtype)

    def close_step(w_, db_, u_, b_, first):
        ub = w_ + 2 * db_[:, None] * u_
        if first:
            ub = np.where(u_ > 0, ub, 0)       # only u_1 can be clipped (v0 may be anything)
        rb = dt(1) / (b_ + eps)
        return (ub * rb[:, None]).astype(x.dtype), (-(ub * u_).sum(-1) * rb).astype(x.dtype)

    ab, bbar = close_step(w, -e * rdT, uT, bs[-1], T == 1)
    for t in range(T - 2, T - K - 1, -1):
        u, rd = us[t], rds[t]
        brd = 2 * bbar * rd
        kappa = brd * eps
        z = (ab + brd[:, None] * u).astype(x.dtype)
        Mm += rd[:, None, None] * (u[:, :, None] * z[:, None, :] + ab[:, :, None] * u[:, None, :])
        m += rd[:, None] * (kappa[:, None] * u + eps * ab)
        Gz = np.einsum("nij,nj->ni", Gm, z)
        Gu = np.einsum("nij,nj->ni", Gm, u)
        w = (rd[:, None] * (Gz + kappa[:, None] * r)).astype(x.dtype)
        qv = rd * ((z * Gu).sum(-1) + eps * (z * r).sum(-1) + kappa * (u * r).sum(-1) + dt(N) * kappa * eps)
        ab, bbar = close_step(w, (-qv * rd).astype(x.dtype), u, bs[t], t == 0)
    dx = rdT[:, None, None] * uT[:, :, None] * gv[:, None, :]
    if K >= T:
        dx = dx + ab[:, :, None] * v0[None, None, :]
    else:   # truncated unroll: v_{L-1} = rd (X^T u_{L-1} + eps) is a constant, a_L = X v_{L-1} still reads X
        u, rd = us[T - K - 1], rds[T - K - 1]
        Mm += rd[:, None, None] * ab[:, :, None] * u[:, None, :]
        m += rd[:, None] * eps * ab
    dx = dx + np.einsum("nij,njk->nik", Mm, x) + m[:, :, None]
    return y, dx.astype(x.dtype)


# ----------------------------------------------------------------------------------------------
# FactorizerBlock glue around the core (factorizer/factorizer.py:34-57, 74-77): channels-first
# LayerNorm (layers/norm.py:29-34), pointwise Linear = Conv1d k=1 (layers/linear.py:53-58),
# MLP = Linear -> exact GELU -> Linear (layers/mlp.py:54-60).  Forward and hand-derived backward,
# pinned against the reference's autograd by tests/test_oracle.py (golden block.npz).
# ----------------------------------------------------------------------------------------------
def _erf(a):
    return np.vectorize(math.erf, otypes=[np.float64])(a).astype(a.dtype)


def layernorm_cf(x, gamma, beta, eps=1e-5):
    """LayerNorm over axis 1 of (B, C, V): biased variance, eps inside the root.  Returns (y, xhat, rstd)."""
    mean = x.mean(axis=1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + x.dtype.type(eps))
    xhat = (x - mean) * rstd
    return xhat * gamma[None, :, None] + beta[None, :, None], xhat, rstd


def layernorm_cf_backward(dy, xhat, rstd, gamma):
    t = dy * gamma[None, :, None]
    dx = rstd * (t - t.mean(axis=1, keepdims=True) - xhat * (t * xhat).mean(axis=1, keepdims=True))
    return dx, (dy * xhat).sum(axis=(0, 2)), dy.sum(axis=(0, 2))


def linear_cf(x, W, b=None):
    y = np.einsum("oc,bcv->bov", W, x)
    return y if b is None else y + b[None, :, None]


def gelu(h):
    return 0.5 * h * (1.0 + _erf(h * h.dtype.type(1.0 / math.sqrt(2.0))))


def gelu_grad(h):
    cdf = 0.5 * (1.0 + _erf(h * h.dtype.type(1.0 / math.sqrt(2.0))))
    return cdf + h * np.exp(-0.5 * h * h) * h.dtype.type(1.0 / math.sqrt(2.0 * math.pi))


def _block_params(sd, dtype):
    g = lambda k: np.asarray(sd[k], dtype=dtype)
    sq = lambda k: g(k)[:, :, 0]
    return dict(g1=g("norm1.norm.weight"), b1=g("norm1.norm.bias"), win=sq("fact.in_proj.linear.weight"),
                wout=sq("fact.out_proj.linear.weight"), bout=g("fact.out_proj.linear.bias"),
                g2=g("norm2.norm.weight"), b2=g("norm2.norm.bias"), w1=sq("mlp.block.0.linear.weight"),
                bb1=g("mlp.block.0.linear.bias"), w2=sq("mlp.block.3.linear.weight"), bb2=g("mlp.block.3.linear.bias"),
                u0=g("fact.factorize.init.u0"), v0=g("fact.factorize.init.v0"))


def block_forward(x, sd, H, d, grid, patch, shifts, num_iters=5, eps_ln=1e-5, keep=False):
    """FactorizerBlock.forward with norm = LayerNorm, act = ReLU, HALS NMF, no dropout.
    ``sd``: the block's state_dict as numpy arrays (reference parameter names)."""
    p = _block_params(sd, x.dtype)
    shape = x.shape
    xf = x.reshape(shape[0], shape[1], -1)
    n1, xh1, r1 = layernorm_cf(xf, p["g1"], p["b1"], eps_ln)
    z = linear_cf(n1, p["win"])
    m = swnmf_forward(np.ascontiguousarray(z.reshape(shape)), p["u0"], p["v0"], H, d, grid, patch, shifts,
                      relu=True, solver="hals", num_iters=num_iters).reshape(xf.shape)
    x1 = xf + linear_cf(m, p["wout"], p["bout"])
    n2, xh2, r2 = layernorm_cf(x1, p["g2"], p["b2"], eps_ln)
    h = linear_cf(n2, p["w1"], p["bb1"])
    a = gelu(h)
    out = x1 + linear_cf(a, p["w2"], p["bb2"])
    if keep:
        return out.reshape(shape), dict(p=p, xh1=xh1, r1=r1, n1=n1, z=z, m=m, xh2=xh2, r2=r2, n2=n2, h=h, a=a)
    return out.reshape(shape)


def block_backward(x, dout, sd, H, d, grid, patch, shifts, num_iters=5, eps_ln=1e-5):
    """Gradients of <dout, block_forward(x)> w.r.t. x and every parameter (reference parameter names)."""
    shape = x.shape
    _, c = block_forward(x, sd, H, d, grid, patch, shifts, num_iters, eps_ln, keep=True)
    p = c["p"]
    do = dout.reshape(shape[0], shape[1], -1)
    grads = {}
    grads["mlp.block.3.linear.weight"] = np.einsum("bov,bjv->oj", do, c["a"])[:, :, None]
    grads["mlp.block.3.linear.bias"] = do.sum(axis=(0, 2))
    dh = np.einsum("oj,bov->bjv", p["w2"], do) * gelu_grad(c["h"])
    grads["mlp.block.0.linear.weight"] = np.einsum("bjv,bcv->jc", dh, c["n2"])[:, :, None]
    grads["mlp.block.0.linear.bias"] = dh.sum(axis=(0, 2))
    dn2 = np.einsum("jc,bjv->bcv", p["w1"], dh)
    dx1, grads["norm2.norm.weight"], grads["norm2.norm.bias"] = layernorm_cf_backward(dn2, c["xh2"], c["r2"], p["g2"])
    dx1 = dx1 + do
    grads["fact.out_proj.linear.weight"] = np.einsum("bov,bcv->oc", dx1, c["m"])[:, :, None]
    grads["fact.out_proj.linear.bias"] = dx1.sum(axis=(0, 2))
    dm = np.einsum("oc,bov->bcv", p["wout"], dx1)
    dz = swnmf_backward(np.ascontiguousarray(c["z"].reshape(shape)), np.ascontiguousarray(dm.reshape(shape)), p["u0"],
                        p["v0"], H, d, grid, patch, shifts, relu=True, solver="hals",
                        num_iters=num_iters).reshape(do.shape)
    grads["fact.in_proj.linear.weight"] = np.einsum("bov,bcv->oc", dz, c["n1"])[:, :, None]
    dn1 = np.einsum("oc,bov->bcv", p["win"], dz)
    dx, grads["norm1.norm.weight"], grads["norm1.norm.bias"] = layernorm_cf_backward(dn1, c["xh1"], c["r1"], p["g1"])
    return (dx + dx1).reshape(shape), grads
