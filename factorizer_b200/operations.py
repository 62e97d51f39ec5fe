"""Matricize / SWMatricize: the reshape that turns a (B, C, *spatial) activation into the batch of
``head_dim x patch-volume`` matrices the NMF layer factorises, and back.

Drop-in for factorizer/factorization/operations.py:147-434 of the reference: same constructor
arguments, ``output_size``, ``forward`` and ``inverse_forward``.  The reference builds these from
einops ``Rearrange`` + ``torch.roll`` + ``torch.cat`` (four full-size materialisations); here the window
geometry is resolved once at construction and each direction is a single gather / scatter kernel
(csrc/fz_swmat.cu) whose results are bit-identical to the reference's.
"""
from __future__ import annotations

import math
import re
from typing import Optional, Sequence

import torch
from torch import Tensor, nn

from . import _ops

__all__ = ["Reshape", "Matricize", "SWMatricize", "dot", "norm2", "relative_error"]


# ---- diagnostic helpers (reference operations.py:13-51, 99-122); not on the fwd/bwd path ---------
def dot(x: Tensor, y: Tensor) -> Tensor:
    return (x * y).sum(dim=(-2, -1)).unsqueeze(-1)


def norm2(x: Tensor, w: Optional[Tensor] = None) -> Tensor:
    y = x.flatten(1).square()
    if w is not None:
        y = y * w.flatten(1)
    return torch.sqrt(y.sum(dim=1))


def relative_error(x: Tensor, y: Tensor, w: Optional[Tensor] = None, eps: float = 1e-16) -> Tensor:
    return (norm2(x - y, w) + eps) / (norm2(x, w) + eps)


def _ntuple(value, n: int) -> tuple:
    if isinstance(value, (tuple, list)):
        return tuple(value)
    return (value,) * n


_MATRICIZE_RE = re.compile(r"^b \(h d\)( \(g\d+ p\d+\))+$")


class Reshape(nn.Module):
    """Base of Matricize.  The reference accepts any einops equation here
    (operations.py:147-280); only the identity and the matricize pattern are on the hot path, so
    anything else raises instead of silently running somewhere slow."""

    def __init__(self, input_size: Sequence[Optional[int]], equation: Optional[str] = None,
                 shifts: Optional[Sequence[int]] = None, dims: Optional[Sequence[int]] = None,
                 **kwargs) -> None:
        super().__init__()
        self.input_size = input_size
        self._geom: Optional[_ops.Geometry] = None
        if equation is None:
            self.output_size = input_size
        else:
            self.equation = equation
            left, right = (side.strip() for side in equation.split("->"))
            self.left, self.right = left, right
            self.equation_inv = " -> ".join([right, left])
            if not _MATRICIZE_RE.match(left):
                raise NotImplementedError(
                    f"factorizer_b200.Reshape implements the Matricize pattern only, got {equation!r}")
            self._resolve(input_size, kwargs)
        if shifts is not None:
            # the reference only defines these attributes for shifted windows (operations.py:191-194)
            self.shifts = tuple(shifts)
            self.shifts_inv = tuple(-s for s in self.shifts)
            self.dims = dims
        if equation is not None:
            n = len(self._patch)
            roll = (0,) * n if shifts is None else tuple(int(s) for s in shifts)
            self._geom = _ops.Geometry(self._channels, self._spatial, self._patch, self.dim_lengths["d"], [roll])

    # -- geometry inference, same rules as Reshape.infer_dims / compute_size (operations.py:196-264)
    def _resolve(self, input_size, known) -> None:
        batch, channels, *spatial = input_size
        n = len(spatial)
        if channels is None or any(s is None for s in spatial):
            raise ValueError("Matricize needs concrete channel and spatial sizes in `input_size`")
        lengths = {}
        h, d = known.get("h"), known.get("d")
        if h is None and d is None:
            raise ValueError("'num_heads' or 'head_dim' must be specified.")
        if h is None:
            h = channels // d
        elif d is None:
            d = channels // h
        lengths["h"], lengths["d"] = h, d
        if h * d != channels:
            raise ValueError(f"channels={channels} cannot be split into {h} heads of dim {d}")
        grid, patch = [], []
        for k, s in enumerate(spatial):
            g, p = known.get(f"g{k}"), known.get(f"p{k}")
            if g is None and p is None:
                raise ValueError("'grid_size' or 'patch_size' must be specified.")
            if g is None:
                g = s // p
            elif p is None:
                p = s // g
            if g * p != s:
                raise ValueError(f"spatial size {s} (axis {k}) is not grid {g} x patch {p}")
            lengths[f"g{k}"], lengths[f"p{k}"] = g, p
            grid.append(g)
            patch.append(p)
        if batch is not None:
            lengths["b"] = batch
        self.dim_lengths = lengths
        self._channels, self._spatial, self._patch, self._grid = channels, tuple(spatial), tuple(patch), tuple(grid)
        lead = None if batch is None else batch * h
        self.output_size = (lead, math.prod(grid), d, math.prod(patch))

    def forward(self, x: Tensor) -> Tensor:
        if self._geom is None:
            return x
        return _ops.SWMatForward.apply(x, self._geom)

    def inverse_forward(self, x: Tensor) -> Tensor:
        if self._geom is None:
            return x
        return _ops.SWMatInverse.apply(x, self._geom, False)


class Matricize(Reshape):
    """``b (h d) (g0 p0) (g1 p1) .. -> (b h) (g0 g1 ..) d (p0 p1 ..)`` with an optional cyclic shift
    applied first (reference operations.py:283-355)."""

    def __init__(self, input_size: Sequence[Optional[int]], num_heads: Optional[int] = None,
                 head_dim: Optional[int] = None, grid_size=None, patch_size=None, shifts=None,
                 **kwargs) -> None:
        assert (num_heads, head_dim) != (None, None), "'num_heads' or 'head_dim' must be specified."
        assert (grid_size, patch_size) != (None, None), "'grid_size' or 'kernel_size' must be specified."
        n = len(input_size) - 2
        left = "b (h d) " + " ".join(f"(g{i} p{i})" for i in range(n))
        right = "(b h) (" + " ".join(f"g{i}" for i in range(n)) + ") d (" + " ".join(f"p{i}" for i in range(n)) + ")"
        lengths = {}
        if num_heads is not None:
            lengths["h"] = max(num_heads, 1)
        if head_dim is not None:
            lengths["d"] = max(head_dim, 1)
        for k, g in enumerate(_ntuple(grid_size, n)):
            if g is not None:
                lengths[f"g{k}"] = max(g, 1)
        for k, p in enumerate(_ntuple(patch_size, n)):
            if p is not None:
                lengths[f"p{k}"] = max(p, 1)
        if shifts is not None:
            dims = tuple(k + 2 for k in range(n))
            shifts = _ntuple(shifts, n)
        else:
            dims = None
        super().__init__(input_size, equation=f"{left} -> {right}", shifts=shifts, dims=dims, **lengths, **kwargs)


class SWMatricize(nn.Module):
    """Shifted-window matricize: one Matricize per shift, concatenated along dim 0; the inverse
    averages the window sets (reference operations.py:358-434)."""

    def __init__(self, input_size: Sequence[Optional[int]], num_heads: Optional[int] = None,
                 head_dim: Optional[int] = None, grid_size=None, patch_size=None,
                 shifts: Optional[Sequence] = None, **kwargs) -> None:
        super().__init__()
        n = len(input_size) - 2
        patch_size = _ntuple(patch_size, n)
        grid_size = _ntuple(grid_size, n)
        if shifts is None:
            shifts = [None, tuple(p // 2 for p in patch_size)]
        self.shifted_windows = nn.ModuleList(
            Matricize(input_size, num_heads=num_heads, head_dim=head_dim, grid_size=grid_size,
                      patch_size=patch_size, shifts=s, **kwargs)
            for s in shifts)
        first = self.shifted_windows[0]
        self.output_size = first.output_size
        self._geom = _ops.Geometry(first._channels, first._spatial, first._patch, first.dim_lengths["d"],
                                   [w._geom.shifts[0] for w in self.shifted_windows])

    def forward(self, x: Tensor) -> Tensor:
        return _ops.SWMatForward.apply(x, self._geom)

    def inverse_forward(self, x: Tensor) -> Tensor:
        return _ops.SWMatInverse.apply(x, self._geom, True)
