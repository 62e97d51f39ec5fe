"""Differentiable NMF layer.  Drop-in for the NMF part of
factorizer/factorization/matrix_factorization.py of the reference: ``NMF`` /
``MatrixFactorization`` keep their constructor and ``forward / decompose / reconstruct / loss``
signatures, ``init.u0`` / ``init.v0`` stay registered buffers with the same RNG consumption, so
reference checkpoints load and same-seed modules get the same buffers.

What differs is the execution: the reference unrolls the solver in Python (12 ATen launches per
iteration, autograd replay for the backward); here the whole unrolled solve -- and its hand-derived
adjoint -- is one CUDA kernel per direction (csrc/fz_nmf_generic.cu, csrc/fz_swnmf_fast.cu).  Only
the solvers named by the north star are implemented ('mu' = MultiplicativeUpdate,
'hals'/'nncd' = CoordinateDescent + ReLU); the other names stay importable and raise
NotImplementedError -- there is no CPU or eager fallback.
"""
from __future__ import annotations

import math
from typing import Any, Optional, Sequence

import torch
from torch import Tensor, nn

from . import _lib, _ops
from .helpers import as_tuple, is_partializable, partialize
from .operations import relative_error

__all__ = [
    "Initializer", "RandomInit", "SVDInit", "NNDSVDInit", "BCDSolver", "LeastSquares",
    "ProjectedGradient", "CoordinateDescent", "MultiplicativeUpdate", "FastMultiplicativeUpdate",
    "WeightedMultiplicativeUpdate", "SemiMultiplicativeUpdate", "Compose", "SVD",
    "MatrixFactorization", "NMF", "INIT_DISPATCH_MAP", "SOLVER_DISPATCH_MAP",
]


def _out_of_scope(name: str, where: str):
    class _Unsupported(nn.Module):
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(
                f"factorizer_b200: `{name}` ({where} in the reference) is outside the B200 hot path "
                "(NMF with the 'mu' and 'hals' solvers and RandomInit); no CPU/eager fallback exists")

    _Unsupported.__name__ = _Unsupported.__qualname__ = name
    return _Unsupported


# ---- initialisers ------------------------------------------------------------------------------------
class Initializer(nn.Module):
    def forward(self, x: Tensor):
        raise NotImplementedError(f"Subclass {self.__class__.__name__} must implement this method.")


class RandomInit(Initializer):
    """Two fixed random buffers shared by every matrix (reference matrix_factorization.py:28-58).
    ``u0`` is drawn before ``v0`` from the global CPU generator, exactly like the reference."""

    def __init__(self, rank: int, size, method="uniform") -> None:
        super().__init__()
        method = as_tuple(method)
        if len(method) == 1:
            self.method = (method[0], method[0])
        elif len(method) == 2:
            self.method = tuple(method)
        else:
            raise ValueError("`method` not valid.")
        self.register_buffer("u0", torch.empty(size[0], rank))
        getattr(nn.init, f"{self.method[0]}_")(self.u0)
        self.register_buffer("v0", torch.empty(size[1], rank))
        getattr(nn.init, f"{self.method[1]}_")(self.v0)

    def forward(self, x: Tensor):
        batch = tuple(x.shape[:-2])
        return self.u0.expand(*batch, *self.u0.shape), self.v0.expand(*batch, *self.v0.shape)


SVDInit = _out_of_scope("SVDInit", "matrix_factorization.py:61-72")
NNDSVDInit = _out_of_scope("NNDSVDInit", "matrix_factorization.py:75-100")


# ---- solvers: configuration carriers; the arithmetic lives in the CUDA kernels -------------------------
class BCDSolver(nn.Module):
    """Block-coordinate-descent solver description (reference matrix_factorization.py:108-136)."""

    kind: Optional[int] = None

    def __init__(self, factor: Sequence[int] = (0, 1), *args, **kwargs) -> None:
        super().__init__()
        self.factor = as_tuple(factor)
        assert set(self.factor).issubset({0, 1}), "`factor` elements must be 0 or 1."

    def kernel_kind(self) -> int:
        raise NotImplementedError(
            f"factorizer_b200: solver {self.__class__.__name__} has no CUDA kernel (only 'mu' and 'hals')")

    def forward(self, x: Tensor, factor_matrices):
        raise NotImplementedError(
            "factorizer_b200 runs the whole unrolled solve inside one kernel; call the solver through "
            "NMF.forward / NMF.decompose instead of stepping it from Python")


class MultiplicativeUpdate(BCDSolver):
    """'mu' (reference matrix_factorization.py:232-247)."""

    def __init__(self, factor: Sequence[int] = (0, 1), eps: float = 1e-16, **kwargs) -> None:
        super().__init__(factor=factor)
        self.eps = eps

    def kernel_kind(self) -> int:
        if self.factor != (0, 1):
            raise NotImplementedError("factorizer_b200: only factor=(0, 1) sweeps are implemented")
        return _lib.FZ_SOLVER_MU


class CoordinateDescent(BCDSolver):
    """'hals' / 'nncd' when ``project`` is ReLU (reference matrix_factorization.py:194-229, :609)."""

    def __init__(self, factor: Sequence[int] = (0, 1), eps: float = 1e-16, project=None, **kwargs):
        super().__init__(factor=factor)
        self.eps = eps
        self.project = partialize(nn.Identity if project is None else project)()

    def kernel_kind(self) -> int:
        if self.factor != (0, 1):
            raise NotImplementedError("factorizer_b200: only factor=(0, 1) sweeps are implemented")
        if not isinstance(self.project, nn.ReLU):
            raise NotImplementedError(
                "factorizer_b200: CoordinateDescent is implemented with project=nn.ReLU ('hals') only")
        return _lib.FZ_SOLVER_HALS


LeastSquares = _out_of_scope("LeastSquares", "matrix_factorization.py:139-165")
ProjectedGradient = _out_of_scope("ProjectedGradient", "matrix_factorization.py:168-191")
FastMultiplicativeUpdate = _out_of_scope("FastMultiplicativeUpdate", "matrix_factorization.py:250-274")
WeightedMultiplicativeUpdate = _out_of_scope("WeightedMultiplicativeUpdate", "matrix_factorization.py:277-316")
SemiMultiplicativeUpdate = _out_of_scope("SemiMultiplicativeUpdate", "matrix_factorization.py:319-341")
Compose = _out_of_scope("Compose", "matrix_factorization.py:344-378")
SVD = _out_of_scope("SVD", "matrix_factorization.py:386-451")

INIT_DISPATCH_MAP = {
    "uniform": (RandomInit, {"method": "uniform"}),
    "normal": (RandomInit, {"method": "normal"}),
    "normal-uniform": (RandomInit, {"method": ("normal", "uniform")}),
    "uniform-normal": (RandomInit, {"method": ("uniform", "normal")}),
    "svd": SVDInit,
    "nndsvd": NNDSVDInit,
}

SOLVER_DISPATCH_MAP = {
    "mu": MultiplicativeUpdate,
    "hals": (CoordinateDescent, {"project": nn.ReLU}),
    "nncd": (CoordinateDescent, {"project": nn.ReLU}),
}
# every other name of the reference's table (matrix_factorization.py:590-618) is out of scope
for _name in ("mu-0", "mu-1", "fmu", "fmu-0", "fmu-1", "wmu", "wmu-0", "wmu-1", "smu", "smu-0", "smu-1",
              "cd", "cd-0", "cd-1", "nncd-0", "nncd-1", "hals-0", "hals-1", "ls", "ls-0", "ls-1", "nnls",
              "nnls-0", "nnls-1"):
    SOLVER_DISPATCH_MAP[_name] = _out_of_scope(f"solver '{_name}'", "matrix_factorization.py:590-618")


def _parse(obj: Any, table: dict, what: str):
    if isinstance(obj, str):
        if obj not in table:
            raise ValueError(f"unknown {what} {obj!r}")
        return table[obj]
    if is_partializable(obj):
        return obj
    raise NotImplementedError(f"factorizer_b200: {what} specification {obj!r} is not supported")


class MatrixFactorization(nn.Module):
    """X ~ U V^T with U, V updated by ``num_iters`` block-coordinate sweeps
    (reference matrix_factorization.py:454-546)."""

    def __init__(self, size: Sequence[int], rank: Optional[int] = None, compression: float = 10,
                 init="normal", solver="cd", num_iters: int = 5, num_grad_steps: Optional[int] = None,
                 verbose: bool = False, **kwargs) -> None:
        super().__init__()
        self.size = M, N = tuple(size)
        self.num_iters = num_iters
        self.num_grad_steps = num_iters if num_grad_steps is None else num_grad_steps
        assert (rank, compression) != (None, None), "'rank' or 'compression' must be specified."
        if rank is None:
            rank = max(math.ceil(M * N / (compression * (M + N))), 1)
        self.rank = rank
        self.compression = M * N / (rank * (M + N))
        self.init = partialize(_parse(init, INIT_DISPATCH_MAP, "init"))(size=self.size, rank=rank)
        self.solver = partialize(_parse(solver, SOLVER_DISPATCH_MAP, "solver"))(size=self.size, rank=rank)
        self.verbose = verbose
        if not isinstance(self.init, RandomInit):
            raise NotImplementedError("factorizer_b200: only RandomInit initialisers are implemented")
        if rank > _lib.FZ_MAX_RANK:
            raise NotImplementedError(f"factorizer_b200: rank {rank} > {_lib.FZ_MAX_RANK} not implemented")
        self._kind = self.solver.kernel_kind()

    # -- kernel plumbing -------------------------------------------------------------------------------
    def solver_spec(self, num_iters: Optional[int] = None) -> _ops.SolverSpec:
        T = self.num_iters if num_iters is None else num_iters
        k = max(0, min(self.num_grad_steps, T))
        return _ops.SolverSpec(self._kind, self.rank, T, k, getattr(self.solver, "eps", 1e-16))

    def decompose(self, x: Tensor, *args, **kwargs):
        if self.verbose:
            self._print_losses(x)
        return _ops.NMFDecompose.apply(x, self.init.u0, self.init.v0, self.solver_spec(), self.size)

    def reconstruct(self, u: Tensor, v: Tensor) -> Tensor:
        return u @ v.mT

    def loss(self, x: Tensor, u: Tensor, v: Tensor, w: Optional[Tensor] = None) -> Tensor:
        return relative_error(x, self.reconstruct(u, v), w)

    def forward(self, x: Tensor) -> Tensor:
        if self.verbose:
            self._print_losses(x)
        return _ops.NMFReconstruct.apply(x, self.init.u0, self.init.v0, self.solver_spec(), self.size)

    def _print_losses(self, x: Tensor) -> None:
        # reference prints the loss of the current factors before each sweep (:524-526)
        with torch.no_grad():
            for it in range(1, self.num_iters + 1):
                u, v = _ops.NMFDecompose.apply(x, self.init.u0, self.init.v0, self.solver_spec(it - 1), self.size)
                print(f"iter {it}, loss = {self.loss(x, u, v)}")


class NMF(MatrixFactorization):
    """Non-negative matrix factorisation layer: defaults init='uniform', solver='hals'
    (reference matrix_factorization.py:549-578)."""

    def __init__(self, size, rank: Optional[int] = None, compression: float = 10, num_iters: int = 5,
                 num_grad_steps: Optional[int] = None, init="uniform", solver="hals",
                 verbose: bool = False, **kwargs):
        super().__init__(size, rank=rank, compression=compression, num_iters=num_iters,
                         num_grad_steps=num_grad_steps, init=init, solver=solver, verbose=verbose, **kwargs)
