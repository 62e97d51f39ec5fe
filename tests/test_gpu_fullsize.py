"""Parity of the PRODUCTION kernels at the sizes that are benchmarked (BASELINE configs 2 and 3).

The golden cases are small (<= 48 tiles), so every streaming warp of the octant kernels sees at most one tile
there; here each warp streams dozens of tiles (buffer reuse, mbarrier parity flips, prefetch, dependency
waits), forward AND backward, against `oracle/nmf_oracle.c` -- the plain-C restatement of
factorizer/factorizer.py:41-50 + matrix_factorization.py:224-227 that tests/test_oracle.py pins on the
reference-generated golden vectors.  Tolerance: rtol 1e-4 / atol 1e-5 (north_star, fp32)."""
import numpy as np
import pytest
import torch
from torch import nn

from conftest import assert_close, tol_ratio
from oracle import c_oracle as CO
from oracle.block_reference import block_reference

pytestmark = pytest.mark.gpu

SHIFTS = [(0, 0, 0), (4, 4, 4)]


@pytest.fixture(scope="module")
def ft():
    import factorizer_b200
    return factorizer_b200


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    CO.use_all_cores()
    return torch.device("cuda:0")


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("shape", [(1, 32, 64, 64, 64), (1, 32, 128, 128, 128), (2, 16, 64, 128, 32), (3, 8, 40, 24, 56)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("k", [None, 2], ids=["full", "k2"])
@pytest.mark.parametrize("scheme", ["pipeline", "three-launch"])
def test_core_full_size_vs_c_oracle(ft, dev, shape, k, scheme):
    """BASELINE config 2 (and neighbours with batch > 1 / non-power-of-two grids): y and dx of the fused
    SWMatricize + ReLU + HALS rank-1 + inverse op through the production path."""
    from factorizer_b200 import _lib, _ops
    if k is not None and shape[2] == 128:
        pytest.skip("truncated unroll is covered at the smaller sizes")
    rng = np.random.default_rng(abs(hash(shape)) % (1 << 31))
    x_np = rng.standard_normal(shape, dtype=np.float32)
    gy_np = rng.standard_normal(shape, dtype=np.float32)
    sw = ft.SWMatricize((None, *shape[1:]), head_dim=8, patch_size=8)
    sw._geom.path = _lib.FZ_PATH_OCTANT_PIPELINE if scheme == "pipeline" else _lib.FZ_PATH_OCTANT_3LAUNCH
    torch.manual_seed(3)
    nmf = ft.NMF((8, 512), rank=1, num_iters=5, num_grad_steps=k, init="uniform", solver="hals").to(dev)
    x = torch.from_numpy(x_np).to(dev).requires_grad_(True)
    y = _ops.SWNMF.apply(x, nmf.init.u0, nmf.init.v0, sw._geom, nmf.solver_spec(), True)
    assert _lib.lib().fz_last_path() == (6 if scheme == "pipeline" else 2)
    (gx,) = torch.autograd.grad((y * torch.from_numpy(gy_np).to(dev)).sum(), x)
    torch.cuda.synchronize()
    v0 = _np(nmf.init.v0)
    y_ref = CO.swnmf_forward(x_np, v0, 8, (8, 8, 8), SHIFTS)
    gx_ref = CO.swnmf_backward(x_np, gy_np, v0, 8, (8, 8, 8), SHIFTS, num_grad_steps=k)
    assert_close(_np(y), y_ref, what="y")
    assert_close(_np(gx), gx_ref, what="gx")
    # a second call on the same buffers gives the same bits (no stale scratch state between calls)
    y2 = _ops.SWNMF.apply(x, nmf.init.u0, nmf.init.v0, sw._geom, nmf.solver_spec(), True)
    (gx2,) = torch.autograd.grad((y2 * torch.from_numpy(gy_np).to(dev)).sum(), x)
    assert torch.equal(y2, y) and torch.equal(gx2, gx)


@pytest.mark.parametrize("shape", [(1, 32, 64, 64, 64), (2, 16, 32, 48, 64), (1, 8, 16, 16, 16), (1, 32, 128, 128, 128)],
                         ids=lambda s: "x".join(map(str, s)))
def test_core_bf16_activations(ft, dev, shape):
    """bf16 volumes (north_star: "a stated looser bound applies for bf16 activations"): x, y, dy, dx travel as bf16, the
    arithmetic and the factors stay fp32.  Reference = the fp32 C oracle on the SAME bf16-representable inputs, so the
    only differences are the kernels' fp32 arithmetic and the final rounding of y / dx to bf16 (relative 2^-9):
    bound rtol 1e-2 / atol 1e-3."""
    from factorizer_b200 import _lib, _ops
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32)).to(dev).bfloat16()
    gy = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32)).to(dev).bfloat16()
    sw = ft.SWMatricize((None, *shape[1:]), head_dim=8, patch_size=8)
    torch.manual_seed(3)
    nmf = ft.NMF((8, 512), rank=1, num_iters=5, init="uniform", solver="hals").to(dev)
    xr = x.clone().requires_grad_(True)
    y = _ops.SWNMF.apply(xr, nmf.init.u0, nmf.init.v0, sw._geom, nmf.solver_spec(), True)
    assert y.dtype == torch.bfloat16 and _lib.lib().fz_last_path() == 2
    (gx,) = torch.autograd.grad(y, xr, gy)
    assert gx.dtype == torch.bfloat16
    torch.cuda.synchronize()
    v0 = _np(nmf.init.v0)
    x_np, gy_np = _np(x.float()), _np(gy.float())
    y_ref = CO.swnmf_forward(x_np, v0, 8, (8, 8, 8), SHIFTS)
    gx_ref = CO.swnmf_backward(x_np, gy_np, v0, 8, (8, 8, 8), SHIFTS)
    assert_close(_np(y.float()), y_ref, rtol=1e-2, atol=1e-3, what="y (bf16)")
    assert_close(_np(gx.float()), gx_ref, rtol=1e-2, atol=1e-3, what="gx (bf16)")
    # and against the fp32 kernels on the same inputs: only the output rounding differs (half a bf16 ulp, plus slack)
    y32 = _ops.SWNMF.apply(x.float(), nmf.init.u0, nmf.init.v0, sw._geom, nmf.solver_spec(), True)
    assert_close(_np(y.float()), _np(y32), rtol=4e-3, atol=1e-6, what="y bf16 vs fp32 kernels")
    # geometries the bf16 kernels do not cover are refused, not silently converted
    odd = ft.SWMatricize((None, 8, 16, 16, 24), head_dim=8, patch_size=8)
    with pytest.raises(NotImplementedError):
        _ops.SWNMF.apply(torch.zeros(1, 8, 16, 16, 24, device=dev, dtype=torch.bfloat16), nmf.init.u0, nmf.init.v0,
                         odd._geom, nmf.solver_spec(), True)


@pytest.mark.parametrize("n", [64, 128])
def test_block_full_size_vs_oracle(ft, dev, n):
    """BASELINE config 3: FactorizerBlock(32, n^3, LayerNorm, SWMatricize, HALS rank 1, mlp_ratio 2) through the
    fused glue kernels + production core: output, input gradient and every parameter gradient."""
    C = 32
    torch.manual_seed(11)
    blk = ft.FactorizerBlock(channels=C, spatial_size=(n, n, n), norm=ft.LayerNorm,
                             reshape=(ft.SWMatricize, {"head_dim": 8, "patch_size": 8}), act=nn.ReLU,
                             factorize=ft.NMF, rank=1, num_iters=5, init="uniform", solver="hals", mlp_ratio=2,
                             dropout=0.0).to(dev)
    with torch.no_grad():      # non-trivial affine parameters
        for nm in ("norm1", "norm2"):
            getattr(blk, nm).norm.weight.add_(0.2 * torch.randn(C, device=dev))
            getattr(blk, nm).norm.bias.add_(0.2 * torch.randn(C, device=dev))
    x = torch.randn(1, C, n, n, n, device=dev).requires_grad_(True)
    gy = torch.randn(1, C, n, n, n, device=dev)
    assert blk._fused_args(x) is not None
    y = blk(x)
    assert type(y.grad_fn).__name__ == "FactorizerBlockFnBackward"
    params = dict(blk.named_parameters())
    grads = torch.autograd.grad((y * gy).sum(), [x] + list(params.values()))
    torch.cuda.synchronize()
    # the product's own z, through the same C entry point the block calls
    from factorizer_b200 import _lib
    z = torch.empty_like(x)
    n1 = blk.norm1.norm
    _lib.check(_lib.lib().fz_ln_linear_forward(x.data_ptr(), n1.weight.data_ptr(), n1.bias.data_ptr(),
                                               blk.fact.in_proj.linear.weight.data_ptr(), z.data_ptr(), 1, C, n ** 3,
                                               float(n1.eps), torch.cuda.current_stream().cuda_stream))
    y_ref, gx_ref, gp_ref, _ = block_reference(blk.state_dict(), x, gy, z,
                                               lambda got, ref: assert_close(_np(got), _np(ref), what="block z = in_proj(norm1(x))"))
    assert_close(_np(y), _np(y_ref), what="block y")
    assert_close(_np(grads[0]), _np(gx_ref), what="block gx")
    for (k, _), gp in zip(params.items(), grads[1:]):
        ref = _np(gp_ref[k]).reshape(_np(gp).shape)
        # a parameter gradient is a sum over n^3 voxels of fp32 products: bound relative to its largest entry
        scale = max(1.0, float(np.abs(ref).max()))
        assert tol_ratio(_np(gp) / scale, ref / scale, rtol=1e-4, atol=1e-4) <= 1.0, k


@pytest.mark.parametrize("C,n,B,ratio", [(64, 32, 1, 2), (128, 16, 2, 2), (512, 8, 1, 2), (72, 16, 1, 4), (16, 16, 2, 1.5)])
def test_wide_block_vs_oracle(ft, dev, C, n, B, ratio):
    """FactorizerBlock at the widths of the Swin Factorizer's deeper stages (and odd ones): the channel-map kernels with fused
    epilogues (FZ_EPILOGUE_*) + LayerNorm backward with the residual gradient around the production core
    (_ops.FactorizerBlockWideFn): output, input gradient and every parameter gradient against oracle/block_reference.py."""
    torch.manual_seed(13 + C)
    blk = ft.FactorizerBlock(channels=C, spatial_size=(n, n, n), norm=ft.LayerNorm,
                             reshape=(ft.SWMatricize, {"head_dim": 8, "patch_size": 8}), act=nn.ReLU,
                             factorize=ft.NMF, rank=1, num_iters=5, init="uniform", solver="hals", mlp_ratio=ratio,
                             dropout=0.0).to(dev)
    with torch.no_grad():
        for nm in ("norm1", "norm2"):
            getattr(blk, nm).norm.weight.add_(0.2 * torch.randn(C, device=dev))
            getattr(blk, nm).norm.bias.add_(0.2 * torch.randn(C, device=dev))
    x = torch.randn(B, C, n, n, n, device=dev).requires_grad_(True)
    gy = torch.randn(B, C, n, n, n, device=dev)
    assert blk._fused_args(x) is not None
    y = blk(x)
    assert type(y.grad_fn).__name__ == "FactorizerBlockWideFnBackward"
    params = dict(blk.named_parameters())
    grads = torch.autograd.grad((y * gy).sum(), [x] + list(params.values()))
    with torch.no_grad():
        assert torch.equal(blk(x.detach()), y.detach())            # inference path: same kernels, nothing saved
        # the product's own z, through the same kernels the block calls (LayerNorm kernel + channel-map kernel)
        from factorizer_b200 import _ops
        n1 = blk.norm1.norm
        z = _ops._channel_map(_ops._ln_forward(x.detach().view(B, C, -1), n1.weight, n1.bias, n1.eps),
                              blk.fact.in_proj.linear.weight.squeeze(-1), None).view(x.shape)
    torch.cuda.synchronize()
    y_ref, gx_ref, gp_ref, _ = block_reference(blk.state_dict(), x, gy, z,
                                               lambda got, ref: assert_close(_np(got), _np(ref), what="block z = in_proj(norm1(x))"))
    assert_close(_np(y), _np(y_ref), what="block y")
    assert_close(_np(grads[0]), _np(gx_ref), what="block gx")
    for (k, _), gp in zip(params.items(), grads[1:]):
        ref = _np(gp_ref[k]).reshape(_np(gp).shape)
        scale = max(1.0, float(np.abs(ref).max()))
        assert tol_ratio(_np(gp) / scale, ref / scale, rtol=1e-4, atol=1e-4) <= 1.0, k
