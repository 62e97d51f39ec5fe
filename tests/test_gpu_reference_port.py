"""The reference's own tests/test_factorizer.py (:14-165), ported onto this package on CUDA tensors: same
constructor calls, same shapes (64^3, Matricize(num_heads=1, grid_size=1) = one 16..32 x 262144 matrix per sample,
MU rank 1; the whole model with SWMatricize(num_heads=8, patch_size=4) = d x 64 windows, HALS), same assertions,
plus a backward pass through each (the reference only checks the forward)."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

SPATIAL = (64, 64, 64)
GLOBAL = dict(reshape=None, act=nn.ReLU, factorize=None, rank=1, num_iters=5, init="uniform", solver="mu", dropout=0.1)


@pytest.fixture(scope="module")
def ft():
    import factorizer_b200
    return factorizer_b200


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _kw(ft, **extra):
    kw = dict(GLOBAL, reshape=(ft.Matricize, {"num_heads": 1, "grid_size": 1}), factorize=ft.NMF)
    kw.update(extra)
    return kw


def _check(y, shape, module, x):
    assert tuple(y.shape) == shape
    assert torch.isfinite(y).all()
    y.sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
    for p in module.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()


def test_factmixer(ft, dev):
    from factorizer_b200 import _lib
    x = torch.rand(1, 16, *SPATIAL, device=dev, requires_grad=True)
    m = ft.FactMixer(in_channels=16, out_channels=16, spatial_size=SPATIAL, num_grad_steps=None, **_kw(ft)).to(dev)
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) > 0
    y = m(x)
    assert _lib.lib().fz_last_path() == 5        # one grid-wide pass per sweep (csrc/fz_nmf_big.cu)
    _check(y, tuple(x.shape), m, x)


def test_factorizer_block(ft, dev):
    x = torch.rand(1, 16, *SPATIAL, device=dev, requires_grad=True)
    m = ft.FactorizerBlock(channels=16, spatial_size=SPATIAL, num_grad_steps=None, mlp_ratio=2, **_kw(ft)).to(dev)
    _check(m(x), tuple(x.shape), m, x)


def test_factorizer_stage(ft, dev):
    x = torch.rand(1, 16, *SPATIAL, device=dev, requires_grad=True)
    m = ft.FactorizerStage(in_channels=16, out_channels=32, spatial_size=SPATIAL, depth=2, mlp_ratio=3, **_kw(ft)).to(dev)
    _check(m(x), (1, 32, *SPATIAL), m, x)


def test_factorizer_model(ft, dev):
    m = ft.Factorizer(in_channels=4, out_channels=3, spatial_size=SPATIAL, encoder_depth=(1, 1, 1, 1),
                      encoder_width=(32, 64, 128, 256), strides=(1, 2, 2, 2), decoder_depth=(1, 1, 1),
                      reshape=(ft.SWMatricize, {"num_heads": 8, "patch_size": 4}), act=nn.ReLU, factorize=ft.NMF, rank=1,
                      num_iters=5, num_grad_steps=None, init="uniform", solver="hals", mlp_ratio=2, dropout=0.1).to(dev)
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) > 0
    x = torch.rand(1, 4, *SPATIAL, device=dev, requires_grad=True)
    _check(m(x), (1, 3, *SPATIAL), m, x)
    for batch in (2, 3):
        with torch.no_grad():
            assert m(torch.rand(batch, 4, *SPATIAL, device=dev)).shape[0] == batch
