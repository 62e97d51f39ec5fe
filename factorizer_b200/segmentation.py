"""Encoder-decoder scaffold that carries FactorizerStages, and ``ft.Factorizer`` itself (BASELINE configs 4-5;
SURVEY.md section 8(f) row 3).  Constructor signatures, forward semantics and ``state_dict`` keys follow the
reference (factorizer/unet.py:177-276 ``UNet``, factorizer/factorizer.py:125-171 ``Factorizer``) so its
checkpoints load; the stem / strided down- and up-sampling convolutions / 1x1 head stay cuDNN library calls
(they are not on the hot path), every FactorizerBlock inside runs the kernels of this package.

Module tree (= parameter names):
    stem | encoder.blocks[i].{downsample, block} | decoder.blocks[i].{upsample, block} | head or heads[j]
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
from torch import nn

from .factorizer import FactorizerStage
from .helpers import as_tuple, partialize
from . import layers
from .layers import PositionalEmbedding

__all__ = ["UNet", "Factorizer"]


class _Down(nn.Module):
    """One encoder level: strided convolution (identity at stride 1), then the stage."""

    def __init__(self, cin, cout, depth, stride, downsample, block, **kw):
        super().__init__()
        if math.prod(as_tuple(stride)) == 1:
            self.downsample = nn.Identity()
        else:
            # the reference always builds the down-sampler with stride=2 (unet.py:53)
            self.downsample = partialize(downsample)(cin, cout, stride=2)
        self.block = partialize(block)(cout, cout, depth=depth, **kw)

    def forward(self, x):
        return self.block(self.downsample(x))


class _Up(nn.Module):
    """One decoder level: transposed convolution, concatenate the skip (skip first), then the stage."""

    def __init__(self, cin, cout, depth, stride, upsample, block, **kw):
        super().__init__()
        self.upsample = partialize(upsample)(cin, cout, stride=stride)
        self.block = partialize(block)(2 * cout, cout, depth=depth, **kw)

    def forward(self, deep, skip):
        return self.block(torch.cat([skip, self.upsample(deep)], dim=1))


class _Levels(nn.Module):
    def __init__(self, levels):
        super().__init__()
        self.blocks = nn.ModuleList(levels)


def _scaled(size, stride, up: bool):
    if not isinstance(size, Sequence):
        return size
    return tuple(d * stride if up else d // stride for d in size)


class UNet(nn.Module):
    """Generic U-shaped network: ``block[i]`` builds level i (encoder levels first, then decoder levels) as
    ``block[i](in_channels, out_channels, depth=..., spatial_size=..., **kwargs)``.  Unlike the reference there
    is no default convolutional block (its DoubleConv U-Net is outside the path this package covers)."""

    def __init__(self, in_channels, out_channels, spatial_dims=3, spatial_size=None,
                 encoder_depth=(1, 1, 1, 1, 1), encoder_width=(32, 64, 128, 256, 512), strides=(1, 2, 2, 2, 2),
                 decoder_depth=(1, 1, 1, 1), stem=None, downsample=None, block=None, upsample=None, head=None,
                 num_deep_supr=False, **kwargs):
        super().__init__()
        if block is None:
            raise NotImplementedError("factorizer_b200.UNet needs `block` (per-level stage specs); the reference's "
                                      "default DoubleConv U-Net is not part of this package")
        self.spatial_dims, self.spatial_size = spatial_dims, spatial_size
        conv = getattr(layers, f"Conv{spatial_dims}d")     # nn.ConvNd with a patch (kernel == stride) fast path
        tconv = getattr(layers, f"ConvTranspose{spatial_dims}d")
        downsample = downsample or (conv, {"kernel_size": 2})
        upsample = upsample or (tconv, {"kernel_size": 2})
        head = partialize(head or (conv, {"kernel_size": 1}))
        n_enc, n_dec = len(encoder_depth), len(decoder_depth)

        if stem in (None, nn.Identity):
            self.stem, width = nn.Identity(), in_channels
        else:
            self.stem, width = partialize(stem)(in_channels, encoder_width[0]), encoder_width[0]

        size, levels, chans = spatial_size, [], [width, *encoder_width]
        for i in range(n_enc):
            size = _scaled(size, strides[i], up=False)
            levels.append(_Down(chans[i], chans[i + 1], encoder_depth[i], strides[i], downsample, block[i],
                                spatial_size=size, **kwargs))
        self.encoder = _Levels(levels)

        levels, widths, up_strides = [], encoder_width[::-1], strides[::-1][:n_dec]
        for i in range(len(widths) - 1):
            size = _scaled(size, up_strides[i], up=True)
            levels.append(_Up(widths[i], widths[i + 1], decoder_depth[i], up_strides[i], upsample, block[n_enc + i],
                              spatial_size=size, **kwargs))
        self.decoder = _Levels(levels)

        if num_deep_supr in (False, None):
            self.num_deep_supr = False
            self.head = head(encoder_width[0], out_channels)
        else:
            self.num_deep_supr = 3 if num_deep_supr is True else num_deep_supr
            # as the reference (factorizer/unet.py:255-258): the stored count is 3 for True, but the heads are built
            # over range(num_deep_supr) of the ARGUMENT, i.e. a single head for True -- state_dict keys must agree
            self.heads = nn.ModuleList(head(encoder_width[j], out_channels) for j in range(int(num_deep_supr)))

    def forward_features(self, x):
        """Feature maps, finest first: decoder outputs where there is a decoder level, encoder outputs below."""
        feats = []
        out = self.stem(x)
        for level in self.encoder.blocks:
            out = level(out)
            feats.append(out)
        for i, level in enumerate(self.decoder.blocks):
            feats[-2 - i] = level(feats[-1 - i], feats[-2 - i])
        return feats

    def forward(self, x):
        feats = self.forward_features(x)
        if self.num_deep_supr:
            return [h(feats[j]) for j, h in enumerate(self.heads)]
        return self.head(feats[0])


class Factorizer(UNet):
    """Factorizer segmentation network: a U-shaped scaffold whose every level is a FactorizerStage; only the
    bottleneck gets the positional embedding (reference factorizer/factorizer.py:125-171)."""

    def __init__(self, in_channels, out_channels, spatial_size, encoder_depth=(1, 1, 1, 1, 1),
                 encoder_width=(32, 64, 128, 256, 512), strides=(1, 2, 2, 2, 2), decoder_depth=(1, 1, 1, 1),
                 stem=None, downsample=None, upsample=None, head=None, pos_embed=PositionalEmbedding,
                 num_deep_supr=False, **kwargs):
        nd = len(spatial_size)
        if stem is None:
            stem = (getattr(layers, f"Conv{nd}d"), {"kernel_size": 3, "padding": 1, "bias": False})
        plain = (FactorizerStage, kwargs)
        bottleneck = (FactorizerStage, {"pos_embed": pos_embed, **kwargs})
        blocks = [plain] * (len(encoder_depth) - 1) + [bottleneck] + [plain] * len(decoder_depth)
        super().__init__(in_channels, out_channels, spatial_dims=nd, spatial_size=spatial_size,
                         encoder_depth=encoder_depth, encoder_width=encoder_width, strides=strides,
                         decoder_depth=decoder_depth, stem=stem, downsample=downsample, block=blocks,
                         upsample=upsample, head=head, num_deep_supr=num_deep_supr)
