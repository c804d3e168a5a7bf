"""Minimal link classes with Chainer's parameter naming (`W`, `b`; ChainList children named
by index) so that the parameter tree of the reference -- and therefore the snapshot key
prefixes generate.py:67-81 loads -- maps 1:1 onto torch `named_parameters()`
(`a.b.W` <-> `a/b/W`)."""
from __future__ import annotations

import math
from typing import Iterator, Tuple

import torch
from torch import nn

from . import functions as Fn


def _lecun_normal_(t: torch.Tensor) -> torch.Tensor:
    fan_in = t[0].numel()
    with torch.no_grad():
        return t.normal_(0.0, 1.0 / math.sqrt(fan_in))


class Convolution2D(nn.Module):
    """L.Convolution2D / L.DilatedConvolution2D restricted to (k,1) kernels on (B,C,T,1)
    tensors: W (out,in,k,1) ~ LeCunNormal, b = 0 (Chainer defaults)."""

    def __init__(self, in_channels, out_channels, ksize, stride=1, pad=0, dilate=1):
        super().__init__()
        k = ksize[0] if isinstance(ksize, (tuple, list)) else ksize
        self.stride = stride[0] if isinstance(stride, (tuple, list)) else stride
        self.pad = pad[0] if isinstance(pad, (tuple, list)) else pad
        self.dilate = dilate[0] if isinstance(dilate, (tuple, list)) else dilate
        self.W = nn.Parameter(_lecun_normal_(torch.empty(out_channels, in_channels, k, 1)))
        self.b = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, relu=False, out_len=None, relu_in=False):
        return Fn.conv(x, self.W, self.b, self.stride, self.pad, self.dilate, relu, out_len,
                       relu_in)


DilatedConvolution2D = Convolution2D


class EmbedID(nn.Module):
    """L.EmbedID: row gather, W (n, dim) ~ N(0, 1)."""

    def __init__(self, in_size, out_size):
        super().__init__()
        self.W = nn.Parameter(torch.randn(in_size, out_size))

    def forward(self, ids):
        return self.W[ids.long()]


def namedparams(module: nn.Module) -> Iterator[Tuple[str, nn.Parameter]]:
    """Chainer-style `/a/b/W` paths (Link.namedparams)."""
    for name, p in module.named_parameters():
        yield "/" + name.replace(".", "/"), p
