"""VQ link, straight-through op, decoder weight-EMA wrapper and mu-law companding behind the
reference's names (utils.py:12-29, 131-255)."""
from __future__ import annotations

import copy
import math

import numpy
import torch
from torch import nn

from . import functions as Fn
from .functions import straight_through  # noqa: F401  (utils.py:234-236)


class MuLaw(object):
    """utils.py:12-29 -- mu = 256 in both the companding law and the bin count, and the
    inverse uses mu**|y| (reference quirks).  Host-side (NumPy):
    it runs in the data workers of the reference, not on the device."""

    def __init__(self, mu=256, int_type=numpy.int32, float_type=numpy.float32):
        self.mu = mu
        self.int_type = int_type
        self.float_type = float_type

    def transform(self, x):
        x = x.astype(self.float_type)
        y = numpy.sign(x) * numpy.log(1 + self.mu * numpy.abs(x)) / numpy.log(1 + self.mu)
        edges = 2 * numpy.arange(self.mu) / self.mu - 1
        return (numpy.digitize(y, edges) - 1).astype(self.int_type)

    def itransform(self, y):
        y = y.astype(self.float_type)
        y = 2 * y / self.mu - 1
        x = numpy.sign(y) / self.mu * ((self.mu) ** numpy.abs(y) - 1)
        return x.astype(self.float_type)


class VQ(nn.Module):
    """utils.py:239-255: owns the codebook W (k, d); `__call__(x)` = straight_through(x, W).
    `initialW=None` is Chainer's default initialiser (LeCunNormal, std 1/sqrt(d));
    `d=None` defers allocation to the first call (utils.py:253-254)."""

    def __init__(self, k, d=None, initialW=None):
        super().__init__()
        self.k = k
        self._initialW = initialW
        self.W = nn.UninitializedParameter()
        self.indexes = None
        if d is not None:
            self._initialize_params(d)

    def _initialize_params(self, d, device=None):
        w = torch.empty(self.k, d, device=device)
        if self._initialW is None:
            w.normal_(0.0, 1.0 / math.sqrt(d))
        elif callable(self._initialW):
            self._initialW(w)
        else:
            w.copy_(torch.as_tensor(self._initialW))
        if isinstance(self.W, nn.UninitializedParameter):
            self.W.materialize((self.k, d), device=device)
        with torch.no_grad():
            self.W.copy_(w)

    def forward(self, x, cached=None):
        if isinstance(self.W, nn.UninitializedParameter):
            self._initialize_params(x.shape[1], device=x.device)
        e, idx = Fn.straight_through(x, self.W, cached, return_indexes=True)
        self.indexes = idx        # int32, immutable (the reference overwrites it in backward,
        return e                  # utils.py:227)


class ExponentialMovingAverage(nn.Module):
    """utils.py:131-158: Polyak copy of the DECODER weights, refreshed inside every
    training-mode forward as ema = decay*target + (1-decay)*ema  (decay multiplies the
    target: reference quirk, utils.py:153-154); evaluation runs the `ema` copy."""

    def __init__(self, target, decay=0.999):
        super().__init__()
        self.decay = decay
        self.target = target
        self.ema = copy.deepcopy(target)
        for p in self.ema.parameters():
            p.requires_grad_(False)

    def set_mode(self, mode) -> None:
        """Arithmetic mode of BOTH copies (the evaluation copy is a deepcopy: it does not see a
        later set_mode on the target)."""
        for m in (self.target, self.ema):
            if hasattr(m, "set_mode"):
                m.set_mode(mode)

    def run(self, method, *args, **kwargs):
        """`forward` for another entry point of the wrapped link (e.g. WaveNet.forward_loss):
        training calls the target and refreshes the average, evaluation calls the copy."""
        if self.training:
            ys = getattr(self.target, method)(*args, **kwargs)
            self.update_average()
        else:
            ys = getattr(self.ema, method)(*args, **kwargs)
        return ys

    def forward(self, *args, **kwargs):
        if self.training:                                    # configuration.config.train
            ys = self.target(*args, **kwargs)
            self.update_average()
        else:
            ys = self.ema(*args, **kwargs)
        return ys

    @staticmethod
    def _flat_range(params):
        """(first tensor, numel) if the parameters are adjacent views of one buffer, else None."""
        if not params or not params[0].is_cuda:
            return None
        ptr, total = params[0].data_ptr(), 0
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.data_ptr() != ptr + 4 * total:
                return None
            total += p.numel()
        return params[0], total

    @torch.no_grad()
    def update_average(self):
        tp = [p for p in self.target.parameters()]
        ep = [p for p in self.ema.parameters()]
        # same pairing as the reference's name match (utils.py:146-148), without the O(P^2) loop
        if tp and tp[0].is_cuda and self._flat_range(ep) is None:
            flat = torch.cat([p.detach().reshape(-1) for p in ep])
            off = 0
            for p in ep:
                p.data = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self._ema_flat = flat
        rt, re = self._flat_range(tp), self._flat_range(ep)
        if rt is not None and re is not None and rt[1] == re[1]:
            from . import _lib as L          # one kernel over the flat ranges (Adam.setup made
            L.check(L.lib.vqw_ema_update(re[0].data_ptr(), rt[0].data_ptr(), rt[1],   # them flat)
                                         float(self.decay), L.stream()), "vqw_ema_update")
            return
        torch._foreach_mul_(ep, 1.0 - self.decay)
        torch._foreach_add_(ep, tp, alpha=self.decay)
