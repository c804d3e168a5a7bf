"""The training step (updaters.py:5-77 of the reference) and its optimiser.

`VQVAE_StandardUpdater.update_core` keeps the reference's ordering -- forward, cleargrads,
loss1.backward(), vq.cleargrads(), loss2.backward(), loss3.backward(), optimizer.update()
(updaters.py:13-19).  `VQVAE_ParallelUpdater` replaces the single-process gather-to-GPU-0 /
broadcast scheme (updaters.py:71-77) with one process per GPU and ONE summed all-reduce over
a flat gradient bucket; every rank then applies the identical Adam step, so no parameter
broadcast is needed.  Gradients are SUMMED, not averaged, and the caller divides the learning
rate by the number of replicas exactly like train.py:101.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import numpy
import torch
import torch.distributed as dist
from torch import nn


class GradBucket:
    """All gradients of a model as views into one contiguous fp32 buffer (plus an optional
    tail for the VQ per-code statistics), so `cleargrads` is one memset and the data-parallel
    reduction is one collective."""

    def __init__(self, model: nn.Module, extra: int = 0):
        self.params = [p for p in model.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n + extra, device=dev, dtype=torch.float32)
        self.n_grad = n
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.extra = self.flat[n:]

    def zero(self) -> None:
        self.flat.zero_()

    def allreduce(self, group=None) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


class Adam:
    """chainer.optimizers.Adam (train.py:101) [dep]:
        m += (1-b1)(g-m);  v += (1-b2)(g*g-v);
        p -= alpha*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)      (eps is not bias-corrected)."""

    def __init__(self, alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-8):
        self.alpha, self.beta1, self.beta2, self.eps = alpha, beta1, beta2, eps
        self.t = 0
        self.target: Optional[nn.Module] = None

    def setup(self, model: nn.Module, extra: int = 0) -> "Adam":
        self.target = model
        self.bucket = GradBucket(model, extra)
        self.params = self.bucket.params
        n = self.bucket.n_grad
        dev = self.params[0].device
        # parameters and both moments as views of flat buffers (same order as the gradient
        # bucket): on CUDA the whole update is ONE libvqw kernel over the flat range
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m, self.v = [], []
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)
                self.m.append(self.flat_m[off:off + k].view_as(p))
                self.v.append(self.flat_v[off:off + k].view_as(p))
                off += k
        return self

    @property
    def lr(self) -> float:
        t = max(self.t, 1)
        return self.alpha * math.sqrt(1 - self.beta2 ** t) / (1 - self.beta1 ** t)

    @torch.no_grad()
    def update(self) -> None:
        self.t += 1
        if self.flat_p.is_cuda:
            from . import _lib as L
            n = self.bucket.n_grad
            L.check(L.lib.vqw_adam_step(L.ptr(self.flat_p), L.ptr(self.bucket.flat), L.ptr(self.flat_m),
                                        L.ptr(self.flat_v), n, self.lr, self.beta1, self.beta2,
                                        self.eps, L.stream()), "vqw_adam_step")
            return
        grads = [p.grad for p in self.params]
        # m += (1-b1)(g-m)
        torch._foreach_lerp_(self.m, grads, 1 - self.beta1)
        g2 = torch._foreach_mul(grads, grads)
        torch._foreach_lerp_(self.v, g2, 1 - self.beta2)
        den = torch._foreach_sqrt(self.v)
        torch._foreach_add_(den, self.eps)
        torch._foreach_addcdiv_(self.params, self.m, den, value=-self.lr)


def cleargrads(model: nn.Module) -> None:
    for p in model.parameters():
        if p.grad is not None:
            p.grad.zero_()


def concat_examples(batch: Sequence[tuple], device) -> tuple:
    """chainer.dataset.concat_examples for Preprocess tuples (utils.py:100-110): stack each
    field, move to `device` through pinned memory."""
    fields = []
    for i in range(len(batch[0])):
        arr = numpy.stack([numpy.asarray(ex[i]) for ex in batch])
        t = torch.from_numpy(arr)
        if device is not None and torch.device(device).type == "cuda":
            t = t.pin_memory().to(device, non_blocking=True)
        fields.append(t)
    return tuple(fields)


class VQVAE_StandardUpdater:
    """updaters.py:5-19."""

    def __init__(self, iterator, optimizer: Adam, converter: Callable = concat_examples,
                 device=None, loss_func=None):
        self._iterators = {"main": iterator}
        self._optimizers = {"main": optimizer}
        self.converter = converter
        self.device = device
        self.loss_func = loss_func
        self.iteration = 0

    def backward_three(self, model, loss1, loss2, loss3) -> None:
        from . import functions as Fn
        model.zero_grad(set_to_none=False) if not hasattr(self._optimizers["main"], "bucket") \
            else self._optimizers["main"].bucket.zero()             # optimizer.target.cleargrads()
        prev, Fn.ACCUMULATE_INTO_GRAD = Fn.ACCUMULATE_INTO_GRAD, True
        try:
            loss1.backward(retain_graph=True)                       # :15
            cleargrads(model.vq)                                    # :16
            loss2.backward(retain_graph=True)                       # :17
            loss3.backward()                                        # :18
        finally:
            Fn.ACCUMULATE_INTO_GRAD = prev

    def update_core(self):
        batch = self._iterators["main"].next()
        in_arrays = self.converter(batch, self.device)
        optimizer = self._optimizers["main"]
        loss_func = self.loss_func or optimizer.target
        loss1, loss2, loss3 = loss_func(*in_arrays)                 # :13
        self.backward_three(optimizer.target, loss1, loss2, loss3)
        optimizer.update()                                          # :19
        return loss1, loss2, loss3

    def update(self):
        out = self.update_core()
        self.iteration += 1
        return out


class VQVAE_ParallelUpdater(VQVAE_StandardUpdater):
    """updaters.py:22-77, one process per GPU.  Rank r of n consumes `batch[r::n]`
    (updaters.py:36-38); the flat gradient bucket is all-reduced with SUM (addgrads,
    updaters.py:71-72); every rank applies the same Adam update (replaces copyparams,
    updaters.py:76-77)."""

    def __init__(self, iterator, optimizer: Adam, converter: Callable = concat_examples,
                 device=None, loss_func=None, group=None):
        super().__init__(iterator, optimizer, converter, device, loss_func)
        self.group = group

    @staticmethod
    def split(batch: Sequence, rank: int, n: int):
        return batch[rank::n]

    def update_core(self):
        n = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        batch = self._iterators["main"].next()
        in_arrays = self.converter(self.split(batch, rank, n), self.device)
        optimizer = self._optimizers["main"]
        loss_func = self.loss_func or optimizer.target
        loss1, loss2, loss3 = loss_func(*in_arrays)
        self.backward_three(optimizer.target, loss1, loss2, loss3)
        optimizer.bucket.allreduce(self.group)
        optimizer.update()
        return loss1, loss2, loss3
