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

    def offset_of(self, grad: torch.Tensor) -> int:
        """Element offset of a gradient view inside the flat bucket."""
        return (grad.data_ptr() - self.flat.data_ptr()) // self.flat.element_size()


def block_segments(n_blocks: int, n_segments: int):
    """Split blocks 0..n-1 into <= n_segments contiguous groups, returned in the order the
    backward finishes them (highest blocks first): [(first_block, last_block_exclusive), ...]."""
    n_segments = max(1, min(n_segments, n_blocks))
    bounds = [round(i * n_blocks / n_segments) for i in range(n_segments + 1)]
    segs = [(bounds[i], bounds[i + 1]) for i in range(n_segments) if bounds[i + 1] > bounds[i]]
    return segs[::-1]


class _OverlappedReduce:
    """Gradient all-reduce of the residual stack overlapped with its own backward.  The stack
    backward walks the blocks from the last to the first and records an event per block
    (vqw_resnet_desc.block_events); the bucket range of every finished group of blocks is
    all-reduced on a side stream while the remaining blocks still compute.  What is left of the
    bucket (encoder, codebook, ConditionEmbed, embed and head parameters) is reduced after the
    last backward pass.  Same sum, same result as one all-reduce over the whole bucket."""

    def __init__(self, bucket: GradBucket, group, n_segments: int = 4):
        self.bucket, self.group, self.n_segments = bucket, group, n_segments
        self.side = torch.cuda.Stream()
        self.pending = []
        self.done_ranges = []
        self._events = None

    def events(self, n_blocks):
        if self._events is None or len(self._events) != n_blocks:
            self._events = [torch.cuda.Event() for _ in range(n_blocks)]
            for e in self._events:
                e.record()          # materialises the CUDA event handle
        return self._events

    def launched(self, params, events):
        n = len(events)
        per = len(params) // n
        flat = self.bucket.flat
        for a, b in block_segments(n, self.n_segments):
            lo = self.bucket.offset_of(params[a * per].grad)
            last = params[b * per - 1].grad
            hi = self.bucket.offset_of(last) + last.numel()
            self.side.wait_event(events[a])          # block a is the last of the group to finish
            with torch.cuda.stream(self.side):
                self.pending.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM,
                                                    group=self.group, async_op=True))
            self.done_ranges.append((lo, hi))

    def finish(self):
        """Reduce whatever `launched` did not cover and join the side stream."""
        flat = self.bucket.flat
        ranges = sorted(self.done_ranges)
        pos = 0
        for lo, hi in ranges + [(flat.numel(), flat.numel())]:
            if lo > pos:
                dist.all_reduce(flat[pos:lo], op=dist.ReduceOp.SUM, group=self.group)
            pos = max(pos, hi)
        for w in self.pending:
            w.wait()
        self.pending, self.done_ranges = [], []


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
        self._update_torch()

    def update_captured(self, lr_dev: torch.Tensor) -> None:
        """The same update with the learning rate read from `lr_dev` (a 1-element CUDA tensor)
        and WITHOUT advancing `t`: the form recorded into a CUDA graph; the caller advances `t`
        and refreshes `lr_dev` before every replay."""
        from . import _lib as L
        L.check(L.lib.vqw_adam_step_dev(L.ptr(self.flat_p), L.ptr(self.bucket.flat), L.ptr(self.flat_m),
                                        L.ptr(self.flat_v), self.bucket.n_grad, L.ptr(lr_dev),
                                        self.beta1, self.beta2, self.eps, L.stream()),
                "vqw_adam_step_dev")

    @torch.no_grad()
    def _update_torch(self) -> None:
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
                 device=None, loss_func=None, use_cuda_graph: bool = False, graph_warmup: int = 3):
        self._iterators = {"main": iterator}
        self._optimizers = {"main": optimizer}
        self.converter = converter
        self.device = device
        self.loss_func = loss_func
        self.iteration = 0
        # use_cuda_graph: after `graph_warmup` ordinary steps the whole step (forward, three
        # backward passes, gradient reduction, Adam, weight EMA: ~500 kernel launches) is
        # captured once into a CUDA graph and replayed from then on; batches are copied into the
        # graph's static input buffers.  Same kernels, same order, same results.
        self.use_cuda_graph = use_cuda_graph
        self.graph_warmup = graph_warmup
        self._graph = None
        self._eager_steps = 0

    MERGE_ENCODER_BACKWARD = True

    def backward_three(self, model, loss1, loss2, loss3) -> None:
        from . import functions as Fn
        model.zero_grad(set_to_none=False) if not hasattr(self._optimizers["main"], "bucket") \
            else self._optimizers["main"].bucket.zero()             # optimizer.target.cleargrads()
        prev, Fn.ACCUMULATE_INTO_GRAD = Fn.ACCUMULATE_INTO_GRAD, True
        try:
            # :15-16 -- the codebook gradient of loss1 is cleared right after it is computed, so
            # it is not computed (cleargrads above already left vq.W.grad at zero).
            # :15 + :18 in one pass: loss1 and loss3 meet at the encoder output z, gradients
            # accumulate (no cleargrads between :15 and :18 touches the encoder), and backward is
            # linear -- so the encoder's backward runs ONCE on d(loss1)/dz + d(loss3)/dz instead of
            # twice.  loss3 reaches nothing but the encoder (e is detached, net.py:91), so the
            # codebook still only sees loss2 (:16-17).  MERGE_ENCODER_BACKWARD = False restores
            # the literal three calls.
            Fn.DISCARD_CODEBOOK_GRAD = True
            Fn.STACK_BACKWARD_OBSERVER = self._overlap()     # decoder grads come from loss1 only
            merged = self.MERGE_ENCODER_BACKWARD and loss3.requires_grad and loss1.requires_grad
            try:
                if merged:
                    torch.autograd.backward((loss1, loss3), retain_graph=True)   # :15, :18
                else:
                    loss1.backward(retain_graph=True)               # :15
            finally:
                Fn.DISCARD_CODEBOOK_GRAD = False
                Fn.STACK_BACKWARD_OBSERVER = None
            cleargrads(model.vq)                                    # :16
            loss2.backward(retain_graph=not merged)                 # :17
            if not merged:
                loss3.backward()                                    # :18
        finally:
            Fn.ACCUMULATE_INTO_GRAD = prev

    def _reduce(self, optimizer) -> None:
        """gradient exchange between replicas (none for the single-device updater)"""

    def _overlap(self):
        """observer of the stack backward (see functions.STACK_BACKWARD_OBSERVER) or None"""
        return None

    def _step(self, in_arrays, captured_lr=None):
        optimizer = self._optimizers["main"]
        loss_func = self.loss_func or optimizer.target
        loss1, loss2, loss3 = loss_func(*in_arrays)                 # :13
        self.backward_three(optimizer.target, loss1, loss2, loss3)  # :14-18
        self._reduce(optimizer)
        if captured_lr is None:
            optimizer.update()                                      # :19
        else:
            optimizer.update_captured(captured_lr)
        # detached: nothing may keep this step's autograd graph (and its AccumulateGrad nodes,
        # which are bound to the stream they were created on) alive into a later graph capture
        return loss1.detach(), loss2.detach(), loss3.detach()

    def update_from_arrays(self, in_arrays):
        """One step on already converted (device-resident) arrays."""
        if not (self.use_cuda_graph and in_arrays[0].is_cuda):
            return self._step(in_arrays)
        optimizer = self._optimizers["main"]
        sig = tuple((tuple(t.shape), t.dtype) for t in in_arrays)
        if self._graph is not None and sig != self._graph_sig:
            self._graph = None                                      # new batch shape: re-capture
        if self._graph is None:
            if self._eager_steps < self.graph_warmup:
                self._eager_steps += 1
                return self._step(in_arrays)
            self._static_in = tuple(t.clone() for t in in_arrays)
            self._lr_dev = torch.zeros(1, device=in_arrays[0].device, dtype=torch.float32)
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            try:
                with torch.cuda.graph(graph):
                    self._static_out = self._step(self._static_in, captured_lr=self._lr_dev)
            except Exception as exc:                                # capture unsupported here:
                import sys                                          # stay on the ordinary path
                print(f"[vqw] CUDA-graph capture of the training step failed ({exc!r}); "
                      "continuing without it", file=sys.stderr)
                self.use_cuda_graph = False
                torch.cuda.synchronize()
                return self._step(in_arrays)
            self._graph, self._graph_sig = graph, sig
        else:
            for dst, src in zip(self._static_in, in_arrays):
                dst.copy_(src, non_blocking=True)
        optimizer.t += 1
        self._lr_dev.fill_(optimizer.lr)
        self._graph.replay()
        return tuple(o.detach().clone() for o in self._static_out)

    def release_graph(self) -> None:
        """Drop the captured step (and its static buffers).  Call it before tearing down a
        process group whose collectives were captured: destroying the NCCL communicator while a
        live graph still references its kernels hangs."""
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph = None
            self._static_in = self._static_out = None
            torch.cuda.synchronize()

    def update_core(self):
        batch = self._iterators["main"].next()
        in_arrays = self.converter(batch, self.device)
        return self.update_from_arrays(in_arrays)

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
                 device=None, loss_func=None, group=None, use_cuda_graph: bool = False,
                 graph_warmup: int = 3):
        super().__init__(iterator, optimizer, converter, device, loss_func, use_cuda_graph,
                         graph_warmup)
        self.group = group
        self._synced = False
        # all-reduce of the stack's gradients under the stack's own backward (tensor-core modes)
        import os
        self.overlap_allreduce = os.environ.get("VQW_OVERLAP_ALLREDUCE", "1") != "0"
        self.n_reduce_segments = 4
        self._ovl = None

    def sync_replicas(self) -> None:
        """The reference re-broadcasts the main model's parameters after every step
        (`copyparams`, updaters.py:76-77).  Here every rank applies the identical Adam update to
        the identical all-reduced gradient, so ONE broadcast from rank 0 -- parameters, both Adam
        moments, the step count and the decoder's EMA copy -- before the first step is
        equivalent and makes the replicas independent of per-process seeds / resume state."""
        self._synced = True
        if not (dist.is_available() and dist.is_initialized()) or \
                dist.get_world_size(self.group) == 1:
            return
        opt = self._optimizers["main"]
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        bufs = [getattr(opt, n) for n in ("flat_p", "flat_m", "flat_v") if hasattr(opt, n)]
        if not bufs:                                   # an optimiser without flat buffers
            bufs = [p.data for p in opt.target.parameters()]
        seen = {b.data_ptr() for b in bufs}
        for m in opt.target.modules():                 # EMA copies are not optimiser parameters
            ema = getattr(m, "ema", None)
            if isinstance(ema, nn.Module):
                bufs += [p.data for p in ema.parameters() if p.data_ptr() not in seen]
        for b in bufs:
            dist.broadcast(b, src=src, group=self.group)
        t = torch.tensor([opt.t], device=bufs[0].device, dtype=torch.int64)
        dist.broadcast(t, src=src, group=self.group)
        opt.t = int(t)

    @staticmethod
    def split(batch: Sequence, rank: int, n: int):
        return batch[rank::n]

    def update_core(self):
        n = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        batch = self._iterators["main"].next()
        in_arrays = self.converter(self.split(batch, rank, n), self.device)
        return self.update_from_arrays(in_arrays)

    def update_from_arrays(self, in_arrays):
        if not self._synced:
            self.sync_replicas()
        return super().update_from_arrays(in_arrays)

    def _overlap(self):
        """Overlap the all-reduce of the residual stack's gradients with the stack's own backward
        (NCCL, more than one rank, not while a CUDA graph is being captured or replayed)."""
        if not self.overlap_allreduce or self.use_cuda_graph:
            return None
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) < 2:
            return None
        opt = self._optimizers["main"]
        if not opt.bucket.flat.is_cuda:
            return None
        if self._ovl is None:
            self._ovl = _OverlappedReduce(opt.bucket, self.group, self.n_reduce_segments)
        return self._ovl

    def _reduce(self, optimizer) -> None:
        if self._ovl is not None and (self._ovl.pending or self._ovl.done_ranges):
            self._ovl.finish()                                      # addgrads, :71-72, in pieces
        else:
            optimizer.bucket.allreduce(self.group)                  # addgrads, :71-72
