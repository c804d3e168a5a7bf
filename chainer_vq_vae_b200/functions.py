"""torch.autograd.Function wrappers over the C-ABI kernels (include/vqw.h).

Tensors keep the reference layout (B, C, T, 1) float32.  Every function launches on the
current CUDA stream and raises on non-CUDA input: there is no eager/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib as L


def _as3(x: torch.Tensor):
    """(B,C,T,1) or (B,C,T) -> sizes; the memory is identical."""
    if x.dim() == 4:
        if x.shape[3] != 1:
            raise ValueError("the last (width) axis must be 1, as in the reference's (k,1) convs")
    elif x.dim() != 3:
        raise ValueError(f"expected a (B,C,T,1) tensor, got shape {tuple(x.shape)}")
    return x.shape[0], x.shape[1], x.shape[2]


def _f32c(x: torch.Tensor) -> torch.Tensor:
    if x.dtype != torch.float32:
        raise TypeError(f"float32 expected, got {x.dtype}")
    return x.contiguous()


# ---------------------------------------------------------------------------------------
# generic strided / dilated convolution with (k,1) kernels
# ---------------------------------------------------------------------------------------
def conv_out_len(L_in: int, k: int, stride: int, pad: int, dilate: int) -> int:
    """Chainer's get_conv_outsize: (L + 2p - dil*(k-1) - 1)//s + 1."""
    return (L_in + 2 * pad - dilate * (k - 1) - 1) // stride + 1


def _conv_launch(desc: L.ConvDesc, out: torch.Tensor, what: str) -> None:
    L.check(L.lib.vqw_conv_forward(C.byref(desc), L.ptr(out), L.stream()), what)


def _fill_src(s: L.ConvSrc, inp, w_ptr, K, Tin, wm, wk, mul, shift, div, relu_in=0, in_mask=None):
    s.in_ = L.ptr(inp)
    s.w = w_ptr
    s.in_mask = L.ptr(in_mask)
    s.K, s.Tin, s.wm, s.wk = K, Tin, wm, wk
    s.mul, s.shift, s.div, s.relu_in = mul, shift, div, relu_in


class _Conv(torch.autograd.Function):
    """y = [relu](conv(x, W) + b) over the time axis; cross-correlation, symmetric zero pad."""

    @staticmethod
    def forward(ctx, x, W, b, stride, pad, dilate, relu_out, out_len, relu_in=False):
        x, W = _f32c(x), _f32c(W)
        B, Cin, Tin = _as3(x)
        Cout, Cin_w, kh = W.shape[0], W.shape[1], W.shape[2]
        if Cin_w != Cin:
            raise ValueError(f"conv: input has {Cin} channels, weight expects {Cin_w}")
        Tout = conv_out_len(Tin, kh, stride, pad, dilate)
        if out_len is not None:
            Tout = min(Tout, out_len)        # the reference slices [:, :, :length]
        out = torch.empty((B, Cout, max(Tout, 0), 1), device=x.device, dtype=torch.float32)
        if b is not None:
            b = _f32c(b)
        wp = L.ptr(W)
        for j0 in range(0, kh, L.VQW_MAX_SRC):
            d = L.ConvDesc()
            d.B, d.M, d.T = B, Cout, Tout
            taps = range(j0, min(kh, j0 + L.VQW_MAX_SRC))
            d.nsrc = len(taps)
            for n, j in enumerate(taps):
                _fill_src(d.src[n], x, wp + 4 * j, Cin, Tin, Cin * kh, kh, stride,
                          j * dilate - pad, 1, int(relu_in))
            last = j0 + L.VQW_MAX_SRC >= kh
            d.bias = L.ptr(b) if (b is not None and j0 == 0) else None
            d.accumulate = 1 if j0 > 0 else 0
            d.relu_out = 1 if (relu_out and last) else 0
            if relu_out and not last:
                raise NotImplementedError("relu fusion needs filter_size <= 4")
            _conv_launch(d, out, "vqw_conv_forward")
        ctx.cfg = (stride, pad, dilate, relu_out, B, Cin, Tin, Cout, kh, Tout, b is not None,
                   relu_in)
        ctx.save_for_backward(x, W, out if relu_out else None)
        return out

    @staticmethod
    def backward(ctx, gy):
        stride, pad, dilate, relu_out, B, Cin, Tin, Cout, kh, Tout, has_b, relu_in = ctx.cfg
        x, W, y = ctx.saved_tensors
        gy = _f32c(gy)
        gx = gW = gb = None
        wp = L.ptr(W)
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            for j0 in range(0, kh, L.VQW_MAX_SRC):
                d = L.ConvDesc()
                d.B, d.M, d.T = B, Cin, Tin
                taps = range(j0, min(kh, j0 + L.VQW_MAX_SRC))
                d.nsrc = len(taps)
                for n, j in enumerate(taps):
                    _fill_src(d.src[n], gy, wp + 4 * j, Cout, Tout, kh, Cin * kh, 1,
                              pad - j * dilate, stride, 0, y if relu_out else None)
                d.accumulate = 1 if j0 > 0 else 0
                if relu_in:
                    if j0 + L.VQW_MAX_SRC < kh:
                        raise NotImplementedError("relu_in fusion needs filter_size <= 4")
                    d.out_mask = L.ptr(x)        # relu'(x) = (x > 0)
                _conv_launch(d, gx, "vqw_conv_forward(dgrad)")
        if ctx.needs_input_grad[1] or (has_b and ctx.needs_input_grad[2]):
            gW = torch.zeros_like(W)
            gb = torch.zeros(Cout, device=x.device, dtype=torch.float32) if has_b else None
            g = L.WgradDesc()                      # all kh taps in one launch
            g.B, g.M, g.T = B, Cout, Tout
            g.a, g.a_mask = L.ptr(gy), (L.ptr(y) if relu_out else None)
            g.in_, g.in_mul = L.ptr(x), None
            g.K, g.Tin = Cin, Tin
            g.mul, g.shift, g.div, g.relu_in = stride, -pad, 1, int(relu_in)
            g.gm, g.gk = Cin * kh, kh
            g.ntaps, g.tap_dshift, g.tap_gw = kh, dilate, 1
            L.check(L.lib.vqw_conv_wgrad(C.byref(g), L.ptr(gW), L.ptr(gb) if has_b else None,
                                         L.stream()), "vqw_conv_wgrad")
        return gx, gW, gb, None, None, None, None, None, None


def conv(x, W, b=None, stride=1, pad=0, dilate=1, relu=False, out_len=None, relu_in=False):
    """Chainer `L.Convolution2D` / `L.DilatedConvolution2D` arithmetic for (k,1) kernels;
    `relu_in` / `relu` fuse an F.relu before / after the convolution."""
    return _Conv.apply(x, W, b, stride, pad, dilate, relu, out_len, relu_in)


# ---------------------------------------------------------------------------------------
# VQ straight-through (utils.py:161-236)
# ---------------------------------------------------------------------------------------
def check_type_forward(x: torch.Tensor, W: torch.Tensor) -> None:
    """StraightThrough.check_type_forward, utils.py:162-174 (InvalidType -> TypeError/ValueError)."""
    if not (x.is_floating_point() and W.is_floating_point()):
        raise TypeError("straight_through: x and W must be floating point (utils.py:168-169)")
    if not (3 <= x.dim() <= 4):
        raise ValueError("straight_through: x.ndim must be 3 or 4 (utils.py:170-171)")
    if W.dim() != 2:
        raise ValueError("straight_through: W.ndim must be 2 (utils.py:172)")
    if x.shape[1] != W.shape[1]:
        raise ValueError("straight_through: x.shape[1] != W.shape[1] (utils.py:173)")
    if x.device != W.device:
        raise ValueError("straight_through: x and W must live on the same device "
                         "(numpy and cupy must not be used together, utils.py:183-186)")


def vq_lookup(x: torch.Tensor, W: torch.Tensor, stats: bool = False):
    """One launch of the fused distance + argmin + gather kernel.
    Returns (e, indexes int32 shaped like x without the channel axis, count, zsum, sqerr)."""
    check_type_forward(x, W)
    x, W = _f32c(x), _f32c(W)
    B, d = x.shape[0], x.shape[1]
    T = x.shape[2] if x.dim() == 3 else x.shape[2] * x.shape[3]
    k = W.shape[0]
    idx = torch.empty((B,) + tuple(x.shape[2:]), device=x.device, dtype=torch.int32)
    e = torch.empty_like(x)
    count = zsum = sqerr = None
    if stats:
        count = torch.zeros(k, device=x.device, dtype=torch.float32)
        zsum = torch.zeros(k, d, device=x.device, dtype=torch.float32)
        sqerr = torch.zeros(1, device=x.device, dtype=torch.float64)
    L.check(L.lib.vqw_vq_forward(L.ptr(x), L.ptr(W), L.ptr(idx), L.ptr(e), L.ptr(count),
                                 L.ptr(zsum), L.ptr(sqerr), B, d, T, k, L.stream()),
            "vqw_vq_forward")
    return e, idx, count, zsum, sqerr


# Set by the VQ-VAE updaters around loss1.backward(): the reference computes d loss1 / d W of
# the codebook and discards it at once (`model.vq.cleargrads()`, updaters.py:15-16); with the
# flag on, the backward simply does not produce it (same state afterwards, one kernel less).
DISCARD_CODEBOOK_GRAD = False


class _StraightThrough(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, cached):
        if cached is None:
            e, idx, _, _, _ = vq_lookup(x, W)
        else:
            e, idx = cached
        ctx.save_for_backward(idx)
        ctx.wshape = tuple(W.shape)
        ctx.mark_non_differentiable(idx)
        return e, idx

    @staticmethod
    def backward(ctx, gy, _gidx):
        (idx,) = ctx.saved_tensors
        gx = gy if ctx.needs_input_grad[0] else None          # utils.py:218-219
        gW = None
        if ctx.needs_input_grad[1] and not DISCARD_CODEBOOK_GRAD:   # utils.py:220-230
            gyc = _f32c(gy)
            k, d = ctx.wshape
            B = gyc.shape[0]
            T = gyc.numel() // (B * d)
            gW = torch.empty(k, d, device=gy.device, dtype=torch.float32)
            L.check(L.lib.vqw_vq_backward_w(L.ptr(gyc), L.ptr(idx), L.ptr(gW), B, d, T, k,
                                            L.stream()), "vqw_vq_backward_w")
        return gx, gW, None


def straight_through(x, W, cached=None, return_indexes=False):
    """utils.py:234-236.  `cached=(e, idx)` reuses a previous lookup of the same (x, W) values
    (VAE.__call__ quantises z twice, net.py:82-83; the kernel runs once)."""
    e, idx = _StraightThrough.apply(x, W, cached)
    return (e, idx) if return_indexes else e


# ---------------------------------------------------------------------------------------
# fused residual block stack (modules.py:30-56, 89-96)
# ---------------------------------------------------------------------------------------
def _rb_desc(B, T, Cr, Cd, Cs, Cc, fs, dil, accumulate, write_res, mode) -> L.ResblockDesc:
    d = L.ResblockDesc()
    d.B, d.T, d.Cr, d.Cd, d.Cs, d.Cc = B, T, Cr, Cd, Cs, Cc
    d.fs, d.dilation = fs, dil
    d.skip_accumulate, d.write_residual, d.mode = int(accumulate), int(write_res), mode
    return d


def _rb_weights(ws: Sequence[Optional[torch.Tensor]]) -> L.ResblockWeights:
    w = L.ResblockWeights()
    for name, t in zip(("conv_w", "conv_b", "cond_w", "cond_b", "res_w", "res_b", "skip_w",
                        "skip_b"), ws):
        setattr(w, name, L.ptr(t))
    return w


# When True (set by the VQ-VAE updaters around their backward passes) the residual stack's
# backward accumulates weight gradients straight into the parameters' existing `.grad` buffers
# (the views of the flat gradient bucket) instead of materialising 8*n_blocks zero-filled
# tensors that autograd then adds -- same result, ~300 fewer tiny kernels per step.
ACCUMULATE_INTO_GRAD = False

# Data-parallel overlap (set by VQVAE_ParallelUpdater around loss1.backward()): an object with
#   .events(n_blocks) -> list of torch.cuda.Event (or None) to be recorded by the library when the
#                        gradients of block i are final, and
#   .launched(weights) called right after the backward kernels of the stack have been ENQUEUED,
# so that the gradient all-reduce of the blocks already finished runs on a side stream under the
# rest of the backward (vqw_resnet_desc.block_events).
STACK_BACKWARD_OBSERVER = None


class _ResidualStack(torch.autograd.Function):
    """All blocks of a ResidualNet in one autograd node: forward launches one fused kernel
    per block (skip accumulated in place, modules.py:92-95); backward walks the blocks in
    reverse with the shared g_skip and the accumulated g_condition."""

    @staticmethod
    def forward(ctx, x, cond, cond_global, dilations, fs, mode, keep_last_residual, grad_targets,
                grad_enabled, *weights):
        ctx.grad_targets = grad_targets
        x, cond = _f32c(x), _f32c(cond)
        B, Cr, T = _as3(x)
        Bc, Cc, Tc = _as3(cond)
        if (Bc, Tc) != (B, T):
            raise ValueError(f"condition shape {tuple(cond.shape)} does not match x {tuple(x.shape)}")
        Cg = 0
        if cond_global is not None:      # hoisted time-constant condition channels (B, Cg)
            cond_global = _f32c(cond_global)
            if cond_global.dim() != 2 or cond_global.shape[0] != B:
                raise ValueError("cond_global must be (B, Cg)")
            Cg = cond_global.shape[1]
            Cc += Cg                     # Cc = columns of condition_proj.W
        n = len(dilations)
        assert len(weights) == 8 * n
        weights = [_f32c(w) for w in weights]
        Cd = weights[0].shape[0]
        Cs = weights[6].shape[0]
        skip = torch.empty((B, Cs, T, 1), device=x.device, dtype=torch.float32)
        # needs_input_grad mirrors requires_grad, NOT the grad mode (inside forward() grad mode
        # is always off): inference under no_grad must not take the save-for-backward path
        need_grad = grad_enabled and any(ctx.needs_input_grad)
        tc_mode = mode != L.MODE_FP32

        def new(ch):
            return torch.empty((B, ch, T, 1), device=x.device, dtype=torch.float32)

        # residual[i] = output of block i = saved input of block i+1.  Training keeps all of
        # them; inference chains the blocks through the library's workspace.
        if need_grad and not tc_mode:
            res = [new(Cr) for _ in range(n - 1)] + [new(Cr) if keep_last_residual else None]
        else:
            res = [None] * (n - 1) + [new(Cr) if keep_last_residual else None]
        gates: List[torch.Tensor] = []
        if need_grad:
            # bf16x3 keeps only the sigmoid (tanh = z / sigmoid from the saved z planes)
            only_sig = mode in L.X3_MODES
            for _ in range(n):
                gates += [x.new_empty(0) if only_sig else new(Cd // 2), new(Cd // 2)]

        d = L.ResnetDesc()
        d.B, d.T, d.Cr, d.Cd, d.Cs, d.Cc, d.fs = B, T, Cr, Cd, Cs, Cc, fs
        d.n_blocks = n
        dil_arr = (C.c_int * n)(*dilations)
        d.dilations = C.cast(dil_arr, C.POINTER(C.c_int))
        d.mode, d.keep_last_residual = mode, int(keep_last_residual)
        d.Cg, d.cond_global = Cg, L.ptr(cond_global)
        if weights[2].shape[1] != Cc:
            raise ValueError(f"condition_proj expects {weights[2].shape[1]} channels, got {Cc}")
        warr = (L.ResblockWeights * n)()
        for i in range(n):
            for name, t in zip(("conv_w", "conv_b", "cond_w", "cond_b", "res_w", "res_b",
                                "skip_w", "skip_b"), weights[8 * i:8 * i + 8]):
                setattr(warr[i], name, L.ptr(t))
        rarr = (C.c_void_p * n)(*[L.ptr(r) for r in res])
        garr_t = garr_s = None
        if gates:
            garr_t = (C.c_void_p * n)(*[(L.ptr(gates[2 * i]) if gates[2 * i].numel() else None)
                                        for i in range(n)])
            garr_s = (C.c_void_p * n)(*[L.ptr(gates[2 * i + 1]) for i in range(n)])
        ws_bytes = L.lib.vqw_resnet_forward_workspace(C.byref(d))
        workspace = torch.empty(max(int(ws_bytes), 1), device=x.device, dtype=torch.uint8)
        # tensor-core training: the library keeps the bf16 hi/lo planes the backward consumes
        # (condition, every block's input and gated activation) in this caller-owned buffer
        saved = None
        if need_grad and tc_mode:
            saved = torch.empty(int(L.lib.vqw_resnet_saved_bytes(C.byref(d))), device=x.device,
                                dtype=torch.uint8)
        if tc_mode:
            L.probe_forward_kernels()
        with L.timed("resnet_forward"):
            L.check(L.lib.vqw_resnet_forward(C.byref(d), L.ptr(x), L.ptr(cond), warr, rarr,
                                             L.ptr(skip), garr_t, garr_s, L.ptr(workspace),
                                             L.ptr(saved), L.stream()), "vqw_resnet_forward")
        xs: List[torch.Tensor] = [x] + [r for r in res[:n - 1]]
        ctx.tc_saved = saved
        if saved is not None:
            xs = [x] + [None] * (n - 1)
        residual = res[n - 1]
        ctx.cfg = (tuple(dilations), fs, mode, B, T, Cr, Cd, Cs, Cc, keep_last_residual, Cg)
        if need_grad:
            keep = [t if t is not None else x.new_empty(0) for t in xs]
            cg = cond_global if cond_global is not None else x.new_empty(0)
            ctx.save_for_backward(cond, cg, *keep, *gates, *weights)
        if keep_last_residual:
            return skip, residual
        return skip

    @staticmethod
    def backward(ctx, g_skip, g_last_res=None):
        dilations, fs, mode, B, T, Cr, Cd, Cs, Cc, keep_last, Cg = ctx.cfg
        n = len(dilations)
        saved = ctx.saved_tensors
        cond, cond_global = saved[0], saved[1]
        xs = saved[2:2 + n]
        tc_saved = getattr(ctx, "tc_saved", None)
        gates = saved[2 + n:2 + 3 * n]
        weights = saved[2 + 3 * n:]
        g_skip = _f32c(g_skip)
        dev = g_skip.device
        gcond = torch.zeros((B, Cc - Cg, T, 1), device=dev, dtype=torch.float32)
        g_glob = torch.zeros((B, Cg), device=dev, dtype=torch.float32) if Cg else None
        targets = ctx.grad_targets
        direct = (ACCUMULATE_INTO_GRAD and targets is not None and
                  all(p.grad is not None and p.grad.is_contiguous() for p in targets))
        gws = [p.grad for p in targets] if direct else [torch.zeros_like(w) for w in weights]
        d = L.ResnetDesc()
        d.B, d.T, d.Cr, d.Cd, d.Cs, d.Cc, d.fs = B, T, Cr, Cd, Cs, Cc, fs
        d.n_blocks = n
        dil_arr = (C.c_int * n)(*dilations)
        d.dilations = C.cast(dil_arr, C.POINTER(C.c_int))
        d.mode, d.keep_last_residual = mode, int(keep_last)
        d.Cg, d.cond_global, d.g_cond_global = Cg, (L.ptr(cond_global) if Cg else None), L.ptr(g_glob)
        warr = (L.ResblockWeights * n)()
        gwarr = (L.ResblockWeights * n)()
        names = ("conv_w", "conv_b", "cond_w", "cond_b", "res_w", "res_b", "skip_w", "skip_b")
        for i in range(n):
            for j, name in enumerate(names):
                setattr(warr[i], name, L.ptr(weights[8 * i + j]))
                setattr(gwarr[i], name, L.ptr(gws[8 * i + j]))
        rarr = (C.c_void_p * n)(*([None if tc_saved is not None else L.ptr(xs[i + 1])
                                   for i in range(n - 1)] + [None]))
        garr_t = garr_s = None
        if gates:
            garr_t = (C.c_void_p * n)(*[(L.ptr(gates[2 * i]) if gates[2 * i].numel() else None)
                                        for i in range(n)])
            garr_s = (C.c_void_p * n)(*[L.ptr(gates[2 * i + 1]) for i in range(n)])
        ws_bytes = L.lib.vqw_resnet_backward_workspace(C.byref(d))
        workspace = torch.empty(int(ws_bytes), device=dev, dtype=torch.uint8)
        obs = STACK_BACKWARD_OBSERVER if (direct and tc_saved is not None) else None
        evs = obs.events(n) if obs is not None else None
        if evs is not None:
            ev_arr = (C.c_void_p * n)(*[e.cuda_event for e in evs])
            d.block_events = C.cast(ev_arr, C.POINTER(C.c_void_p))
        g_last = _f32c(g_last_res) if (keep_last and g_last_res is not None) else None
        g_res = torch.empty((B, Cr, T, 1), device=dev, dtype=torch.float32) \
            if ctx.needs_input_grad[0] else None
        with L.timed("resnet_backward"):
            L.check(L.lib.vqw_resnet_backward(
                C.byref(d), L.ptr(g_skip), L.ptr(g_last), L.ptr(xs[0]), L.ptr(cond), rarr, garr_t,
                garr_s, warr, L.ptr(g_res), L.ptr(gcond), gwarr, L.ptr(workspace),
                L.ptr(tc_saved), L.stream()), "vqw_resnet_backward")
        if evs is not None:
            obs.launched(targets, evs)
        if direct:
            gws = [None] * len(gws)
        return (g_res, gcond, g_glob, None, None, None, None, None, None, *gws)


def residual_stack(x, cond, dilations, fs, weights, mode=L.MODE_FP32, keep_last_residual=False,
                   grad_targets=None, cond_global=None):
    """`cond` (B,Cc,T,1) is the full condition, or -- with `cond_global` (B,Cg), tensor-core modes
    only -- its Cc-Cg time-varying channels: the time-constant channels are then projected once per
    (item, block) into the gate bias instead of being contracted at every time step."""
    return _ResidualStack.apply(x, cond, cond_global, tuple(dilations), fs, mode, keep_last_residual,
                                grad_targets, torch.is_grad_enabled(), *weights)


# ---------------------------------------------------------------------------------------
# causal embedding of mu-law indices (modules.py:151-152 on a one-hot input)
# ---------------------------------------------------------------------------------------
class _EmbedGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, W, b, mode=L.MODE_FP32):
        W = _f32c(W)
        if q.dtype != torch.int32:
            raise TypeError("embed indices must be int32")
        q = q.contiguous()
        B, T = q.shape[0], q.shape[1]
        Cr, Q = W.shape[0], W.shape[1]
        out = torch.empty((B, Cr, T, 1), device=W.device, dtype=torch.float32)
        L.check(L.lib.vqw_embed_gather_forward(L.ptr(q), L.ptr(W), L.ptr(b), L.ptr(out), B, T, Cr,
                                               Q, L.stream()), "vqw_embed_gather_forward")
        ctx.save_for_backward(q)
        ctx.cfg = (tuple(W.shape), b is not None, mode)
        return out

    @staticmethod
    def backward(ctx, g):
        (q,) = ctx.saved_tensors
        wshape, has_b, mode = ctx.cfg
        g = _f32c(g)
        Cr, Q = wshape[0], wshape[1]
        B, T = q.shape[0], q.shape[1]
        gW = torch.zeros(wshape, device=g.device, dtype=torch.float32)
        gb = torch.zeros(Cr, device=g.device, dtype=torch.float32) if has_b else None
        ws_bytes = -1
        if mode != L.MODE_FP32:
            ws_bytes = int(L.lib.vqw_embed_gather_backward_tc_workspace(B, T, Cr, Q))
        if ws_bytes > 0:      # two tcgen05 GEMMs against a one-hot plane
            ws = torch.empty(ws_bytes, device=g.device, dtype=torch.uint8)
            L.check(L.lib.vqw_embed_gather_backward_tc(L.ptr(q), L.ptr(g), L.ptr(gW), L.ptr(gb), B, T,
                                                       Cr, Q, mode, L.ptr(ws), L.stream()),
                    "vqw_embed_gather_backward_tc")
        else:                 # shared-memory histogram on the CUDA cores
            L.check(L.lib.vqw_embed_gather_backward(L.ptr(q), L.ptr(g), L.ptr(gW), L.ptr(gb), B, T,
                                                    Cr, Q, L.stream()), "vqw_embed_gather_backward")
        return None, gW, gb, None


def embed_gather(q, W, b, mode=L.MODE_FP32):
    """q (B,T) int32 -> (B,Cr,T,1); W is the embed conv's weight (Cr, Q, 2, 1).  `mode` selects
    the weight-gradient kernel (tensor-core GEMMs in the bf16 modes)."""
    return _EmbedGather.apply(q, W, b, mode)


# ---------------------------------------------------------------------------------------
# output head on the tensor cores (modules.py:155-159)
# ---------------------------------------------------------------------------------------
def head_supported(skip: torch.Tensor, mode: int) -> bool:
    B, Cs, T = _as3(skip)
    return mode != L.MODE_FP32 and Cs % 256 == 0 and T >= 128 and T % 8 == 0 and B >= 1


class _Head(torch.autograd.Function):
    """y = proj2(relu(proj1(relu(skip)))) as two tcgen05 GEMMs; the backward is two data-gradient
    GEMMs (ReLU masks applied in the epilogue) and one grouped weight-gradient launch."""

    @staticmethod
    def forward(ctx, skip, W1, b1, W2, b2, mode, grad_enabled):
        skip, W1, b1, W2, b2 = (_f32c(t) for t in (skip, W1, b1, W2, b2))
        B, Cs, T = _as3(skip)
        Q = W2.shape[0]
        d = L.HeadDesc()
        d.B, d.T, d.Cs, d.Q, d.mode = B, T, Cs, Q, mode
        need_grad = grad_enabled and any(ctx.needs_input_grad)
        y = torch.empty((B, Q, T, 1), device=skip.device, dtype=torch.float32)
        ws = torch.empty(int(L.lib.vqw_head_workspace(C.byref(d))), device=skip.device,
                         dtype=torch.uint8)
        # inference: the activation planes live in the workspace (vqw_head_forward, saved = NULL)
        saved = torch.empty(int(L.lib.vqw_head_saved_bytes(C.byref(d))), device=skip.device,
                            dtype=torch.uint8) if (need_grad or d.Q < d.Cs) else None
        with L.timed("head_forward"):
            L.check(L.lib.vqw_head_forward(C.byref(d), L.ptr(skip), L.ptr(W1), L.ptr(b1), L.ptr(W2),
                                           L.ptr(b2), L.ptr(y), L.ptr(ws), L.ptr(saved), L.stream()),
                    "vqw_head_forward")
        if need_grad:
            ctx.cfg = (B, T, Cs, Q, mode)
            ctx.tc_saved = saved
            ctx.save_for_backward(W1, W2)
        return y

    @staticmethod
    def backward(ctx, gy):
        B, T, Cs, Q, mode = ctx.cfg
        W1, W2 = ctx.saved_tensors
        gy = _f32c(gy)
        d = L.HeadDesc()
        d.B, d.T, d.Cs, d.Q, d.mode = B, T, Cs, Q, mode
        dev = gy.device
        gskip = torch.empty((B, Cs, T, 1), device=dev, dtype=torch.float32)
        gW1, gW2 = torch.zeros_like(W1), torch.zeros_like(W2)
        gb1 = torch.zeros(Cs, device=dev, dtype=torch.float32)
        gb2 = torch.zeros(Q, device=dev, dtype=torch.float32)
        ws = torch.empty(int(L.lib.vqw_head_workspace(C.byref(d))), device=dev, dtype=torch.uint8)
        with L.timed("head_backward"):
            L.check(L.lib.vqw_head_backward(C.byref(d), L.ptr(gy), L.ptr(W1), L.ptr(W2),
                                            L.ptr(gskip), L.ptr(gW1), L.ptr(gb1), L.ptr(gW2),
                                            L.ptr(gb2), L.ptr(ws), L.ptr(ctx.tc_saved), L.stream()),
                    "vqw_head_backward")
        return gskip, gW1, gb1, gW2, gb2, None, None


def head(skip, W1, b1, W2, b2, mode):
    return _Head.apply(skip, W1, b1, W2, b2, mode, torch.is_grad_enabled())


def head_loss_supported(skip: torch.Tensor, Q: int, mode: int, use_logistic: bool) -> bool:
    """The fused head + loss (vqw_head_loss_*): tensor-core modes, one accumulator tile of logits."""
    if mode == L.MODE_FP32 or not head_supported(skip, mode):
        return False
    return (Q % 3 == 0 and Q <= 32) if use_logistic else Q <= 256


class _HeadLoss(torch.autograd.Function):
    """loss = L(proj2(relu(proj1(relu(skip)))), t) with the loss and d loss / d y computed in the
    epilogue of the proj2 GEMM (modules.py:155-160 + train.py:92-95 / modules.py:169-230): the
    logits are only materialised when `keep_logits` asks for them."""

    @staticmethod
    def forward(ctx, skip, W1, b1, W2, b2, target, mode, use_logistic, quantize, log_scale_min,
                keep_logits, grad_enabled):
        skip, W1, b1, W2, b2 = (_f32c(t) for t in (skip, W1, b1, W2, b2))
        B, Cs, T = _as3(skip)
        Q = W2.shape[0]
        d = L.HeadDesc()
        d.B, d.T, d.Cs, d.Q, d.mode = B, T, Cs, Q, mode
        if target.numel() != B * T:
            raise ValueError(f"target has {target.numel()} elements, expected B*T = {B * T}")
        t_lab = t_val = None
        if use_logistic:
            t_val = target.reshape(B, T).to(torch.float32).contiguous()
        else:
            t_lab = target.reshape(B, T).to(torch.int32).contiguous()
        y = torch.empty((B, Q, T, 1), device=skip.device, dtype=torch.float32) if keep_logits else None
        loss = torch.zeros(2, device=skip.device, dtype=torch.float64)
        ws = torch.empty(int(L.lib.vqw_head_workspace(C.byref(d))), device=skip.device, dtype=torch.uint8)
        saved = torch.empty(int(L.lib.vqw_head_saved_bytes(C.byref(d))), device=skip.device,
                            dtype=torch.uint8)
        with L.timed("head_forward"):
            L.check(L.lib.vqw_head_loss_forward(C.byref(d), L.ptr(skip), L.ptr(W1), L.ptr(b1), L.ptr(W2),
                                                L.ptr(b2), L.ptr(t_lab), L.ptr(t_val), int(quantize),
                                                float(log_scale_min), L.ptr(loss), L.ptr(y), L.ptr(ws),
                                                L.ptr(saved), L.stream()), "vqw_head_loss_forward")
        if grad_enabled and any(ctx.needs_input_grad):
            ctx.cfg = (B, T, Cs, Q, mode)
            ctx.tc_saved = saved
            ctx.save_for_backward(W1, W2)
        out_y = y if y is not None else skip.new_empty(0)
        ctx.mark_non_differentiable(out_y)
        return loss[0].to(torch.float32).reshape(()), out_y

    @staticmethod
    def backward(ctx, g_loss, _gy):
        B, T, Cs, Q, mode = ctx.cfg
        W1, W2 = ctx.saved_tensors
        d = L.HeadDesc()
        d.B, d.T, d.Cs, d.Q, d.mode = B, T, Cs, Q, mode
        dev = W1.device
        g = g_loss.reshape(1).to(torch.float32).contiguous()
        gskip = torch.empty((B, Cs, T, 1), device=dev, dtype=torch.float32)
        gW1, gW2 = torch.zeros_like(W1), torch.zeros_like(W2)
        gb1 = torch.zeros(Cs, device=dev, dtype=torch.float32)
        gb2 = torch.zeros(Q, device=dev, dtype=torch.float32)
        ws = torch.empty(int(L.lib.vqw_head_workspace(C.byref(d))), device=dev, dtype=torch.uint8)
        with L.timed("head_backward"):
            L.check(L.lib.vqw_head_loss_backward(C.byref(d), L.ptr(g), L.ptr(W1), L.ptr(W2), L.ptr(gskip),
                                                 L.ptr(gW1), L.ptr(gb1), L.ptr(gW2), L.ptr(gb2), L.ptr(ws),
                                                 L.ptr(ctx.tc_saved), L.stream()), "vqw_head_loss_backward")
        return gskip, gW1, gb1, gW2, gb2, None, None, None, None, None, None, None


def head_loss(skip, W1, b1, W2, b2, target, mode, use_logistic, quantize, log_scale_min,
              keep_logits=False):
    """Returns (loss, logits or None)."""
    loss, y = _HeadLoss.apply(skip, W1, b1, W2, b2, target, mode, bool(use_logistic), quantize,
                              log_scale_min, bool(keep_logits), torch.is_grad_enabled())
    return loss, (y if y.numel() else None)


# ---------------------------------------------------------------------------------------
# ConditionEmbed tail: x64 linear upsampling + speaker broadcast + concat (net.py:58-63)
# ---------------------------------------------------------------------------------------
class _UpsampleConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, local, glob, out_len):
        local = _f32c(local)
        B, Cl, H = _as3(local)
        Cg = 0
        if glob is not None:
            glob = _f32c(glob)
            Cg = glob.shape[1]
            if glob.shape[0] != B:
                raise ValueError("local and global condition batch sizes differ")
        out = torch.empty((B, Cl + Cg, out_len, 1), device=local.device, dtype=torch.float32)
        L.check(L.lib.vqw_upsample_concat_forward(L.ptr(local), L.ptr(glob), L.ptr(out), B, Cl, Cg, H,
                                                  out_len, L.stream()), "vqw_upsample_concat_forward")
        ctx.cfg = (B, Cl, Cg, H, out_len)
        return out

    @staticmethod
    def backward(ctx, g):
        B, Cl, Cg, H, out_len = ctx.cfg
        g = _f32c(g)
        g_local = torch.empty((B, Cl, H, 1), device=g.device, dtype=torch.float32)
        g_glob = torch.empty((B, Cg), device=g.device, dtype=torch.float32) if Cg else None
        L.check(L.lib.vqw_upsample_concat_backward(L.ptr(g), L.ptr(g_local), L.ptr(g_glob), B, Cl, Cg, H,
                                                   out_len, L.stream()), "vqw_upsample_concat_backward")
        return g_local, g_glob, None


def upsample_concat(local, glob, out_len):
    """local (B,Cl,H,1), glob (B,Cg) -> (B, Cl+Cg, out_len, 1): F.resize_images of both (the
    global one from length 1 = broadcast) and F.concat, net.py:58-63.  glob=None: the resize
    of the local part alone."""
    return _UpsampleConcat.apply(local, glob, out_len)
