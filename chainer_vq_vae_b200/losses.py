"""Output losses of the path: softmax cross entropy (train.py:95) and the discretised
mixture-of-logistics NLL (WaveNet.calculate_logistic_loss, modules.py:169-230)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


class _SoftmaxCE(torch.autograd.Function):
    """Loss and d loss / d y in one pass over the logits (vqw_softmax_ce)."""

    @staticmethod
    def forward(ctx, y, t, grad_enabled):
        from . import _lib as L
        yc = y.contiguous()
        B, Q, T = yc.shape[0], yc.shape[1], yc.shape[2]
        tc = t.reshape(B, T).to(torch.int32).contiguous()
        # needs_input_grad reflects requires_grad, not the grad mode: under no_grad / in
        # evaluation the gradient tensor must not be allocated
        need = grad_enabled and ctx.needs_input_grad[0]
        gy = torch.empty_like(yc) if need else None
        loss = torch.zeros(2, device=y.device, dtype=torch.float64)   # {loss, valid-label count}
        L.check(L.lib.vqw_softmax_ce(L.ptr(yc), L.ptr(tc), L.ptr(gy), L.ptr(loss), B, Q, T, L.stream()),
                "vqw_softmax_ce")
        if need:
            ctx.save_for_backward(gy)
        return loss[0].to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        (gy,) = ctx.saved_tensors
        return gy * g, None, None


def softmax_cross_entropy(y, t):
    """chainer.functions.softmax_cross_entropy: log-softmax over axis 1, mean over all labels.
    y (B,Q,T,1) f32, t (B,T,1) int.  CUDA tensors run the fused libvqw kernel; the torch
    formula below is the generic definition for other tensors."""
    if y.is_cuda and y.dtype == torch.float32 and y.dim() == 4 and y.shape[3] == 1:
        return _SoftmaxCE.apply(y, t, torch.is_grad_enabled())
    # generic definition (host-side checks only; CUDA tensors never take it)
    return F.cross_entropy(y, t.long().reshape(y.shape[0], *y.shape[2:]), ignore_index=-1)


class _MolLoss(torch.autograd.Function):
    """modules.py:169-230 and its gradient in one pass over the decoder output (vqw_mol_loss)."""

    @staticmethod
    def forward(ctx, y, t, quantize, log_scale_min, grad_enabled):
        from . import _lib as L
        yc = y.contiguous()
        B, C3, T = yc.shape[0], yc.shape[1], yc.shape[2]
        tc = t.reshape(B, T).to(torch.float32).contiguous()
        need = grad_enabled and ctx.needs_input_grad[0]
        gy = torch.empty_like(yc) if need else None
        loss = torch.zeros(1, device=y.device, dtype=torch.float64)
        L.check(L.lib.vqw_mol_loss(L.ptr(yc), L.ptr(tc), L.ptr(gy), L.ptr(loss), B, C3 // 3, T,
                                   int(quantize), float(log_scale_min), L.stream()), "vqw_mol_loss")
        if need:
            ctx.save_for_backward(gy)
        return loss.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        (gy,) = ctx.saved_tensors
        return gy * g, None, None, None, None


def logistic_loss(y, t, quantize, log_scale_min):
    """modules.py:169-230.  CUDA tensors run the fused libvqw kernel; the torch formula below is
    the same definition term by term for other tensors."""
    if (y.is_cuda and y.dtype == torch.float32 and y.dim() == 4 and y.shape[3] == 1
            and y.shape[1] % 3 == 0 and t.numel() == y.shape[0] * y.shape[2]):
        return _MolLoss.apply(y, t, quantize, log_scale_min, torch.is_grad_enabled())
    nr_mix = y.shape[1] // 3
    logit_probs = y[:, :nr_mix]
    means = y[:, nr_mix:2 * nr_mix]
    log_scales = torch.clamp_min(y[:, 2 * nr_mix:3 * nr_mix], log_scale_min)      # :178-179
    t = (127.5 * t).expand_as(means)                                              # :181
    centered_t = t - means
    inv_std = torch.exp(-log_scales)
    half = 127.5 / (quantize - 1)
    plus_in = inv_std * (centered_t + half)
    cdf_plus = torch.sigmoid(plus_in)
    min_in = inv_std * (centered_t - half)
    cdf_min = torch.sigmoid(min_in)
    log_cdf_plus = plus_in - F.softplus(plus_in)
    log_one_minus_cdf_min = -F.softplus(min_in)
    cdf_delta = cdf_plus - cdf_min
    inner = torch.log(torch.clamp_min(cdf_delta, 1e-12))                          # :214-215
    lo = torch.tensor(127.5 * -0.999, dtype=torch.float32, device=y.device)
    hi = torch.tensor(127.5 * 0.999, dtype=torch.float32, device=y.device)
    log_probs = torch.where(t < lo, log_cdf_plus,
                            torch.where(t > hi, log_one_minus_cdf_min, inner))    # :198-226
    log_probs = log_probs + F.log_softmax(logit_probs, dim=1)                     # :228
    return -torch.mean(torch.logsumexp(log_probs, dim=1))                         # :229
