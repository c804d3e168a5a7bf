"""WaveNet decoder behind the reference's callable surface (WaveNet/modules.py:7-255).

ResidualBlock / ResidualNet / WaveNet keep the reference's constructor arguments, attribute
names and parameter tree; the arithmetic runs in libvqw.so (one fused kernel per block)."""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from . import _lib as L
from . import functions as Fn
from .links import Convolution2D


class ResidualBlock(nn.Module):
    """modules.py:7-74."""

    def __init__(self, filter_size, dilation, residual_channels, dilated_channels, skip_channels,
                 condition_dim, dropout_zero_rate):
        super().__init__()
        if dropout_zero_rate:
            raise NotImplementedError("dropout_zero_rate != 0 is outside the hot path "
                                      "(params.py:42 default is 0)")
        self.conv = Convolution2D(residual_channels, dilated_channels, (filter_size, 1),
                                  pad=(dilation * (filter_size - 1), 0), dilate=(dilation, 1))
        self.condition_proj = Convolution2D(condition_dim, dilated_channels, 1)
        self.res = Convolution2D(dilated_channels // 2, residual_channels, 1)
        self.skip = Convolution2D(dilated_channels // 2, skip_channels, 1)
        self.filter_size = filter_size
        self.dilation = dilation
        self.residual_channels = residual_channels
        self.condition_dim = condition_dim
        self.dropout_zero_rate = dropout_zero_rate
        self.mode = L.MODE_FP32

    def weights(self) -> List[torch.Tensor]:
        return [self.conv.W, self.conv.b, self.condition_proj.W, self.condition_proj.b,
                self.res.W, self.res.b, self.skip.W, self.skip.b]

    def forward(self, x, condition):
        """(x, condition) -> (residual, skip), modules.py:30-56."""
        skip, residual = Fn.residual_stack(x, condition, [self.dilation], self.filter_size,
                                           self.weights(), self.mode, keep_last_residual=True)
        return residual, skip


class ResidualNet(nn.ModuleList):
    """modules.py:77-110 (a ChainList: children are named '0', '1', ...)."""

    def __init__(self, n_loop, n_layer, filter_size, residual_channels, dilated_channels,
                 skip_channels, condition_dim, dropout_zero_rate):
        super().__init__()
        dilations = [2 ** i for i in range(n_layer)] * n_loop
        for dilation in dilations:
            self.append(ResidualBlock(filter_size, dilation, residual_channels, dilated_channels,
                                      skip_channels, condition_dim, dropout_zero_rate))
        self.filter_size = filter_size
        self.mode = L.MODE_FP32

    def tc_supported(self, x, n_local) -> bool:
        """Shapes the tcgen05 kernels take (resnet_tc_supported, csrc/resblock_tc.cu);
        n_local = time-varying condition channels that are contracted."""
        blk = self[0]
        Cd, Cr = blk.conv.W.shape[0], blk.conv.W.shape[1]
        Cs, T = blk.skip.W.shape[0], x.shape[2]
        return (Cd == 512 and Cr % 256 == 0 and Cs % 256 == 0 and n_local >= 32 and
                n_local % 32 == 0 and T >= 128 and T % 8 == 0 and x.shape[0] <= 65535)

    def forward(self, x, condition):
        """Sum of the blocks' skip outputs (modules.py:89-96); the last block's residual is
        never used by the reference (modules.py:52,91) and is not computed.  Shapes outside
        the tensor-core kernels' tiling (short or odd-length utterances, small channel counts)
        run the fp32 CUDA-core kernels, like the head does."""
        weights: List[torch.Tensor] = []
        for block in self:
            weights += block.weights()
        mode = self.mode
        cond_global = None
        if isinstance(condition, (tuple, list)):
            # (local (B,Cl,T,1), global (B,Cg)) from ConditionEmbed(split=True): the speaker
            # embedding is constant over time, so its projection is hoisted into the gate bias
            local, glob = condition
            if mode != L.MODE_FP32 and self.tc_supported(x, local.shape[1]):
                condition, cond_global = local, glob
            else:
                condition = torch.cat([local, glob.reshape(glob.shape[0], -1, 1, 1).expand(
                    -1, -1, local.shape[2], 1)], dim=1)
        if mode != L.MODE_FP32 and cond_global is None and not self.tc_supported(x, condition.shape[1]):
            mode = L.MODE_FP32
        return Fn.residual_stack(x, condition, [b.dilation for b in self], self.filter_size,
                                 weights, mode, grad_targets=tuple(weights), cond_global=cond_global)


class WaveNet(nn.Module):
    """modules.py:113-255."""

    def __init__(self, n_loop, n_layer, filter_size, input_dim, residual_channels,
                 dilated_channels, skip_channels, quantize, use_logistic, n_mixture,
                 log_scale_min, condition_dim, dropout_zero_rate):
        super().__init__()
        self.embed = Convolution2D(input_dim, residual_channels, (2, 1), pad=(1, 0))
        self.resnet = ResidualNet(n_loop, n_layer, filter_size, residual_channels,
                                  dilated_channels, skip_channels, condition_dim,
                                  dropout_zero_rate)
        self.proj1 = Convolution2D(skip_channels, skip_channels, 1)
        output_dim = n_mixture if use_logistic else quantize        # modules.py:137-140
        self.proj2 = Convolution2D(skip_channels, output_dim, 1)
        self.input_dim = input_dim
        self.use_logistic = bool(use_logistic)
        if self.use_logistic and input_dim != 1:
            # generate.py:116-137 feeds the drawn VALUE back as a (1,1,1,1) input; a one-hot
            # decoder with a mixture-of-logistics output has no feedback rule in the reference
            import warnings
            warnings.warn("use_logistic=True with input_dim != 1: training works, generation "
                          "is undefined in the reference (generate.py:137) and raises here")
        self.quantize = quantize
        self.skip_channels = skip_channels
        self.log_scale_min = log_scale_min
        self.n_loop, self.n_layer, self.filter_size = n_loop, n_layer, filter_size
        self.residual_channels, self.dilated_channels = residual_channels, dilated_channels
        self.condition_dim = condition_dim

    def set_mode(self, mode) -> None:
        """Arithmetic of the residual stack: 'fp32' | 'fp16x3' | 'bf16x3' | 'bf16' | 'fp16' (include/vqw.h)."""
        m = L.MODES[mode] if isinstance(mode, str) else mode
        self.resnet.mode = m
        for blk in self.resnet:
            blk.mode = m

    def forward(self, x, condition, generating=False):
        """modules.py:148-160.  `x` is the reference's (B, input_dim, T, 1) float tensor, or --
        same arithmetic without materialising the one-hot -- a (B, T) int32 tensor of mu-law
        indices when input_dim == quantize."""
        if x.dtype == torch.int32:
            h = Fn.embed_gather(x, self.embed.W, self.embed.b, self.resnet.mode)  # :151-152
        else:
            h = self.embed(x, out_len=x.shape[2])                              # :150-152
        z = self.resnet(h, condition)                                          # :155
        if Fn.head_supported(z, self.resnet.mode):                             # tensor-core head
            return Fn.head(z, self.proj1.W, self.proj1.b, self.proj2.W, self.proj2.b,
                           self.resnet.mode)                                   # :155-159
        z = self.proj1(z, relu=True, relu_in=True)                             # :155,158
        return self.proj2(z)                                                   # :159

    def forward_loss(self, x, condition, t, keep_logits=False):
        """loss_func(self(x, condition), t) for the reference's two losses -- softmax cross entropy
        (train.py:95) or calculate_logistic_loss (train.py:93), by `use_logistic` -- returning
        (loss, logits or None).  On the tensor-core path head and loss are ONE pass: the loss and
        its gradient are computed in the epilogue of the proj2 GEMM and the logits are only
        written when `keep_logits` asks for them."""
        if x.dtype == torch.int32:
            h = Fn.embed_gather(x, self.embed.W, self.embed.b, self.resnet.mode)
        else:
            h = self.embed(x, out_len=x.shape[2])
        z = self.resnet(h, condition)
        Q = self.proj2.W.shape[0]
        if Fn.head_loss_supported(z, Q, self.resnet.mode, self.use_logistic):
            return Fn.head_loss(z, self.proj1.W, self.proj1.b, self.proj2.W, self.proj2.b, t,
                                self.resnet.mode, self.use_logistic, self.quantize,
                                self.log_scale_min, keep_logits)
        if Fn.head_supported(z, self.resnet.mode):
            y = Fn.head(z, self.proj1.W, self.proj1.b, self.proj2.W, self.proj2.b, self.resnet.mode)
        else:
            y = self.proj2(self.proj1(z, relu=True, relu_in=True))
        if self.use_logistic:
            return self.calculate_logistic_loss(y, t), y
        from .losses import softmax_cross_entropy
        return softmax_cross_entropy(y, t), y

    def calculate_logistic_loss(self, y, t):
        from .losses import logistic_loss
        return logistic_loss(y, t, self.quantize, self.log_scale_min)

    # generation state: see generate.py (persistent kernel); the reference's initialize()/
    # generate() pair (modules.py:232-255) is provided there with the same semantics.
    def initialize(self, n):
        from .generate import WaveNetState
        self._state = WaveNetState(self, n)

    def generate(self, x, condition):
        return self._state.step(x, condition)
