"""Encoder / ConditionEmbed / VAE behind the reference's callable surface (net.py:8-96)."""
from __future__ import annotations

import torch
from torch import nn

from . import functions as Fn
from .links import Convolution2D, DilatedConvolution2D, EmbedID
from .losses import softmax_cross_entropy
from .utils import VQ, ExponentialMovingAverage  # noqa: F401


class Encoder(nn.Module):
    """net.py:8-26: six Convolution2D(k=(4,1), stride=(2,1), pad=(1,0)); ReLU after the first
    five (fused into the conv kernel's epilogue)."""

    def __init__(self, d):
        super().__init__()
        self.conv1 = Convolution2D(1, d, (4, 1), (2, 1), (1, 0))
        self.conv2 = Convolution2D(d, d, (4, 1), (2, 1), (1, 0))
        self.conv3 = Convolution2D(d, d, (4, 1), (2, 1), (1, 0))
        self.conv4 = Convolution2D(d, d, (4, 1), (2, 1), (1, 0))
        self.conv5 = Convolution2D(d, d, (4, 1), (2, 1), (1, 0))
        self.conv6 = Convolution2D(d, d, (4, 1), (2, 1), (1, 0))

    def forward(self, x):
        h = self.conv1(x, relu=True)
        h = self.conv2(h, relu=True)
        h = self.conv3(h, relu=True)
        h = self.conv4(h, relu=True)
        h = self.conv5(h, relu=True)
        return self.conv6(h)


class ConditionEmbed(nn.Module):
    """net.py:29-64: five non-causal dilated convs (dil 1,2,4,8,16; ReLU fused) on the VQ
    output, x upscale_factor linear upsampling, speaker EmbedID broadcast, channel concat
    (local first)."""

    def __init__(self, n_global_cond, global_embed_dim, local_embed_dim, upscale_factor=64,
                 local_in_channels=None):
        super().__init__()
        # the reference passes in_channels=None (lazy); the input is the VQ output, so its
        # channel count is d.  local_in_channels=None defers to local_embed_dim == d
        # (the BASELINE configs), otherwise give d explicitly.
        cin = local_embed_dim if local_in_channels is None else local_in_channels
        self.local_embed1 = DilatedConvolution2D(cin, local_embed_dim, (3, 1), pad=(1, 0), dilate=(1, 1))
        self.local_embed2 = DilatedConvolution2D(local_embed_dim, local_embed_dim, (3, 1), pad=(2, 0), dilate=(2, 1))
        self.local_embed3 = DilatedConvolution2D(local_embed_dim, local_embed_dim, (3, 1), pad=(4, 0), dilate=(4, 1))
        self.local_embed4 = DilatedConvolution2D(local_embed_dim, local_embed_dim, (3, 1), pad=(8, 0), dilate=(8, 1))
        self.local_embed5 = DilatedConvolution2D(local_embed_dim, local_embed_dim, (3, 1), pad=(16, 0), dilate=(16, 1))
        self.global_embed = EmbedID(n_global_cond, global_embed_dim)
        self.upscale_factor = upscale_factor

    def forward(self, local_condition, global_condition, split=False):
        """`split=True` returns (upsampled local (B,Cl,T,1), speaker embedding (B,Cg)) instead of
        their concatenation: the decoder's tensor-core path then hoists the projection of the
        time-constant channels out of the per-step contraction (same arithmetic: W_p is linear
        and F.resize_images of a length-1 signal is a broadcast, net.py:60-61)."""
        h = self.local_embed1(local_condition, relu=True)
        h = self.local_embed2(h, relu=True)
        h = self.local_embed3(h, relu=True)
        h = self.local_embed4(h, relu=True)
        h = self.local_embed5(h, relu=True)
        g = self.global_embed(global_condition)                         # :59
        # resize (:58), broadcast-resize of the speaker embedding (:60-61) and concat (:63): one kernel
        if split:
            return Fn.upsample_concat(h, None, self.upscale_factor * h.shape[2]), g
        return Fn.upsample_concat(h, g, self.upscale_factor * h.shape[2])


class VAE(nn.Module):
    """net.py:67-96.  `__call__(x_enc, x_dec, global_condition, t)` returns
    (loss1, loss2, loss3); `.encoder .vq .condition_embed .decoder` as in the reference
    (updaters.py:16,65 reach into `.vq`).  The VQ kernel runs once: the second quantisation of
    net.py:83 reuses its indices (identical by construction)."""

    def __init__(self, encoder, decoder, condition_embed, d, k, beta, loss_func):
        super().__init__()
        self.beta = beta
        self.loss_func = loss_func
        self.encoder = encoder
        self.vq = VQ(k, d)
        self.condition_embed = condition_embed
        self.decoder = decoder
        self.observation = {}
        # the decoder's logits are a local of the reference's __call__ (net.py:86); set
        # keep_logits to have them in `self.y` (the fused head + loss otherwise never writes them)
        self.keep_logits = False
        self.y = None

    def forward(self, x_enc, x_dec, global_condition, t):
        z = self.encoder(x_enc)                                        # :81
        e = self.vq(z)                                                 # :82
        idx = self.vq.indexes
        zd = z.detach()
        e_ = self.vq(zd, cached=(e.detach(), idx))                     # :83
        dec = getattr(self.decoder, "target", self.decoder)
        split = getattr(getattr(dec, "resnet", None), "mode", 0) != 0   # tensor-core modes hoist it
        condition = self.condition_embed(e, global_condition, split=split)   # :85
        wn = getattr(dec, "forward_loss", None)
        std_loss = wn is not None and (
            self.loss_func is softmax_cross_entropy and not dec.use_logistic or
            getattr(self.loss_func, "__func__", None) is type(dec).calculate_logistic_loss
            and dec.use_logistic)
        if std_loss:
            # :86 + :89 in one pass: loss_func is one of the reference's two losses (train.py:92-95)
            if hasattr(self.decoder, "run"):
                loss1, y = self.decoder.run("forward_loss", x_dec, condition, t, self.keep_logits)
            else:
                loss1, y = self.decoder.forward_loss(x_dec, condition, t, self.keep_logits)
        else:
            y = self.decoder(x_dec, condition)                         # :86
            loss1 = self.loss_func(y, t)                               # :89
        loss2 = torch.mean((zd - e_) ** 2)                             # :90
        loss3 = self.beta * torch.mean((z - e.detach()) ** 2)          # :91
        loss = loss1 + loss2 + loss3
        self.observation = {"loss1": loss1.detach(), "loss2": loss2.detach(),
                            "loss3": loss3.detach(), "loss": loss.detach()}   # reporter, :93-95
        self.y = y.detach() if y is not None else None
        return loss1, loss2, loss3

    @torch.no_grad()
    def generate(self, raw, speaker, use_ema=True, uniforms=None, seed=None, n_steps=None):
        """The reference's convenience entry (models.py:74-105 / generate.py:104-145): encode
        `raw` (1,1,T+1,1), quantise, embed the condition, then draw the whole utterance with the
        persistent kernel.  `use_ema` picks the decoder copy like generate.py:70-76.  `uniforms`
        are the draws numpy.random.choice / numpy.random.uniform would make (generated from
        `seed` when omitted).  Returns the `output` array of generate.py:110,145 (length T, last
        entry 0) as a float64 tensor."""
        import numpy
        from .generate import generate_utterance
        dec = self.decoder
        if hasattr(dec, "ema") and hasattr(dec, "target"):
            dec = dec.ema if use_ema else dec.target
        z = self.encoder(raw)
        e = self.vq(z)
        condition = self.condition_embed(e, speaker)
        T = condition.shape[2]
        steps = T - 1 if n_steps is None else min(n_steps, T - 1)
        if uniforms is None:
            per = dec.proj2.W.shape[0] // 3 if dec.input_dim == 1 else 1
            uniforms = numpy.random.default_rng(seed).uniform(size=(steps, per))
            if per == 1:
                uniforms = uniforms[:, 0]
        return generate_utterance(dec, condition, uniforms, n_steps=steps)
