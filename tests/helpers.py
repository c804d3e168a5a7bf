"""Shared test helpers: build the product modules from an oracle parameter dict."""
import numpy as np
import torch

import chainer_vq_vae_b200 as V
from oracle import vqvae_oracle as O

TOL = 1e-3   # BASELINE.json north_star: fp activations within 1e-3 rel


def rel_err(a, b):
    """max |a-b| / max |b|  (scale-relative error)."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def build_model(cfg: O.Config, params, device="cuda", ema_decay=None, mode="fp32"):
    wavenet = V.WaveNet(cfg.n_loop, cfg.n_layer, cfg.filter_size, cfg.input_dim,
                        cfg.residual_channels, cfg.dilated_channels, cfg.skip_channels,
                        cfg.quantize, cfg.use_logistic, cfg.n_mixture, cfg.log_scale_min,
                        cfg.condition_dim, 0)
    encoder = V.Encoder(cfg.d)
    cond = V.ConditionEmbed(cfg.n_speaker, cfg.global_condition_dim, cfg.local_condition_dim,
                            cfg.upscale_factor, local_in_channels=cfg.d)
    decoder = V.ExponentialMovingAverage(wavenet, ema_decay) if ema_decay else wavenet
    loss = wavenet.calculate_logistic_loss if cfg.use_logistic else V.softmax_cross_entropy
    model = V.VAE(encoder, decoder, cond, cfg.d, cfg.k, cfg.beta, loss)
    model.keep_logits = True          # the tests compare model.y with the oracle's logits
    load_params(model, params, ema=bool(ema_decay))
    decoder.set_mode(mode)      # both the training copy and the EMA (evaluation) copy
    return model.to(device)


def load_params(model, params, ema=False):
    own = dict(model.named_parameters())
    with torch.no_grad():
        for name, val in params.items():
            key = name.replace("/", ".")
            if ema and key.startswith("decoder."):
                for sub in ("target", "ema"):
                    own["decoder." + sub + key[len("decoder"):]].copy_(val)
            else:
                own[key].copy_(val)


def grads_by_name(model, ema=False):
    out = {}
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        key = name.replace(".", "/")
        if ema:
            if key.startswith("decoder/ema/"):
                continue
            key = key.replace("decoder/target/", "decoder/")
        out[key] = p.grad.detach().cpu()
    return out


def to_dev(inp, cfg, device="cuda", indices=False):
    x_enc = torch.from_numpy(inp["x_enc"]).to(device)
    if indices and cfg.input_dim != 1:
        x_dec = torch.from_numpy(inp["quantized"][:, :-1].copy()).to(device)
    else:
        x_dec = torch.from_numpy(inp["x_dec"]).to(device)
    spk = torch.from_numpy(inp["speaker"]).to(device)
    t = torch.from_numpy(inp["t"]).to(device)
    return x_enc, x_dec, spk, t
