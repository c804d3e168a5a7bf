"""CPU oracle: a literal restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED BY THE REFERENCE: dhgrs/chainer-VQ-VAE ships no tests, fixtures or golden
vectors, and Chainer/CuPy/librosa are not installable here, so the reference itself cannot
be executed.  What pins this file instead:
  * `oracle/chainer_shim` + `oracle/make_golden.py` import the reference's OWN `net.py`,
    `utils.py` and `WaveNet/modules.py` from /root/reference unmodified on top of a NumPy
    shim of the Chainer primitives they call and freeze their forward outputs under
    `tests/golden/`; `tests/test_oracle_golden.py` checks this file against those vectors.
  * the invariants of SURVEY.md section 4 (tests/test_oracle_invariants.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module.  The product package never does.

Every function cites the reference file:line it follows (paths relative to /root/reference).
All tensors use the reference layout (B, C, T, 1) float; arithmetic is torch-CPU in `dtype`
(float32 mirrors the reference, float64 is the tolerance ground truth).  VQ index selection
follows NumPy's float32 semantics exactly (sequential sum over d, first-minimum argmin).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# configuration (params.py:1-49; the named configs of BASELINE.json / SURVEY.md section 8)
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class Config:
    batch: int = 2
    length: int = 1024
    n_loop: int = 1
    n_layer: int = 4
    filter_size: int = 3
    input_dim: int = 256
    residual_channels: int = 32
    dilated_channels: int = 32
    skip_channels: int = 32
    quantize: int = 256
    use_logistic: bool = False
    n_mixture: int = 30          # number of OUTPUT channels (modules.py:137-138, params.py:38)
    log_scale_min: float = -40.0
    d: int = 64
    k: int = 128
    local_condition_dim: int = 64
    global_condition_dim: int = 128
    n_speaker: int = 109
    beta: float = 0.25
    upscale_factor: int = 64

    @property
    def condition_dim(self) -> int:     # train.py:87
        return self.local_condition_dim + self.global_condition_dim

    @property
    def dilations(self) -> List[int]:   # modules.py:82
        return [2 ** i for i in range(self.n_layer)] * self.n_loop

    @property
    def output_dim(self) -> int:        # modules.py:137-140
        return self.n_mixture if self.use_logistic else self.quantize


def config_cpu() -> Config:
    return Config()


def config_b200() -> Config:
    return Config(batch=16, length=7680, n_loop=2, n_layer=10, filter_size=3,
                  residual_channels=512, dilated_channels=512, skip_channels=256,
                  k=512, d=64)


def config_mol() -> Config:
    c = config_b200()
    c.use_logistic = True
    c.input_dim = 1
    return c


def config_gen() -> Config:
    c = config_b200()
    c.batch = 1
    c.n_loop = 4
    return c


# --------------------------------------------------------------------------------------
# mu-law (utils.py:12-29)
# --------------------------------------------------------------------------------------
class MuLaw:
    def __init__(self, mu=256, int_type=np.int32, float_type=np.float32):   # utils.py:13-16
        self.mu = mu
        self.int_type = int_type
        self.float_type = float_type

    def transform(self, x):                                                 # utils.py:18-23
        x = x.astype(self.float_type)
        y = np.sign(x) * np.log(1 + self.mu * np.abs(x)) / np.log(1 + self.mu)
        y = np.digitize(y, 2 * np.arange(self.mu) / self.mu - 1) - 1
        return y.astype(self.int_type)

    def itransform(self, y):                                                # utils.py:25-29
        y = y.astype(self.float_type)
        y = 2 * y / self.mu - 1
        x = np.sign(y) / self.mu * ((self.mu) ** np.abs(y) - 1)
        return x.astype(self.float_type)


# --------------------------------------------------------------------------------------
# synthetic inputs and parameters
# --------------------------------------------------------------------------------------
def make_inputs(cfg: Config, seed: int = 71, batch: Optional[int] = None):
    """Synthetic batch shaped like Preprocess.__call__'s tuple (utils.py:100-110)."""
    rng = np.random.default_rng(seed)
    B = cfg.batch if batch is None else batch
    T = cfg.length
    sr = 16000.0
    n = np.arange(T + 1, dtype=np.float64)
    raws = []
    for _ in range(B):
        f = rng.uniform(80.0, 4000.0, size=3)
        ph = rng.uniform(0.0, 2 * np.pi, size=3)
        w = sum(np.sin(2 * np.pi * f[i] * n / sr + ph[i]) for i in range(3))
        w = w + rng.normal(0.0, 0.01, size=T + 1)
        w = w / np.abs(w).max()                                             # utils.py:58
        raws.append(w.astype(np.float32))
    raw = np.stack(raws)                                                    # (B, T+1)
    quantized = MuLaw(cfg.quantize).transform(raw)                          # utils.py:63
    speaker = rng.integers(0, cfg.n_speaker, size=B).astype(np.int32)       # utils.py:96-97
    x_enc = raw[:, None, :, None]                                           # utils.py:88-89
    if cfg.input_dim != 1:                                                  # utils.py:84-87,102
        one_hot = np.identity(cfg.quantize, dtype=np.float32)[quantized]    # (B, T+1, Q)
        x_dec = np.ascontiguousarray(one_hot.transpose(0, 2, 1)[:, :, :-1, None])
    else:                                                                   # utils.py:104
        x_dec = np.ascontiguousarray(x_enc[:, :, :-1])
    if cfg.use_logistic:                                                    # utils.py:107
        t = np.ascontiguousarray(x_enc[:, :, 1:])
    else:                                                                   # utils.py:91,109
        t = np.ascontiguousarray(quantized[:, 1:, None]).astype(np.int32)
    return dict(x_enc=np.ascontiguousarray(x_enc), x_dec=x_dec, speaker=speaker, t=t,
                quantized=quantized.astype(np.int32), raw=raw)


def param_shapes(cfg: Config) -> Dict[str, Tuple[int, ...]]:
    """Chainer link-path names and shapes (net.py:12-17,34-44; modules.py:13-22,127-141)."""
    s: Dict[str, Tuple[int, ...]] = {}
    d = cfg.d
    s["encoder/conv1/W"] = (d, 1, 4, 1)
    s["encoder/conv1/b"] = (d,)
    for i in range(2, 7):
        s[f"encoder/conv{i}/W"] = (d, d, 4, 1)
        s[f"encoder/conv{i}/b"] = (d,)
    s["vq/W"] = (cfg.k, d)
    L = cfg.local_condition_dim
    for i in range(1, 6):
        s[f"condition_embed/local_embed{i}/W"] = (L, d if i == 1 else L, 3, 1)
        s[f"condition_embed/local_embed{i}/b"] = (L,)
    s["condition_embed/global_embed/W"] = (cfg.n_speaker, cfg.global_condition_dim)
    Cr, Cd, Cs, Cc = (cfg.residual_channels, cfg.dilated_channels, cfg.skip_channels,
                      cfg.condition_dim)
    s["decoder/embed/W"] = (Cr, cfg.input_dim, 2, 1)
    s["decoder/embed/b"] = (Cr,)
    for i in range(len(cfg.dilations)):
        p = f"decoder/resnet/{i}/"
        s[p + "conv/W"] = (Cd, Cr, cfg.filter_size, 1)
        s[p + "conv/b"] = (Cd,)
        s[p + "condition_proj/W"] = (Cd, Cc, 1, 1)
        s[p + "condition_proj/b"] = (Cd,)
        s[p + "res/W"] = (Cr, Cd // 2, 1, 1)
        s[p + "res/b"] = (Cr,)
        s[p + "skip/W"] = (Cs, Cd // 2, 1, 1)
        s[p + "skip/b"] = (Cs,)
    s["decoder/proj1/W"] = (Cs, Cs, 1, 1)
    s["decoder/proj1/b"] = (Cs,)
    s["decoder/proj2/W"] = (cfg.output_dim, Cs, 1, 1)
    s["decoder/proj2/b"] = (cfg.output_dim,)
    return s


def make_params(cfg: Config, seed: int = 1234, dtype=torch.float32) -> Params:
    """Seeded LeCunNormal weights (std 1/sqrt(fan_in)), N(0, 0.01) biases, EmbedID N(0,1),
    codebook N(0, 1/sqrt(d)) -- SURVEY.md section 8d."""
    rng = np.random.default_rng(seed)
    out: Params = {}
    for name, shape in param_shapes(cfg).items():
        if name == "condition_embed/global_embed/W":
            a = rng.normal(0.0, 1.0, size=shape)
        elif name == "vq/W":
            a = rng.normal(0.0, 1.0 / math.sqrt(shape[1]), size=shape)
        elif name.endswith("/b"):
            a = rng.normal(0.0, 0.01, size=shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            a = rng.normal(0.0, 1.0 / math.sqrt(fan_in), size=shape)
        out[name] = torch.from_numpy(a.astype(np.float32)).to(dtype)
    return out


def sub(params: Params, prefix: str) -> Params:
    n = len(prefix)
    return {k[n:]: v for k, v in params.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------------------
# Chainer primitives restated [dep]
# --------------------------------------------------------------------------------------
def conv2d(x, W, b, stride=1, pad=0, dilate=1):
    """L.Convolution2D / L.DilatedConvolution2D with (k,1) kernels: cross-correlation,
    W (out,in,kh,1), symmetric zero pad, out len (L+2p-dil*(k-1)-1)//s+1."""
    return F.conv2d(x, W, b, stride=(stride, 1), padding=(pad, 0), dilation=(dilate, 1))


def resize_images_h(x, out_h):
    """F.resize_images(x, (out_h, 1)) for W == 1 [dep]: float64 linspace coordinates,
    v0 = clip(floor v, 0, H-2), weights cast to x.dtype, y = w0*x[v0] + w1*x[v0+1].
    For H == 1 Chainer's clip(0, -1) yields v0 = -1, v1 = 0, weights (0, 1): a broadcast."""
    B, C, H, Wd = x.shape
    assert Wd == 1
    if H == 1:
        return x.expand(B, C, out_h, 1).clone()
    v = np.linspace(0, H - 1, num=out_h)
    v0 = np.floor(v).astype(np.int32).clip(0, H - 2)
    v1 = v0 + 1
    w0 = torch.from_numpy((v1 - v)).to(x.dtype).reshape(1, 1, out_h, 1)
    w1 = torch.from_numpy((v - v0)).to(x.dtype).reshape(1, 1, out_h, 1)
    i0 = torch.from_numpy(v0.astype(np.int64))
    i1 = torch.from_numpy(v1.astype(np.int64))
    return w0 * x[:, :, i0] + w1 * x[:, :, i1]


def sigmoid(x):
    """chainer.functions.sigmoid on CPU [dep]: tanh(x * 0.5) * 0.5 + 0.5.  The formulation
    matters for calculate_logistic_loss, where cdf_plus - cdf_min cancels catastrophically in
    float32 (pinned by tests/golden/ref_mol_T256.npz)."""
    half = 0.5
    return torch.tanh(x * half) * half + half


def softmax_cross_entropy(y, t):
    """chainer.functions.softmax_cross_entropy (train.py:95) [dep]: log-softmax over axis 1,
    mean over all labels (normalize=True, ignore_label=-1 never hit)."""
    logp = F.log_softmax(y, dim=1)
    t64 = t.to(torch.int64).unsqueeze(1)                   # (B,1,T,1)
    picked = torch.gather(logp, 1, t64)
    return -picked.sum() / t.numel()


# --------------------------------------------------------------------------------------
# Encoder / ConditionEmbed (net.py:8-64)
# --------------------------------------------------------------------------------------
def encoder_forward(p: Params, x):
    """Encoder.__call__ net.py:19-26 (params relative to 'encoder/')."""
    h = x
    for i in range(1, 7):
        h = conv2d(h, p[f"conv{i}/W"], p[f"conv{i}/b"], stride=2, pad=1)   # net.py:12-17
        if i < 6:
            h = F.relu(h)                                                  # net.py:20-24
    return h


def condition_embed_forward(p: Params, local_condition, global_condition, upscale_factor=64):
    """ConditionEmbed.__call__ net.py:48-64 (params relative to 'condition_embed/')."""
    h = local_condition
    for i, dil in enumerate([1, 2, 4, 8, 16], start=1):                    # net.py:34-43
        h = F.relu(conv2d(h, p[f"local_embed{i}/W"], p[f"local_embed{i}/b"],
                          pad=dil, dilate=dil))                            # net.py:49-53
    h = resize_images_h(h, upscale_factor * h.shape[2])                    # net.py:54-55
    g = p["global_embed/W"][global_condition.to(torch.int64)]              # net.py:57 EmbedID
    g = g.reshape(g.shape + (1, 1))                                        # net.py:58-59
    g = resize_images_h(g, h.shape[2])                                     # net.py:60-61
    return torch.cat((h, g), dim=1)                                        # net.py:63


# --------------------------------------------------------------------------------------
# VQ (utils.py:161-255)
# --------------------------------------------------------------------------------------
def vq_indexes_numpy(xs: np.ndarray, W: np.ndarray) -> np.ndarray:
    """StraightThrough.forward utils.py:189-203, the literal NumPy formulation (materialises
    the (B,k,d,T,1) temporary: small inputs only)."""
    e = W
    xs = np.expand_dims(xs, 1)
    shape = list(xs.shape)
    shape[1] = W.shape[0]
    xs = np.broadcast_to(xs, tuple(shape))
    if xs.ndim == 5:
        Wb = np.broadcast_to(np.reshape(W, (1,) + W.shape + (1, 1)), xs.shape)
    else:
        Wb = np.broadcast_to(np.reshape(W, (1,) + W.shape + (1,)), xs.shape)
    return np.argmin(np.sum((xs - Wb) ** 2, axis=2), axis=1).astype(np.int32)


def vq_indexes(xs: np.ndarray, W: np.ndarray) -> np.ndarray:
    """Same result as vq_indexes_numpy without the 5-D temporary: distances accumulated
    sequentially over d in index order in the input dtype, no FMA, first-minimum argmin
    (equality with the literal form is a test)."""
    x3 = xs.reshape(xs.shape[0], xs.shape[1], -1)            # (B, d, T)
    B, d, T = x3.shape
    k = W.shape[0]
    acc = np.zeros((B, k, T), dtype=xs.dtype)
    for i in range(d):
        diff = x3[:, None, i, :] - W[None, :, i, None]
        acc = acc + diff * diff
    idx = np.argmin(acc, axis=1).astype(np.int32)
    return idx.reshape((xs.shape[0],) + xs.shape[2:])


class _StraightThrough(torch.autograd.Function):
    """utils.py:161-231.  forward: indexes + gather + transpose to channel-first;
    backward: gx = gy (utils.py:218-219); gW = eye(k)[idx].T.dot(gy) accumulated in float64
    then cast (utils.py:222-230)."""

    @staticmethod
    def forward(ctx, xs, W):
        idx = vq_indexes(xs.detach().numpy(), W.detach().numpy())
        ctx.indexes = idx
        ctx.k = W.shape[0]
        ctx.wdtype = W.dtype
        ti = torch.from_numpy(idx.astype(np.int64))
        embeded = W.detach()[ti]                              # utils.py:206
        if embeded.ndim == 4:
            embeded = embeded.permute(0, 3, 1, 2)             # utils.py:207-208
        else:
            embeded = embeded.permute(0, 2, 1)                # utils.py:209-210
        return embeded.contiguous()

    @staticmethod
    def backward(ctx, gy):
        gx = gy
        if gy.ndim == 4:
            g2 = gy.permute(0, 2, 3, 1)
        else:
            g2 = gy.permute(0, 2, 1)
        g2 = g2.reshape(-1, g2.shape[-1])
        onehot = np.eye(ctx.k)[ctx.indexes.reshape(-1)]       # float64, utils.py:227
        gW = onehot.T.dot(g2.detach().numpy())                # utils.py:228
        return gx, torch.from_numpy(gW).to(ctx.wdtype)


def straight_through(x, W):                                   # utils.py:234-236
    return _StraightThrough.apply(x, W)


# --------------------------------------------------------------------------------------
# WaveNet (WaveNet/modules.py)
# --------------------------------------------------------------------------------------
def residual_block_forward(p: Params, x, condition, filter_size, dilation, pad=None):
    """ResidualBlock.__call__ modules.py:30-56 (params relative to 'resnet/<i>/').
    `pad=None` = training (pad dil*(fs-1), modules.py:16); `pad=0` after initialize()
    (modules.py:63)."""
    length = x.shape[2]
    if pad is None:
        pad = dilation * (filter_size - 1)
    h = conv2d(x, p["conv/W"], p["conv/b"], pad=pad, dilate=dilation)      # :40
    h = h[:, :, :length]                                                   # :41
    h = h + conv2d(condition, p["condition_proj/W"], p["condition_proj/b"])  # :44
    tanh_z, sig_z = torch.split(h, h.shape[1] // 2, dim=1)                 # :47
    z = torch.tanh(tanh_z) * sigmoid(sig_z)                                # :48
    if x.shape[2] == z.shape[2]:                                           # :51-54
        residual = conv2d(z, p["res/W"], p["res/b"]) + x
    else:
        residual = conv2d(z, p["res/W"], p["res/b"]) + x[:, :, -1:]
    skip = conv2d(z, p["skip/W"], p["skip/b"])                             # :55
    return residual, skip


def residual_net_forward(p: Params, cfg: Config, x, condition, collect=None):
    """ResidualNet.__call__ modules.py:89-96 (params relative to 'decoder/')."""
    skip_connections = None
    for i, dil in enumerate(cfg.dilations):
        x, skip = residual_block_forward(sub(p, f"resnet/{i}/"), x, condition,
                                         cfg.filter_size, dil)
        if collect is not None:
            collect.append((x, skip))
        skip_connections = skip if i == 0 else skip_connections + skip
    return skip_connections


def wavenet_forward(p: Params, cfg: Config, x, condition, collect=None):
    """WaveNet.__call__ modules.py:148-160 (params relative to 'decoder/')."""
    length = x.shape[2]
    x = conv2d(x, p["embed/W"], p["embed/b"], pad=1)                       # :151
    x = x[:, :, :length, :]                                                # :152
    if collect is not None:
        collect.append((x, None))
    z = F.relu(residual_net_forward(p, cfg, x, condition, collect))        # :155
    z = F.relu(conv2d(z, p["proj1/W"], p["proj1/b"]))                      # :158
    return conv2d(z, p["proj2/W"], p["proj2/b"])                           # :159


def calculate_logistic_loss(cfg: Config, y, t):
    """WaveNet.calculate_logistic_loss modules.py:169-230."""
    nr_mix = y.shape[1] // 3                                               # :173
    logit_probs = y[:, :nr_mix]
    means = y[:, nr_mix:2 * nr_mix]
    log_scales = y[:, 2 * nr_mix:3 * nr_mix]
    log_scales = torch.maximum(log_scales, torch.full_like(log_scales, cfg.log_scale_min))
    t = (127.5 * t).expand_as(means)                                       # :181
    centered_t = t - means
    inv_std = torch.exp(-log_scales)
    half = 127.5 / (cfg.quantize - 1)
    plus_in = inv_std * (centered_t + half)                                # :185
    cdf_plus = sigmoid(plus_in)
    min_in = inv_std * (centered_t - half)                                 # :187
    cdf_min = sigmoid(min_in)
    log_cdf_plus = plus_in - F.softplus(plus_in)                           # :190
    log_one_minus_cdf_min = -F.softplus(min_in)                            # :191
    cdf_delta = cdf_plus - cdf_min                                         # :193
    lo = torch.tensor(127.5 * -0.999, dtype=torch.float32).to(t.dtype)     # :200 (float32 scalar)
    hi = torch.tensor(127.5 * 0.999, dtype=torch.float32).to(t.dtype)      # :208
    inner = torch.log(torch.maximum(cdf_delta, torch.full_like(cdf_delta, 1e-12)))  # :214-215
    log_probs = torch.where(t < lo, log_cdf_plus,
                            torch.where(t > hi, log_one_minus_cdf_min, inner))      # :198-226
    log_probs = log_probs + F.log_softmax(logit_probs, dim=1)              # :228
    return -torch.mean(torch.logsumexp(log_probs, dim=1))                  # :229


def calculate_logistic_loss_numpy(cfg: Config, y: np.ndarray, t: np.ndarray) -> float:
    """modules.py:169-230 evaluated with NumPy in y.dtype exactly as Chainer's CPU functions do
    (sigmoid = tanh(x/2)/2 + 1/2, softplus = max(x,0) + log1p(exp(-|x|)), log_softmax and
    logsumexp with max subtraction) [dep].  In float32 `cdf_plus - cdf_min` cancels
    catastrophically wherever the logistic is narrow, so the loss is only reproducible to
    ~1e-2 rel by an implementation with different elementary-function rounding (torch, CUDA);
    this literal form is what pins the formula to the golden vector."""
    dt = y.dtype
    nr_mix = y.shape[1] // 3
    logit_probs, means = y[:, :nr_mix], y[:, nr_mix:2 * nr_mix]
    log_scales = np.maximum(y[:, 2 * nr_mix:3 * nr_mix], np.full_like(means, cfg.log_scale_min))
    tt = np.broadcast_to(dt.type(127.5) * t, means.shape)
    centered = tt - means
    inv_std = np.exp(-log_scales)
    half_bin = 127.5 / (cfg.quantize - 1)

    def sig(v):
        h = dt.type(0.5)
        return np.tanh(v * h) * h + h

    def softplus(v):
        return np.maximum(v, 0) + np.log1p(np.exp(-np.fabs(v)))
    plus_in = inv_std * (centered + dt.type(half_bin))
    min_in = inv_std * (centered - dt.type(half_bin))
    cdf_delta = sig(plus_in) - sig(min_in)
    log_cdf_plus = plus_in - softplus(plus_in)
    log_one_minus_cdf_min = -softplus(min_in)
    lo = np.full(tt.shape, 127.5 * -0.999, dtype=np.float32)
    hi = np.full(tt.shape, 127.5 * 0.999, dtype=np.float32)
    inner = np.log(np.maximum(cdf_delta, np.full(cdf_delta.shape, 1e-12, dtype=np.float32)))
    log_probs = np.where(tt < lo, log_cdf_plus, np.where(tt > hi, log_one_minus_cdf_min, inner))
    m = logit_probs.max(axis=1, keepdims=True)
    ls = (logit_probs - m) - np.log(np.exp(logit_probs - m).sum(axis=1, keepdims=True))
    log_probs = log_probs + ls
    mm = log_probs.max(axis=1, keepdims=True)
    lse = np.log(np.exp(log_probs - mm).sum(axis=1)) + np.squeeze(mm, axis=1)
    return float(-lse.mean())


class WaveNetGenerator:
    """WaveNet.initialize / generate with the reference's concat-shift queues
    (modules.py:58-74, 98-110, 232-255)."""

    def __init__(self, p: Params, cfg: Config, n: int = 1):
        self.p, self.cfg = p, cfg
        dt = p["embed/W"].dtype
        Cr, fs = cfg.residual_channels, cfg.filter_size
        self.queues = [torch.zeros(n, Cr, dil * (fs - 1) + 1, 1, dtype=dt)
                       for dil in cfg.dilations]                           # :59-62
        self.cond_queues = [torch.zeros(n, cfg.condition_dim, 1, 1, dtype=dt)
                            for _ in cfg.dilations]                        # :64-66
        self.embed_queue = torch.zeros(n, cfg.input_dim, 2, 1, dtype=dt)   # :236-237
        self.proj1_queue = torch.zeros(n, cfg.skip_channels, 1, 1, dtype=dt)   # :239-240
        self.proj2_queue3 = torch.zeros(n, cfg.skip_channels, 1, 1, dtype=dt)  # :242-243

    def generate(self, x, condition, collect=None):
        p, cfg = self.p, self.cfg
        self.embed_queue = torch.cat((self.embed_queue[:, :, 1:], x), dim=2)   # :246
        x = conv2d(self.embed_queue, p["embed/W"], p["embed/b"], pad=0)        # :247, pad (0,0) :235
        skip_connections = None
        for i, dil in enumerate(cfg.dilations):                                # :102-110
            self.queues[i] = torch.cat((self.queues[i][:, :, 1:], x), dim=2)   # :72
            self.cond_queues[i] = torch.cat((self.cond_queues[i][:, :, 1:], condition), dim=2)
            x, skip = residual_block_forward(sub(p, f"resnet/{i}/"), self.queues[i],
                                             self.cond_queues[i], cfg.filter_size, dil,
                                             pad=0)                            # :68-69
            if collect is not None:
                collect.append(x)
            skip_connections = skip if i == 0 else skip_connections + skip
        x = F.relu(skip_connections)                                           # :248
        self.proj1_queue = torch.cat((self.proj1_queue[:, :, 1:], x), dim=2)   # :250
        x = F.relu(conv2d(self.proj1_queue, p["proj1/W"], p["proj1/b"]))       # :251
        self.proj2_queue3 = torch.cat((self.proj2_queue3[:, :, 1:], x), dim=2)  # :253
        return conv2d(self.proj2_queue3, p["proj2/W"], p["proj2/b"])           # :254


def choice_from_uniform(prob: np.ndarray, u: float) -> int:
    """numpy.random.choice(n, p=prob) given its uniform draw [dep]:
    cdf = cumsum(p) in float64, cdf /= cdf[-1], searchsorted(cdf, u, side='right')."""
    cdf = np.cumsum(prob.astype(np.float64))
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


def generate_loop(params: Params, cfg: Config, x_enc, speaker, uniforms: np.ndarray,
                  n_steps: Optional[int] = None, return_logits: bool = False):
    """generate.py:104-145.  `uniforms[i]` replaces the i-th random draw (categorical: one
    per step; MoL: nr_mix per step).  Returns the generated `output` array
    (generate.py:110,145) and optionally the per-step decoder outputs."""
    dt = params["vq/W"].dtype
    x_enc = torch.as_tensor(x_enc).to(dt)
    z = encoder_forward(sub(params, "encoder/"), x_enc)                        # :105
    e = straight_through(z, params["vq/W"])                                    # :106
    condition = condition_embed_forward(sub(params, "condition_embed/"), e,
                                        torch.as_tensor(speaker), cfg.upscale_factor)  # :108
    dec = sub(params, "decoder/")
    gen = WaveNetGenerator(dec, cfg, 1)                                        # :109
    total = condition.shape[2]
    output = np.zeros(total)                                                   # :110
    steps = total - 1 if n_steps is None else min(n_steps, total - 1)
    x_dec = torch.zeros(1, cfg.input_dim, 1, 1, dtype=dt)                      # :51
    logits = []
    with torch.no_grad():
        for i in range(steps):                                                 # :112
            out = gen.generate(x_dec, condition[:, :, i:i + 1])                # :115
            if return_logits:
                logits.append(out[0, :, 0, 0].numpy().copy())
            if cfg.use_logistic:                                               # :116-137
                o = out.numpy()
                nr_mix = o.shape[1] // 3
                logit_probs = o[:, :nr_mix]
                means = o[:, nr_mix:2 * nr_mix]
                log_scales = np.maximum(o[:, 2 * nr_mix:3 * nr_mix], cfg.log_scale_min)
                scales = np.exp(log_scales)
                rand = np.asarray(uniforms[i], dtype=np.float64).reshape(logit_probs.shape)
                rand = means + scales * (np.log(rand) - np.log(1 - rand))      # :127-128
                sm = torch.softmax(torch.from_numpy(logit_probs), dim=1).numpy()
                rand = (rand * sm).sum(axis=1)                                 # :131-132
                value = np.squeeze(rand.astype(np.float32))                    # :134
                value = value / np.float32(127.5)
                value = np.clip(value, -1, 1)                                  # :136
                x_dec = torch.full((1, 1, 1, 1), float(value), dtype=dt)       # :137
            else:                                                              # :138-144
                prob = torch.softmax(out, dim=1)[0, :, 0, 0].numpy()
                value = choice_from_uniform(prob, float(uniforms[i]))          # :139-141
                x_dec = torch.zeros(1, cfg.input_dim, 1, 1, dtype=dt)
                x_dec[:, value] = 1                                            # :142-144
            output[i] = value                                                  # :145
    if return_logits:
        return output, np.stack(logits) if logits else np.zeros((0, cfg.output_dim))
    return output


# --------------------------------------------------------------------------------------
# VAE forward + the three-loss step (net.py:67-96, updaters.py:6-19)
# --------------------------------------------------------------------------------------
def vae_forward(params: Params, cfg: Config, x_enc, x_dec, global_condition, t, collect=None):
    """VAE.__call__ net.py:79-96: returns (loss1, loss2, loss3) and a dict of intermediates."""
    z = encoder_forward(sub(params, "encoder/"), x_enc)                        # :81
    W = params["vq/W"]
    e = straight_through(z, W)                                                 # :82
    e_ = straight_through(z.detach(), W)                                       # :83
    condition = condition_embed_forward(sub(params, "condition_embed/"), e,
                                        global_condition, cfg.upscale_factor)  # :85
    y = wavenet_forward(sub(params, "decoder/"), cfg, x_dec, condition, collect)  # :86
    if cfg.use_logistic:                                                       # train.py:92-95
        loss1 = calculate_logistic_loss(cfg, y, t)
    else:
        loss1 = softmax_cross_entropy(y, t)
    loss2 = torch.mean((z.detach() - e_) ** 2)                                 # :90
    loss3 = cfg.beta * torch.mean((z - e.detach()) ** 2)                       # :91
    idx = vq_indexes(z.detach().numpy(), W.detach().numpy())
    return (loss1, loss2, loss3), dict(z=z, e=e, condition=condition, y=y, indexes=idx)


def three_loss_grads(params: Params, cfg: Config, x_enc, x_dec, global_condition, t):
    """VQVAE_StandardUpdater.update_core updaters.py:13-18: forward, cleargrads,
    loss1.backward(), vq.cleargrads(), loss2.backward(), loss3.backward().
    Returns (losses, grads by parameter name, intermediates)."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    (l1, l2, l3), inter = vae_forward(leaf, cfg, x_enc, x_dec, global_condition, t)
    l1.backward(retain_graph=True)                                             # :15
    leaf["vq/W"].grad = None                                                   # :16
    l2.backward(retain_graph=True)                                             # :17
    l3.backward()                                                              # :18
    grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v))
             for k, v in leaf.items()}
    return (l1.detach(), l2.detach(), l3.detach()), grads, inter


def adam_step(param, grad, m, v, t, alpha, beta1=0.9, beta2=0.999, eps=1e-8):
    """chainer.optimizers.Adam update rule (train.py:101) [dep]; in place; t is 1-based."""
    m += (1 - beta1) * (grad - m)
    v += (1 - beta2) * (grad * grad - v)
    lr = alpha * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    param -= lr * m / (torch.sqrt(v) + eps)


def weight_ema_update(target: Params, ema: Params, decay: float) -> None:
    """ExponentialMovingAverage.__call__ utils.py:146-155: decay multiplies the TARGET."""
    for name, tp in target.items():
        ema[name] = decay * tp + (1 - decay) * ema[name]                       # :153-154


def parallel_update_grads(params: Params, cfg: Config, batch, n: int):
    """VQVAE_ParallelUpdater.update_core updaters.py:23-72: the batch is split batch[i::n]
    (:36-38), each replica back-propagates its own mean losses, gradients are SUMMED into
    the main model (:71-72).  Returns (list of per-replica losses, summed grads)."""
    total = None
    losses = []
    for i in range(n):
        sl = slice(i, None, n)
        xe = torch.as_tensor(batch["x_enc"][sl]).to(params["vq/W"].dtype)
        xd = torch.as_tensor(batch["x_dec"][sl]).to(params["vq/W"].dtype)
        sp = torch.as_tensor(batch["speaker"][sl])
        tt = torch.as_tensor(batch["t"][sl])
        if cfg.use_logistic:
            tt = tt.to(params["vq/W"].dtype)
        ls, g, _ = three_loss_grads(params, cfg, xe, xd, sp, tt)
        losses.append(ls)
        total = g if total is None else {k: total[k] + g[k] for k in g}
    return losses, total


def receptive_field(cfg: Config) -> int:
    """n_loop*(fs-1)*(2^n_layer-1)+2."""
    return cfg.n_loop * (cfg.filter_size - 1) * (2 ** cfg.n_layer - 1) + 2
