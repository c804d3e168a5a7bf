"""Autoregressive generation (generate.py:104-145 of the reference) on the persistent kernel.

`generate_utterance` is the whole sample loop in one cooperative launch; `WaveNetState` gives
the reference's step-wise `WaveNet.initialize(n)` / `WaveNet.generate(x, condition)` pair
(modules.py:232-255) on the same kernel, one step per call, with the dilation queues living in
the kernel's workspace (ring buffers) instead of concat-shifted Variables."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy
import torch

from . import _lib as L


def _desc(wavenet, T_total, n_steps, t_start):
    blocks = list(wavenet.resnet)
    n = len(blocks)
    d = L.GenerateDesc()
    d.n_blocks = n
    dil = (C.c_int * n)(*[b.dilation for b in blocks])
    d.dilations = C.cast(dil, C.POINTER(C.c_int))
    d.fs, d.Cr, d.Cd = wavenet.filter_size, wavenet.residual_channels, wavenet.dilated_channels
    d.Cs, d.Cc, d.Q = wavenet.skip_channels, wavenet.condition_dim, wavenet.proj2.W.shape[0]
    d.T_total, d.n_steps, d.t_start = T_total, n_steps, t_start
    use_logistic = bool(getattr(wavenet, "use_logistic", wavenet.input_dim == 1))
    if use_logistic != (wavenet.input_dim == 1):
        raise NotImplementedError(
            "generation needs use_logistic == (input_dim == 1): generate.py:116-145 feeds a scalar "
            "back for the mixture-of-logistics decoder and a one-hot for the categorical one")
    d.use_logistic = 1 if use_logistic else 0
    d.log_scale_min = float(wavenet.log_scale_min)
    warr = (L.ResblockWeights * n)()
    for i, b in enumerate(blocks):
        for name, t in zip(("conv_w", "conv_b", "cond_w", "cond_b", "res_w", "res_b", "skip_w",
                            "skip_b"), b.weights()):
            setattr(warr[i], name, L.ptr(t.detach()))
    return d, dil, warr


class WaveNetState:
    """initialize(n) state: queues are zeroed on the first step (modules.py:59-66,236-243)."""

    def __init__(self, wavenet, n):
        if n != 1:
            raise NotImplementedError("generation supports n = 1, like generate.py:42")
        self.w = wavenet
        self.mol = wavenet.input_dim == 1   # mixture of logistics: scalar input (generate.py:137)
        self.t = 0
        # input of the previous step: index (-1 = all zeros) or, for MoL, the float's bit pattern
        self.prev = 0 if self.mol else -1
        self.workspace = None

    def _launch(self, cond2d, t_start, n_steps, uniforms, forced, want_logits, state=None,
                cond_t0=0):
        w = self.w
        dev = cond2d.device
        d, _dil, warr = _desc(w, cond2d.shape[1], n_steps, t_start)
        d.cond_t0 = cond_t0
        if state is not None:
            d.set_state, d.s1, d.s2 = 1, state[0], state[1]
        if self.workspace is None:
            nbytes = L.lib.vqw_generate_workspace(C.byref(d))
            self.workspace = torch.empty(int(nbytes), device=dev, dtype=torch.uint8)
        samples = torch.empty(n_steps, device=dev, dtype=torch.int32)
        logits = torch.empty(n_steps, d.Q, device=dev, dtype=torch.float32) if want_logits else None
        L.check(L.lib.vqw_generate(
            C.byref(d), warr, L.ptr(w.embed.W.detach()), L.ptr(w.embed.b.detach()),
            L.ptr(w.proj1.W.detach()), L.ptr(w.proj1.b.detach()), L.ptr(w.proj2.W.detach()),
            L.ptr(w.proj2.b.detach()), L.ptr(cond2d), L.ptr(uniforms), L.ptr(forced),
            L.ptr(samples), L.ptr(logits), L.ptr(self.workspace), L.stream()), "vqw_generate")
        return samples, logits

    def step(self, x, condition):
        """One `WaveNet.generate(x, condition)` call (modules.py:245-255): x (1, Q, 1, 1) one-hot
        or all-zero float (generate.py:51,142-144), condition (1, Cc, 1, 1).  Returns the decoder
        output (1, Q, 1, 1).  The embed queue (modules.py:246) is the pair (previous x, x)."""
        if x.shape[0] != 1 or x.shape[2] != 1:
            raise ValueError("generate() takes one time step of one utterance")
        xv = x.reshape(-1)
        if self.mol:
            cur = int(xv.float().cpu().view(torch.int32)[0])
        else:
            cur = int(torch.argmax(xv)) if bool((xv != 0).any()) else -1
        cond2d = condition.reshape(condition.shape[1], 1).contiguous().float()
        u = torch.full((max(1, self.w.proj2.W.shape[0] // 3),), 0.5, device=cond2d.device,
                       dtype=torch.float64)
        _, logits = self._launch(cond2d, self.t, 1, u, None, True, state=(cur, self.prev),
                                 cond_t0=self.t)
        self.prev = cur
        self.t += 1
        return logits.reshape(1, -1, 1, 1)


def generate_utterance(wavenet, condition: torch.Tensor, uniforms, n_steps: Optional[int] = None,
                       forced=None, return_logits: bool = False):
    """The loop of generate.py:109-145 for one utterance in one kernel launch.

    condition (1, Cc, T, 1) from ConditionEmbed; uniforms: the draws numpy.random.choice would
    make, one per step.  Returns `output` as generate.py builds it -- length T, the last entry
    left 0 (generate.py:110-112) -- and optionally the per-step decoder outputs (steps, Q)."""
    if condition.shape[0] != 1:
        raise NotImplementedError("generation supports n = 1, like generate.py:42")
    T = condition.shape[2]
    steps = T - 1 if n_steps is None else min(n_steps, T - 1)
    dev = condition.device
    cond2d = condition.reshape(condition.shape[1], T).contiguous().float()
    mol = wavenet.input_dim == 1
    u = numpy.asarray(uniforms, dtype=numpy.float64)
    if mol:     # one draw per mixture component and step (generate.py:124)
        u = u.reshape(-1, wavenet.proj2.W.shape[0] // 3)
    u = torch.as_tensor(numpy.ascontiguousarray(u[:steps])).to(dev)
    f = None
    if forced is not None:
        fa = numpy.asarray(forced, dtype=numpy.float32 if mol else numpy.int32)[:steps]
        f = torch.as_tensor(fa.view(numpy.int32) if mol else fa).to(dev)
    st = WaveNetState(wavenet, 1)
    samples, logits = st._launch(cond2d, 0, steps, u, f, return_logits)
    out = torch.zeros(T, device=dev, dtype=torch.float64)
    out[:steps] = (samples.view(torch.float32) if mol else samples).double()
    return (out, logits) if return_logits else out


# ---------------------------------------------------------------------------------------
# output formats (generate.py:147-153)
# ---------------------------------------------------------------------------------------
def output_to_wave(output, quantize: int = 256, use_logistic: bool = False) -> numpy.ndarray:
    """generate.py:149-152: the mixture-of-logistics decoder emits the waveform itself, the
    categorical decoder emits mu-law indices that MuLaw(quantize).itransform expands
    (utils.py:25-29, with its mu**|y| quirk)."""
    from .utils import MuLaw
    out = output.detach().cpu().numpy() if torch.is_tensor(output) else numpy.asarray(output)
    if use_logistic:
        return out.astype(numpy.float32)
    return MuLaw(quantize).itransform(out)


def write_wav(path, output, sr: int = 16000, quantize: int = 256, use_logistic: bool = False) -> None:
    """generate.py:153 `librosa.output.write_wav(args.output, wave, params.sr)`: librosa 0.5.1
    hands the float32 array to scipy.io.wavfile.write, i.e. a 32-bit IEEE-float mono WAVE file."""
    from scipy.io import wavfile
    wave = output_to_wave(output, quantize, use_logistic)
    wavfile.write(str(path), int(sr), wave.astype(numpy.float32))
