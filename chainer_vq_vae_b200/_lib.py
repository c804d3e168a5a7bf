"""ctypes binding of libvqw.so (include/vqw.h).  The product path has NO fallback: if the
shared library is missing or a symbol is absent, importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libvqw.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `./build.sh` (or __graft_entry__.build()); "
        "this package has no CPU / eager fallback")

lib = C.CDLL(LIB_PATH)

VQW_MAX_SRC = 4
MODE_FP32, MODE_BF16X3, MODE_BF16, MODE_FP16, MODE_FP16X3 = 0, 1, 2, 3, 4
MODES = {"fp32": MODE_FP32, "bf16x3": MODE_BF16X3, "bf16": MODE_BF16, "fp16": MODE_FP16,
         "fp16x3": MODE_FP16X3}
X3_MODES = (MODE_BF16X3, MODE_FP16X3)      # split hi/lo operand planes, sigma-only gate save

c_float_p = C.c_void_p   # device pointers are passed as integers
c_int = C.c_int


class ConvSrc(C.Structure):
    _fields_ = [("in_", C.c_void_p), ("w", C.c_void_p), ("in_mask", C.c_void_p),
                ("K", c_int), ("Tin", c_int), ("wm", c_int), ("wk", c_int),
                ("mul", c_int), ("shift", c_int), ("div", c_int), ("relu_in", c_int)]


class ConvDesc(C.Structure):
    _fields_ = [("B", c_int), ("M", c_int), ("T", c_int), ("nsrc", c_int),
                ("src", ConvSrc * VQW_MAX_SRC),
                ("bias", C.c_void_p), ("addend", C.c_void_p), ("out_mask", C.c_void_p),
                ("relu_out", c_int), ("accumulate", c_int),
                ("gate_tanh", C.c_void_p), ("gate_sig", C.c_void_p)]


class WgradDesc(C.Structure):
    _fields_ = [("B", c_int), ("M", c_int), ("T", c_int),
                ("a", C.c_void_p), ("a_mask", C.c_void_p), ("in_", C.c_void_p),
                ("in_mul", C.c_void_p), ("K", c_int), ("Tin", c_int),
                ("mul", c_int), ("shift", c_int), ("div", c_int), ("relu_in", c_int),
                ("gm", c_int), ("gk", c_int),
                ("ntaps", c_int), ("tap_dshift", c_int), ("tap_gw", c_int)]


class ResblockDesc(C.Structure):
    _fields_ = [("B", c_int), ("T", c_int), ("Cr", c_int), ("Cd", c_int), ("Cs", c_int),
                ("Cc", c_int), ("fs", c_int), ("dilation", c_int),
                ("skip_accumulate", c_int), ("write_residual", c_int), ("mode", c_int)]


class ResblockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("conv_w", "conv_b", "cond_w", "cond_b", "res_w",
                                          "res_b", "skip_w", "skip_b")]


class ResnetDesc(C.Structure):
    _fields_ = [("B", c_int), ("T", c_int), ("Cr", c_int), ("Cd", c_int), ("Cs", c_int),
                ("Cc", c_int), ("fs", c_int), ("n_blocks", c_int),
                ("dilations", C.POINTER(c_int)), ("mode", c_int), ("keep_last_residual", c_int),
                ("Cg", c_int), ("cond_global", C.c_void_p), ("g_cond_global", C.c_void_p),
                ("block_events", C.POINTER(C.c_void_p))]


class GenerateDesc(C.Structure):
    _fields_ = [("n_blocks", c_int), ("dilations", C.POINTER(c_int)), ("fs", c_int),
                ("Cr", c_int), ("Cd", c_int), ("Cs", c_int), ("Cc", c_int), ("Q", c_int),
                ("T_total", c_int), ("n_steps", c_int), ("t_start", c_int),
                ("set_state", c_int), ("s1", c_int), ("s2", c_int), ("cond_t0", c_int),
                ("use_logistic", c_int), ("log_scale_min", C.c_float)]


class HeadDesc(C.Structure):
    _fields_ = [("B", c_int), ("T", c_int), ("Cs", c_int), ("Q", c_int), ("mode", c_int)]


_SIGNATURES = {
    "vqw_version": (c_int, []),
    "vqw_last_error": (C.c_char_p, []),
    "vqw_launch_count": (C.c_longlong, []),
    "vqw_probe_forward_kernels": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vqw_vq_forward": (c_int, [C.c_void_p] * 7 + [c_int] * 4 + [C.c_void_p]),
    "vqw_vq_backward_w": (c_int, [C.c_void_p] * 3 + [c_int] * 4 + [C.c_void_p]),
    "vqw_conv_forward": (c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p]),
    "vqw_conv_wgrad": (c_int, [C.POINTER(WgradDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "vqw_resblock_forward": (c_int, [C.POINTER(ResblockDesc), C.c_void_p, C.c_void_p,
                                     C.POINTER(ResblockWeights)] + [C.c_void_p] * 5),
    "vqw_resblock_backward_workspace": (C.c_int64, [C.POINTER(ResblockDesc)]),
    "vqw_resblock_backward": (c_int, [C.POINTER(ResblockDesc)] + [C.c_void_p] * 6 +
                              [C.POINTER(ResblockWeights), C.c_void_p, C.c_void_p,
                               C.POINTER(ResblockWeights), C.c_void_p, C.c_void_p]),
    "vqw_resnet_forward_workspace": (C.c_int64, [C.POINTER(ResnetDesc)]),
    "vqw_resnet_saved_bytes": (C.c_int64, [C.POINTER(ResnetDesc)]),
    "vqw_resnet_forward": (c_int, [C.POINTER(ResnetDesc), C.c_void_p, C.c_void_p,
                                   C.POINTER(ResblockWeights), C.POINTER(C.c_void_p), C.c_void_p,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "vqw_resnet_backward_workspace": (C.c_int64, [C.POINTER(ResnetDesc)]),
    "vqw_resnet_backward": (c_int, [C.POINTER(ResnetDesc)] + [C.c_void_p] * 4 +
                            [C.POINTER(C.c_void_p)] * 3 + [C.POINTER(ResblockWeights), C.c_void_p,
                                                           C.c_void_p, C.POINTER(ResblockWeights),
                                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "vqw_generate_workspace": (C.c_int64, [C.POINTER(GenerateDesc)]),
    "vqw_generate": (c_int, [C.POINTER(GenerateDesc), C.POINTER(ResblockWeights)] +
                     [C.c_void_p] * 13),
    "vqw_head_workspace": (C.c_int64, [C.POINTER(HeadDesc)]),
    "vqw_head_saved_bytes": (C.c_int64, [C.POINTER(HeadDesc)]),
    "vqw_head_forward": (c_int, [C.POINTER(HeadDesc)] + [C.c_void_p] * 9),
    "vqw_head_backward": (c_int, [C.POINTER(HeadDesc)] + [C.c_void_p] * 11),
    "vqw_head_loss_forward": (c_int, [C.POINTER(HeadDesc)] + [C.c_void_p] * 7 + [c_int, C.c_float] +
                              [C.c_void_p] * 5),
    "vqw_head_loss_backward": (c_int, [C.POINTER(HeadDesc)] + [C.c_void_p] * 11),
    "vqw_softmax_ce": (c_int, [C.c_void_p] * 4 + [c_int] * 3 + [C.c_void_p]),
    "vqw_mol_loss": (c_int, [C.c_void_p] * 4 + [c_int] * 4 + [C.c_float, C.c_void_p]),
    "vqw_upsample_concat_forward": (c_int, [C.c_void_p] * 3 + [c_int] * 5 + [C.c_void_p]),
    "vqw_upsample_concat_backward": (c_int, [C.c_void_p] * 3 + [c_int] * 5 + [C.c_void_p]),
    "vqw_adam_step": (c_int, [C.c_void_p] * 4 + [C.c_longlong] + [C.c_double] * 4 + [C.c_void_p]),
    "vqw_adam_step_dev": (c_int, [C.c_void_p] * 4 + [C.c_longlong, C.c_void_p] + [C.c_double] * 3 +
                          [C.c_void_p]),
    "vqw_ema_update": (c_int, [C.c_void_p] * 2 + [C.c_longlong, C.c_double, C.c_void_p]),
    "vqw_embed_gather_forward": (c_int, [C.c_void_p] * 4 + [c_int] * 4 + [C.c_void_p]),
    "vqw_embed_gather_backward": (c_int, [C.c_void_p] * 4 + [c_int] * 4 + [C.c_void_p]),
    "vqw_embed_gather_backward_tc_workspace": (C.c_int64, [c_int] * 4),
    "vqw_embed_gather_backward_tc": (c_int, [C.c_void_p] * 4 + [c_int] * 5 + [C.c_void_p] * 2),
}


def _bind():
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args


_bind()

def launch_count() -> int:
    """Number of CUDA kernels launched through libvqw.so by this process."""
    return int(lib.vqw_launch_count())


class VqwError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.vqw_last_error().decode("utf-8", "replace")
        raise VqwError(f"{what} failed (rc={rc}): {msg}")


def ptr(t) -> int:
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise VqwError("libvqw operates on CUDA tensors only (no CPU fallback); got a "
                       f"{t.device} tensor")
    if not t.is_contiguous():
        raise VqwError("libvqw needs contiguous tensors")
    return t.data_ptr()


def stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


# ---------------------------------------------------------------------------------------
# optional per-kernel CUDA-event timing (bench.py's roofline leg)
# ---------------------------------------------------------------------------------------
TIMERS = None   # dict name -> list of (start_event, end_event) when enabled


def enable_timers(on: bool = True) -> None:
    global TIMERS
    TIMERS = {} if on else None


class timed:
    """`with timed("name"):` brackets a C-ABI call with CUDA events on the current stream."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if TIMERS is not None:
            import torch
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if TIMERS is not None:
            self.b.record()
            TIMERS.setdefault(self.name, []).append((self.a, self.b))
        return False


def probe_forward_kernels(name="resblock_forward_kernels") -> None:
    """When timers are on: have the next tensor-core vqw_resnet_forward bracket its block kernels
    (only them, not the operand packing) with two timing events filed under `name`."""
    if TIMERS is None:
        return
    import torch
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    a.record()
    b.record()          # materialise the CUDA event handles; the library records them again
    check(lib.vqw_probe_forward_kernels(C.c_void_p(a.cuda_event), C.c_void_p(b.cuda_event)),
          "vqw_probe_forward_kernels")
    TIMERS.setdefault(name, []).append((a, b))


def timer_summary():
    """name -> (launch count, mean ms); call after torch.cuda.synchronize()."""
    out = {}
    for name, evs in (TIMERS or {}).items():
        ms = [a.elapsed_time(b) for a, b in evs]
        out[name] = (len(ms), sum(ms) / max(len(ms), 1))
    return out
