"""Profiling driver for the fused residual-block forward at the 1xB200 shape.

  python tools/prof_fwd.py time [nblk]      CUDA-event time per block launch for the shipped
                                            kernel, the round-1 schedule and prefetch distances
  python tools/prof_fwd.py run [mode] [nblk] [bwd]   a few blocks (for ncu / VQW_TC_TIMELINE=1)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chainer_vq_vae_b200 as V
from chainer_vq_vae_b200 import _lib as L

B, T, Cr, Cd, Cs, Cc, fs = 16, 7680, 512, 512, 256, 192, 3
dev = "cuda"


def make(nblk, bwd=False):
    torch.manual_seed(0)
    ws = []
    for i in range(nblk):
        ws += [torch.randn(Cd, Cr, fs, 1, device=dev) / (Cr * fs) ** 0.5, torch.randn(Cd, device=dev) * 0.01,
               torch.randn(Cd, Cc, 1, 1, device=dev) / Cc ** 0.5, torch.randn(Cd, device=dev) * 0.01,
               torch.randn(Cr, Cd // 2, 1, 1, device=dev) / 16, torch.randn(Cr, device=dev) * 0.01,
               torch.randn(Cs, Cd // 2, 1, 1, device=dev) / 16, torch.randn(Cs, device=dev) * 0.01]
    if bwd:
        ws = [w.requires_grad_(True) for w in ws]
    x = torch.randn(B, Cr, T, 1, device=dev, requires_grad=bwd)
    c = torch.randn(B, Cc, T, 1, device=dev)
    dil = [2 ** (i % 10) for i in range(nblk)]
    return ws, x, c, dil


def time_variants(nblk):
    ws, x, c, dil = make(nblk, bwd=True)           # training forward: gates and planes saved
    res = {}
    for name, env in (("warm-up", {}), ("shipped", {}), ("shipped pf=8", {"VQW_TC_PREFETCH": "8"})):
        for k in ("VQW_TC_FWD_V1", "VQW_TC_PREFETCH"):
            os.environ.pop(k, None)
        os.environ.update(env)
        L.enable_timers(True)
        for it in range(4):
            skip = V.residual_stack(x, c, dil, fs, ws, L.MODES["bf16x3"])
            del skip
        torch.cuda.synchronize()
        n, ms = L.timer_summary()["resnet_forward"]
        L.enable_timers(False)
        res[name] = ms / nblk
        print(f"{name:28s}: {ms / nblk:.4f} ms per block (call of {nblk} blocks incl. packing: {ms:.3f} ms)",
              flush=True)
    return res


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "time"
    if what == "time":
        time_variants(int(sys.argv[2]) if len(sys.argv) > 2 else 10)
    else:
        mode = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
        nblk = int(sys.argv[3]) if len(sys.argv) > 3 else 3
        bwd = (sys.argv[4] == "bwd") if len(sys.argv) > 4 else False
        ws, x, c, dil = make(nblk, bwd)
        for it in range(2):
            skip = V.residual_stack(x, c, dil, fs, ws, L.MODES[mode])
            if bwd:
                skip.sum().backward()
        torch.cuda.synchronize()
        print("done", float(skip.abs().mean()))
