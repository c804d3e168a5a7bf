"""Profiling driver: a few residual blocks forward+backward at the 1xB200 shape (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chainer_vq_vae_b200 as V
from chainer_vq_vae_b200 import _lib as L

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bwd = (sys.argv[3] == "bwd") if len(sys.argv) > 3 else False
B, T, Cr, Cd, Cs, Cc, fs = 16, 7680, 512, 512, 256, 192, 3
torch.manual_seed(0)
dev = "cuda"
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
for it in range(2):
    skip = V.residual_stack(x, c, dil, fs, ws, L.MODES[mode])
    if bwd:
        skip.sum().backward()
torch.cuda.synchronize()
print("done", float(skip.abs().mean()))
