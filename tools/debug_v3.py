"""Where does the CTA-pair forward kernel (VQW_TC_FWD_V3=1) differ from the shipped one?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chainer_vq_vae_b200 as V
from chainer_vq_vae_b200 import _lib as L

B, T, Cr, Cd, Cs, Cc, fs = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 512, 512, 256, 192, 3
torch.manual_seed(0)
dev = "cuda"
ws = [torch.randn(Cd, Cr, fs, 1, device=dev) / (Cr * fs) ** 0.5, torch.randn(Cd, device=dev) * 0.01,
      torch.randn(Cd, Cc, 1, 1, device=dev) / Cc ** 0.5, torch.randn(Cd, device=dev) * 0.01,
      torch.randn(Cr, Cd // 2, 1, 1, device=dev) / 16, torch.randn(Cr, device=dev) * 0.01,
      torch.randn(Cs, Cd // 2, 1, 1, device=dev) / 16, torch.randn(Cs, device=dev) * 0.01]
x = torch.randn(B, Cr, T, 1, device=dev)
c = torch.randn(B, Cc, T, 1, device=dev)
out = {}
for v3 in ("0", "1"):
    os.environ["VQW_TC_FWD_V3"] = v3
    with torch.no_grad():
        skip, res = V.residual_stack(x, c, [1], fs, ws, L.MODES["bf16x3"], keep_last_residual=True)
    torch.cuda.synchronize()
    out[v3] = (skip[0, :, :, 0].clone(), res[0, :, :, 0].clone())
for name, i in (("skip", 0), ("res", 1)):
    a, b = out["1"][i], out["0"][i]
    print(name, "max |ref|", float(b.abs().max()), "overall max err", float((a - b).abs().max()))
    for h in range((T + 127) // 128):
        row = []
        for cb in range(a.shape[0] // 64):
            e = (a[cb * 64:(cb + 1) * 64, h * 128:(h + 1) * 128] - b[cb * 64:(cb + 1) * 64, h * 128:(h + 1) * 128]).abs().max()
            row.append(f"{float(e):8.1e}")
        print(f"  t-half {h}: " + " ".join(row))
