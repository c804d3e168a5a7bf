"""The CUDA optimiser kernels against the oracle's restatement of chainer.optimizers.Adam
(train.py:101) and of the decoder weight-EMA (utils.py:153-154): vqw_adam_step (host learning
rate), vqw_adam_step_dev (rate read from device memory: the CUDA-graph form) and
vqw_ema_update, five steps each, at 1e-6."""
import numpy as np
import pytest
import torch

import chainer_vq_vae_b200 as V
from chainer_vq_vae_b200 import _lib as L
from oracle import vqvae_oracle as O
from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 1023, 4096 * 37 + 5])
@pytest.mark.parametrize("dev_lr", [False, True])
def test_cuda_adam_kernel_five_steps_vs_oracle(n, dev_lr):
    rng = np.random.default_rng(n)
    p0 = rng.normal(size=n).astype(np.float32)
    grads = [(rng.normal(size=n) * 10.0 ** rng.uniform(-6, 1)).astype(np.float32) for _ in range(5)]
    alpha, b1, b2, eps = 2e-4, 0.9, 0.999, 1e-8
    # oracle (float32 tensors, the reference's dtype)
    po, mo, vo = torch.from_numpy(p0.copy()), torch.zeros(n), torch.zeros(n)
    # float64 ground truth of the same recurrence
    p64, m64, v64 = torch.from_numpy(p0.astype(np.float64)), torch.zeros(n, dtype=torch.float64), \
        torch.zeros(n, dtype=torch.float64)
    pg = torch.from_numpy(p0.copy()).cuda()
    mg, vg = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    lr_dev = torch.zeros(1, device="cuda")
    for t, g in enumerate(grads, start=1):
        O.adam_step(po, torch.from_numpy(g), mo, vo, t, alpha, b1, b2, eps)
        O.adam_step(p64, torch.from_numpy(g.astype(np.float64)), m64, v64, t, alpha, b1, b2, eps)
        gg = torch.from_numpy(g).cuda()
        lr = alpha * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        if dev_lr:
            lr_dev.fill_(lr)
            L.check(L.lib.vqw_adam_step_dev(L.ptr(pg), L.ptr(gg), L.ptr(mg), L.ptr(vg), n,
                                            L.ptr(lr_dev), b1, b2, eps, L.stream()), "adam_dev")
        else:
            L.check(L.lib.vqw_adam_step(L.ptr(pg), L.ptr(gg), L.ptr(mg), L.ptr(vg), n, lr, b1, b2,
                                        eps, L.stream()), "adam")
    torch.cuda.synchronize()
    # the accumulated step (what Adam changed), relative to its own size
    step_g = pg.cpu().double() - torch.from_numpy(p0.astype(np.float64))
    step_64 = p64 - torch.from_numpy(p0.astype(np.float64))
    step_o = po.double() - torch.from_numpy(p0.astype(np.float64))
    assert rel_err(step_g, step_64) < 5e-3        # vs exact arithmetic: only fp32 ulps of p (~1e-7 of |p|, steps ~1e-3)
    assert rel_err(pg, po) < 1e-6 and rel_err(mg, mo) < 1e-6 and rel_err(vg, vo) < 1e-6
    assert rel_err(step_g, step_o) < 5e-3
    assert rel_err(mg, m64) < 1e-6 and rel_err(vg, v64) < 1e-6


def test_cuda_ema_kernel_vs_oracle_quirk():
    """utils.py:153-154: ema = decay * TARGET + (1 - decay) * ema (decay multiplies the target)."""
    n, decay = 100003, 0.9999
    rng = np.random.default_rng(0)
    tgt = {"w": torch.from_numpy(rng.normal(size=n).astype(np.float32))}
    ema = {"w": torch.from_numpy(rng.normal(size=n).astype(np.float32))}
    eg, tg = ema["w"].clone().cuda(), tgt["w"].clone().cuda()
    for _ in range(5):
        O.weight_ema_update(tgt, ema, decay)
        L.check(L.lib.vqw_ema_update(L.ptr(eg), L.ptr(tg), n, decay, L.stream()), "ema")
    torch.cuda.synchronize()
    assert rel_err(eg, ema["w"]) < 1e-6
