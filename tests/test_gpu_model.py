"""GPU parity of the whole path at the CPU-reference config (BASELINE.json configs[0]):
VAE forward, the three-loss backward ordering, one optimiser step, the weight-EMA quirk."""
import numpy as np
import pytest
import torch

import chainer_vq_vae_b200 as V
from oracle import vqvae_oracle as O
from helpers import TOL, build_model, grads_by_name, rel_err, to_dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cpu_case():
    cfg = O.config_cpu()
    params = O.make_params(cfg)
    inp = O.make_inputs(cfg)
    args = [torch.from_numpy(inp[k]) for k in ("x_enc", "x_dec", "speaker", "t")]
    losses, grads, inter = O.three_loss_grads(params, cfg, *args)
    return cfg, params, inp, losses, grads, inter


@pytest.mark.parametrize("indices", [False, True])
def test_vae_forward_matches_oracle(cpu_case, indices):
    cfg, params, inp, losses, grads, inter = cpu_case
    model = build_model(cfg, params)
    l1, l2, l3 = model(*to_dev(inp, cfg, indices=indices))
    assert np.array_equal(model.vq.indexes.cpu().numpy(), inter["indexes"]), \
        "VQ indices must be bit-identical to the reference formulation"
    assert rel_err(model.y, inter["y"]) < TOL
    for got, want in zip((l1, l2, l3), losses):
        assert abs(float(got.detach()) - float(want)) <= TOL * abs(float(want))


def test_three_loss_backward_ordering(cpu_case):
    cfg, params, inp, losses, grads, inter = cpu_case
    model = build_model(cfg, params)
    opt = V.Adam(2e-4).setup(model)
    upd = V.VQVAE_StandardUpdater(None, opt)
    l1, l2, l3 = model(*to_dev(inp, cfg, indices=True))
    upd.backward_three(model, l1, l2, l3)
    got = grads_by_name(model)
    assert set(got) == set(grads)
    for name, g in grads.items():
        assert rel_err(got[name], g) < TOL, name
    # vq.W carries loss2's gradient only (updaters.py:16): (2/M)(n_k W_k - sum z)
    z, idx, W = inter["z"].detach().numpy(), inter["indexes"], params["vq/W"].numpy()
    M = z.size
    want = np.zeros_like(W, dtype=np.float64)
    cnt = np.bincount(idx.ravel(), minlength=cfg.k)
    zs = np.zeros_like(want)
    np.add.at(zs, idx.ravel(), np.transpose(z[..., 0], (0, 2, 1)).reshape(-1, cfg.d))
    want = (2.0 / M) * (cnt[:, None] * W - zs)
    assert rel_err(got["vq/W"], want) < 1e-4


def test_adam_step_and_weight_ema(cpu_case):
    cfg, params, inp, losses, grads, inter = cpu_case
    decay = 0.9999
    model = build_model(cfg, params, ema_decay=decay)
    opt = V.Adam(2e-4).setup(model)

    class It:
        def next(self_inner):
            return None
    upd = V.VQVAE_StandardUpdater(It(), opt, converter=lambda b, d: to_dev(inp, cfg, indices=True))
    model.train()
    upd.update()
    # oracle: EMA refresh happens inside the forward, BEFORE the optimiser step (utils.py:142-155)
    dec = O.sub(params, "decoder/")
    ema = {k: v.clone() for k, v in dec.items()}
    O.weight_ema_update(dec, ema, decay)
    new = {k: v.clone() for k, v in params.items()}
    for k in new:
        m, v = torch.zeros_like(new[k]), torch.zeros_like(new[k])
        O.adam_step(new[k], grads[k], m, v, 1, 2e-4)
    own = dict(model.named_parameters())
    for k in new:
        key = k.replace("/", ".")
        if key.startswith("decoder."):
            tkey = "decoder.target" + key[len("decoder"):]
            ekey = "decoder.ema" + key[len("decoder"):]
            assert rel_err(own[ekey], ema[k[len("decoder/"):]]) < 1e-6, ekey
        else:
            tkey = key
        # Adam's first step moves every weight by ~alpha*sign(g): compare the step itself
        step_got = own[tkey].detach().cpu() - params[k]
        step_want = new[k] - params[k]
        big = grads[k].abs() > 1e-3 * grads[k].abs().max()
        if not bool(big.any()):
            continue
        assert (step_got[big] - step_want[big]).abs().max() <= 0.05 * 2e-4 + 1e-9, k
    model.eval()
    with torch.no_grad():
        model(*to_dev(inp, cfg, indices=True))       # evaluation runs the EMA copy (utils.py:156-157)


def test_causality(cpu_case):
    cfg, params, inp, *_ = cpu_case
    model = build_model(cfg, params)
    dec = model.decoder
    rng = np.random.default_rng(0)
    B, T = 2, 256
    q = torch.from_numpy(rng.integers(0, 256, size=(B, T)).astype(np.int32)).cuda()
    c = torch.from_numpy(rng.normal(size=(B, cfg.condition_dim, T, 1)).astype(np.float32)).cuda()
    with torch.no_grad():
        y0 = dec(q, c)
        t0 = 100
        q2, c2 = q.clone(), c.clone()
        q2[:, t0:] = (q2[:, t0:] + 17) % 256
        c2[:, :, t0:] += 1.0
        y1 = dec(q2, c2)
    assert torch.equal(y0[:, :, :t0], y1[:, :, :t0]), "outputs before t0 must be bit-unchanged"
    assert not torch.equal(y0[:, :, t0:], y1[:, :, t0:])


def test_mol_vae_matches_oracle():
    """BASELINE.json configs[3] in miniature: use_logistic=True, input_dim=1, 30 output
    channels.  Activations and the VQ/commitment losses to 1e-3; the MoL loss itself is only
    reproducible to ~2e-2 by ANY independent float32 evaluation."""
    cfg = O.config_cpu()
    cfg.use_logistic, cfg.input_dim, cfg.length = True, 1, 512
    params = O.make_params(cfg)
    inp = O.make_inputs(cfg)
    args = [torch.from_numpy(inp[k]) for k in ("x_enc", "x_dec", "speaker", "t")]
    losses, grads, inter = O.three_loss_grads(params, cfg, *args)
    model = build_model(cfg, params)
    opt = V.Adam(2e-4).setup(model)
    upd = V.VQVAE_StandardUpdater(None, opt)
    l1, l2, l3 = model(*to_dev(inp, cfg))
    assert np.array_equal(model.vq.indexes.cpu().numpy(), inter["indexes"])
    assert rel_err(model.y, inter["y"]) < TOL
    assert abs(float(l1.detach()) - float(losses[0])) <= 2e-2 * abs(float(losses[0]))
    assert abs(float(l2.detach()) - float(losses[1])) <= TOL * abs(float(losses[1]))
    upd.backward_three(model, l1, l2, l3)
    got = grads_by_name(model)
    assert rel_err(got["vq/W"], grads["vq/W"]) < TOL
    # the float32 MoL gradient inherits the loss's conditioning: check it is finite and has the
    # oracle's direction
    a, b = got["decoder/proj2/W"].flatten().double(), grads["decoder/proj2/W"].flatten().double()
    assert torch.isfinite(a).all()
    assert float(torch.dot(a, b) / (a.norm() * b.norm())) > 0.99


def test_cuda_graph_updater_matches_eager_updater():
    """VQVAE_StandardUpdater(use_cuda_graph=True): three ordinary steps, one capture, then replays
    -- same losses and parameters as issuing every step eagerly (different batch every step, the
    bias-corrected Adam learning rate read from device memory)."""
    cfg = O.config_cpu()
    cfg.length = 256
    params = O.make_params(cfg)
    batches = [O.make_inputs(cfg, seed=100 + i) for i in range(7)]

    class It:
        def __init__(self):
            self.i = 0

        def next(self):
            inp = batches[self.i]
            self.i += 1
            return [(inp["x_enc"][b], inp["quantized"][b, :-1].astype(np.int32), inp["speaker"][b],
                     inp["t"][b]) for b in range(cfg.batch)]

    out = {}
    for graph in (False, True):
        model = build_model(cfg, params, ema_decay=0.9999)
        model.train()
        # eps = 1e-4 instead of 1e-8: Adam normalises every element's step to ~lr whatever the size
        # of its gradient, so with the default eps the last-bit run-to-run differences of the
        # atomically accumulated weight gradients (conv_wgrad_kernel's split-K) flip the direction
        # of elements whose gradient is at noise level -- in eager-vs-eager just as in graph-vs-
        # eager.  The larger eps makes the comparison well conditioned; the code path (device-side
        # learning rate, captured Adam / EMA kernels) is the same.
        opt = V.Adam(1e-3, eps=1e-4).setup(model)
        upd = V.VQVAE_StandardUpdater(It(), opt, device="cuda", use_cuda_graph=graph, graph_warmup=3)
        losses = [[float(v) for v in upd.update()] for _ in range(len(batches))]
        assert (upd._graph is not None) == graph
        assert opt.t == len(batches)
        out[graph] = (losses, {n: p.detach().clone() for n, p in model.named_parameters()})
    for a, b in zip(out[True][0], out[False][0]):
        assert np.allclose(a, b, rtol=2e-4, atol=1e-7), (a, b)
    # parameters after 7 Adam steps
    for n, p in out[False][1].items():
        a, b = out[True][1][n].double(), p.double()
        scale = float(b.abs().max()) or 1.0
        diff = (a - b).abs()
        assert float(diff.max()) <= 2e-4 * scale + 2e-5, (n, float(diff.max()), scale)
