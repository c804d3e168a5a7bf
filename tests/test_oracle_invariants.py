"""Known-answer invariants of the reference algorithm, checked on the
oracle.  CPU only."""
import numpy as np
import torch

from oracle import vqvae_oracle as O


def _small():
    cfg = O.config_cpu()
    cfg.length = 256
    return cfg, O.make_params(cfg), O.make_inputs(cfg)


def _args(inp, dtype=torch.float32):
    return (torch.from_numpy(inp["x_enc"]).to(dtype), torch.from_numpy(inp["x_dec"]).to(dtype),
            torch.from_numpy(inp["speaker"]), torch.from_numpy(inp["t"]))


def test_vq_loop_equals_literal_numpy_and_ties():
    rng = np.random.default_rng(0)
    for shape, k in (((2, 8, 6, 1), 16), ((3, 64, 5, 1), 128), ((2, 8, 6), 16)):
        z = rng.normal(size=shape).astype(np.float32)
        W = rng.normal(size=(k, shape[1])).astype(np.float32)
        W[k - 1] = W[2]                       # duplicate row: the lowest index must win
        z[0, :, 0] = W[2].reshape(z[0, :, 0].shape)
        a, b = O.vq_indexes(z, W), O.vq_indexes_numpy(z, W)
        assert np.array_equal(a, b)
        assert a.reshape(shape[0], -1)[0, 0] == 2


def test_straight_through_gradients():
    rng = np.random.default_rng(1)
    z = torch.from_numpy(rng.normal(size=(2, 8, 6, 1)).astype(np.float32)).requires_grad_(True)
    W = torch.from_numpy(rng.normal(size=(16, 8)).astype(np.float32)).requires_grad_(True)
    gy = torch.from_numpy(rng.normal(size=(2, 8, 6, 1)).astype(np.float32))
    O.straight_through(z, W).backward(gy)
    assert torch.equal(z.grad, gy)                                   # utils.py:218-219
    idx = O.vq_indexes(z.detach().numpy(), W.detach().numpy()).ravel()
    want = np.zeros((16, 8))
    np.add.at(want, idx, gy.numpy()[..., 0].transpose(0, 2, 1).reshape(-1, 8).astype(np.float64))
    assert np.allclose(W.grad.numpy(), want.astype(np.float32), rtol=0, atol=0)


def test_causality_and_receptive_field():
    cfg, params, inp = _small()
    dec = O.sub(params, "decoder/")
    rng = np.random.default_rng(2)
    x = torch.from_numpy(inp["x_dec"]).double()
    c = torch.from_numpy(rng.normal(size=(2, cfg.condition_dim, cfg.length, 1)))
    p64 = {k: v.double() for k, v in dec.items()}
    with torch.no_grad():
        y0 = O.wavenet_forward(p64, cfg, x, c)
        t0 = 100
        x2, c2 = x.clone(), c.clone()
        x2[:, :, t0:] = torch.roll(x2[:, :, t0:], 1, 1)
        c2[:, :, t0:] += 1.0
        y1 = O.wavenet_forward(p64, cfg, x2, c2)
        assert torch.equal(y0[:, :, :t0], y1[:, :, :t0])
        # an impulse at t0 in x alone reaches exactly receptive_field steps
        x3 = x.clone()
        x3[:, :, t0] = torch.roll(x3[:, :, t0], 1, 1)
        y2 = O.wavenet_forward(p64, cfg, x3, c)
        changed = ((y2 - y0).abs().amax(dim=(0, 1, 3)) > 0).nonzero().ravel()
        assert int(changed.min()) == t0
        assert int(changed.max()) == t0 + O.receptive_field(cfg) - 1


def test_incremental_generation_equals_full_forward():
    cfg, params, inp = _small()
    p64 = {k: v.double() for k, v in params.items()}
    with torch.no_grad():
        _, inter = O.vae_forward(p64, cfg, *_args(inp, torch.float64))
        gen = O.WaveNetGenerator(O.sub(p64, "decoder/"), cfg, 2)
        x_dec = torch.from_numpy(inp["x_dec"]).double()
        for i in range(40):
            o = gen.generate(x_dec[:, :, i:i + 1], inter["condition"][:, :, i:i + 1])
            assert (o[:, :, 0] - inter["y"][:, :, i]).abs().max() < 1e-12


def test_three_loss_backward_ordering():
    cfg, params, inp = _small()
    p64 = {k: v.double() for k, v in params.items()}
    losses, grads, inter = O.three_loss_grads(p64, cfg, *_args(inp, torch.float64))
    # vq.W <- d loss2 only: (2/M)(n_k W_k - sum_{idx=k} z)    
    z, idx, W = inter["z"].detach().numpy(), inter["indexes"], p64["vq/W"].numpy()
    cnt = np.bincount(idx.ravel(), minlength=cfg.k)
    zs = np.zeros_like(W)
    np.add.at(zs, idx.ravel(), z[..., 0].transpose(0, 2, 1).reshape(-1, cfg.d))
    assert np.allclose(grads["vq/W"].numpy(), (2.0 / z.size) * (cnt[:, None] * W - zs), atol=1e-14)
    # decoder and ConditionEmbed <- d loss1 only
    leaf = {k: v.clone().requires_grad_(True) for k, v in p64.items()}
    (l1, l2, l3), _ = O.vae_forward(leaf, cfg, *_args(inp, torch.float64))
    l1.backward()
    for k in leaf:
        if k.startswith("decoder/") or k.startswith("condition_embed/"):
            g1 = leaf[k].grad if leaf[k].grad is not None else torch.zeros_like(leaf[k])
            assert torch.allclose(g1, grads[k], atol=1e-14), k   # (last block's res is unused)
    # encoder <- d loss1 (through the straight-through) + d loss3
    assert not torch.allclose(leaf["encoder/conv6/W"].grad, grads["encoder/conv6/W"], atol=1e-12)


def test_dp_equivalence_of_summed_replica_gradients():
    """updaters.py:36-38,71-72 + train.py:101: replicas consume batch[i::n], gradients are
    SUMMED and alpha = lr/n.  The per-replica mean losses make the summed gradient n times the
    full-batch gradient for the batch-mean loss terms."""
    cfg = O.config_cpu()
    cfg.length, cfg.batch = 128, 4
    params = {k: v.double() for k, v in O.make_params(cfg).items()}
    inp = O.make_inputs(cfg)
    losses, total = O.parallel_update_grads(params, cfg, inp, 2)
    _, full, _ = O.three_loss_grads(params, cfg, *_args(inp, torch.float64))
    for k in ("decoder/proj2/W", "decoder/resnet/0/conv/W", "encoder/conv1/W", "vq/W"):
        assert torch.allclose(total[k], 2.0 * full[k], rtol=1e-9, atol=1e-13), k


def test_adam_and_ema_formulae():
    p = torch.tensor([1.0, -2.0], dtype=torch.float64)
    g = torch.tensor([0.5, 0.25], dtype=torch.float64)
    m, v = torch.zeros(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
    O.adam_step(p, g, m, v, 1, 1e-3)
    # t=1: m = 0.1 g, v = 0.001 g^2, lr = a*sqrt(0.001)/0.1, step = lr * m/(sqrt(v)+eps)
    want = torch.tensor([1.0, -2.0], dtype=torch.float64) - \
        1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * g) / (torch.sqrt(0.001 * g * g) + 1e-8)
    assert torch.allclose(p, want, atol=1e-15)
    tgt, ema = {"w": torch.tensor([2.0])}, {"w": torch.tensor([0.0])}
    O.weight_ema_update(tgt, ema, 0.9)
    assert torch.allclose(ema["w"], torch.tensor([1.8]))        # decay multiplies the TARGET


def test_mol_loss_against_float64_and_edges():
    cfg = O.config_mol()
    rng = np.random.default_rng(3)
    y = rng.normal(size=(2, 30, 50, 1)).astype(np.float32)
    y[:, 10:20] *= 40.0                                  # means spread over the +-127.5 range
    y[:, 20:30] = rng.uniform(2.0, 4.0, size=(2, 10, 50, 1))   # wide logistics: well conditioned
    y = torch.from_numpy(y)
    t = torch.from_numpy(rng.uniform(-1, 1, size=(2, 1, 50, 1)).astype(np.float32))
    t[0, 0, :5] = -1.0      # below -0.999: log_cdf_plus branch (modules.py:200-203)
    t[0, 0, 5:10] = 1.0     # above  0.999: log_one_minus_cdf_min branch (modules.py:208-211)
    a = O.calculate_logistic_loss(cfg, y, t)
    b = O.calculate_logistic_loss(cfg, y.double(), t.double())
    assert abs(float(a) - float(b)) < 1e-5 * abs(float(b))
    assert np.isfinite(float(a))
    lit = O.calculate_logistic_loss_numpy(cfg, y.numpy(), t.numpy())
    assert abs(lit - float(b)) < 1e-5 * abs(float(b))


def test_choice_from_uniform_matches_numpy_random_choice():
    rng = np.random.RandomState(5)
    p = rng.dirichlet(np.ones(256))
    st = np.random.RandomState(7)
    u = np.random.RandomState(7).random_sample(20)
    want = [st.choice(256, p=p) for _ in range(20)]
    assert [O.choice_from_uniform(p, ui) for ui in u] == want
