"""Host-side logic that needs no GPU: shapes, error behaviour, naming, optimiser arithmetic."""
import numpy as np
import pytest
import torch

import chainer_vq_vae_b200 as V
from chainer_vq_vae_b200 import functions as Fn
from oracle import vqvae_oracle as O


def test_conv_out_len_matches_chainer_formula():
    assert [Fn.conv_out_len(n, 4, 2, 1, 1) for n in (7681, 3840, 1920, 960, 480, 240)] == \
        [3840, 1920, 960, 480, 240, 120]                       # net.py:12-17 at T+1 = 7681
    assert Fn.conv_out_len(1024, 3, 1, 8, 4) == 1024 + 8      # modules.py:13-16 before the slice


def test_parameter_tree_names_match_the_reference():
    cfg = O.config_cpu()
    from helpers import build_model
    model = build_model(cfg, O.make_params(cfg), device="cpu")
    got = {n.lstrip("/") for n, _ in V.namedparams(model)}
    assert got == set(O.param_shapes(cfg))
    for name, p in V.namedparams(model):
        assert tuple(p.shape) == O.param_shapes(cfg)[name.lstrip("/")]


def test_cpu_tensors_fail_loudly():
    blk = V.ResidualBlock(3, 1, 32, 32, 32, 16, 0)
    with pytest.raises(V.VqwError):
        blk(torch.zeros(1, 32, 8, 1), torch.zeros(1, 16, 8, 1))
    with pytest.raises(V.VqwError):
        V.Encoder(8)(torch.zeros(1, 1, 65, 1))


def test_straight_through_type_checks():
    W = torch.zeros(4, 8)
    with pytest.raises(ValueError):
        V.straight_through(torch.zeros(2, 7, 5, 1), W)
    with pytest.raises(ValueError):
        V.straight_through(torch.zeros(2, 8), W)
    with pytest.raises(ValueError):
        V.straight_through(torch.zeros(2, 8, 5, 1), torch.zeros(4, 8, 1))
    with pytest.raises(TypeError):
        V.straight_through(torch.zeros(2, 8, 5, 1, dtype=torch.int64), W)


def test_generated_output_to_wav_matches_reference_formats(tmp_path):
    """generate.py:147-153: categorical output = mu-law indices -> MuLaw.itransform -> float32
    WAVE; mixture-of-logistics output is the waveform itself."""
    from scipy.io import wavfile
    rng = np.random.default_rng(0)
    out = rng.integers(0, 256, size=4000).astype(np.float64)       # generate.py:110: float array
    out[-1] = 0
    V.write_wav(tmp_path / "a.wav", torch.from_numpy(out), sr=16000, quantize=256)
    sr, wave = wavfile.read(tmp_path / "a.wav")
    assert sr == 16000 and wave.dtype == np.float32 and wave.shape == (4000,)
    assert np.array_equal(wave, O.MuLaw(256).itransform(out))
    raw = rng.uniform(-1, 1, size=1000)
    V.write_wav(tmp_path / "b.wav", raw, sr=22050, use_logistic=True)
    sr, wave = wavfile.read(tmp_path / "b.wav")
    assert sr == 22050 and np.array_equal(wave, raw.astype(np.float32))


def test_adam_matches_chainer_rule_and_bucket_views():
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    opt = V.Adam(1e-3).setup(lin)
    assert opt.bucket.flat.numel() == 18
    p0 = [p.detach().clone() for p in lin.parameters()]
    ms = [torch.zeros_like(p) for p in p0]
    vs = [torch.zeros_like(p) for p in p0]
    for step in range(1, 4):
        opt.bucket.zero()
        lin(torch.randn(7, 5)).pow(2).sum().backward()
        grads = [p.grad.clone() for p in lin.parameters()]
        assert all(p.grad.data_ptr() >= opt.bucket.flat.data_ptr() for p in lin.parameters())
        opt.update()
        for p, g, m, v in zip(p0, grads, ms, vs):
            O.adam_step(p, g, m, v, step, 1e-3)
        for p, q in zip(lin.parameters(), p0):
            assert torch.allclose(p, q, atol=1e-7)


def test_weight_ema_quirk_on_module():
    lin = torch.nn.Linear(2, 2)
    ema = V.ExponentialMovingAverage(lin, 0.9)
    with torch.no_grad():
        for p in ema.ema.parameters():
            p.zero_()
        w = lin.weight.clone()
    ema.update_average()
    assert torch.allclose(ema.ema.weight, 0.9 * w)            # decay multiplies the TARGET
    assert not any(p.requires_grad for p in ema.ema.parameters())


def test_concat_examples_and_split():
    ex = [(np.full((1, 9, 1), i, np.float32), np.arange(8, dtype=np.int32), np.int32(i),
           np.zeros((8, 1), np.int32)) for i in range(6)]
    a, b, c, d = V.updaters.concat_examples(ex, None)
    assert a.shape == (6, 1, 9, 1) and b.shape == (6, 8) and c.shape == (6,) and d.shape == (6, 8, 1)
    assert [int(e[2]) for e in V.VQVAE_ParallelUpdater.split(ex, 1, 2)] == [1, 3, 5]   # batch[i::n]


def test_mulaw_roundtrip():
    m = V.MuLaw(256)
    q = m.transform(np.linspace(-1, 1, 1001).astype(np.float32))
    assert q.min() == 0 and q.max() == 255 and q.dtype == np.int32
    x = m.itransform(q)
    assert np.all(np.diff(x) >= 0) and abs(x[500]) < 0.01


def test_chainer_snapshot_roundtrip_and_reference_keys(tmp_path):
    from helpers import build_model
    from chainer_vq_vae_b200.snapshot import load_chainer_snapshot, save_chainer_snapshot, PREFIX
    cfg = O.config_cpu()
    params = O.make_params(cfg)
    # a snapshot as the reference would write it (EMA wrapper on): decoder/target + decoder/ema
    ref = {}
    for k, v in params.items():
        if k.startswith("decoder/"):
            ref[PREFIX + "decoder/target/" + k[len("decoder/"):]] = v.numpy()
            ref[PREFIX + "decoder/ema/" + k[len("decoder/"):]] = (v * 0.5).numpy()
        else:
            ref[PREFIX + k] = v.numpy()
    ref["updater/optimizer:main/t"] = np.array(7)
    path = tmp_path / "snapshot_iter_7.npz"
    np.savez(path, **ref)
    # generate.py:70-73: a bare WaveNet decoder loads the EMA sub-tree
    plain = build_model(cfg, O.make_params(cfg, seed=99), device="cpu")
    missing, unexpected = load_chainer_snapshot(str(path), plain, use_ema=True)
    assert not missing and not unexpected
    assert torch.equal(plain.decoder.proj2.W, params["decoder/proj2/W"] * 0.5)
    assert torch.equal(plain.encoder.conv3.W, params["encoder/conv3/W"])
    load_chainer_snapshot(str(path), plain, use_ema=False)                 # generate.py:74-76
    assert torch.equal(plain.decoder.proj2.W, params["decoder/proj2/W"])
    # a wrapped decoder gets both copies; save -> load round trip
    wrapped = build_model(cfg, O.make_params(cfg, seed=98), device="cpu", ema_decay=0.9999)
    load_chainer_snapshot(str(path), wrapped)
    assert torch.equal(wrapped.decoder.ema.embed.W, params["decoder/embed/W"] * 0.5)
    assert torch.equal(wrapped.decoder.target.embed.W, params["decoder/embed/W"])
    p2 = tmp_path / "out.npz"
    save_chainer_snapshot(str(p2), wrapped)
    again = build_model(cfg, O.make_params(cfg, seed=97), device="cpu", ema_decay=0.9999)
    assert load_chainer_snapshot(str(p2), again) == ([], [])
    for (n, a), (_, b) in zip(wrapped.named_parameters(), again.named_parameters()):
        assert torch.equal(a, b), n


def test_snapshot_keeps_optimizer_state_and_extensionless_names(tmp_path):
    """`--resume` (train.py:152-153): Adam's moments, step count and the iteration survive a
    save/load; Chainer-style extensionless names are written verbatim (no '.npz' appended)."""
    import chainer_vq_vae_b200 as V
    from helpers import build_model
    cfg = O.config_cpu()
    model = build_model(cfg, O.make_params(cfg), device="cpu")
    opt = V.Adam(1e-3).setup(model)
    opt.flat_m.normal_()
    opt.flat_v.uniform_()
    opt.t = 12
    path = tmp_path / "snapshot_iter_12"
    V.save_chainer_snapshot(str(path), model, optimizer=opt, iteration=12)
    assert path.exists() and not (tmp_path / "snapshot_iter_12.npz").exists()
    model2 = build_model(cfg, O.make_params(cfg, seed=5), device="cpu")
    opt2 = V.Adam(1e-3).setup(model2)
    assert V.load_chainer_snapshot(str(path), model2) == ([], [])
    assert V.load_optimizer_state(str(path), model2, opt2) == 12
    assert opt2.t == 12 and torch.equal(opt2.flat_m, opt.flat_m) and torch.equal(opt2.flat_v, opt.flat_v)
    assert torch.equal(opt2.flat_p, opt.flat_p)


def test_overlapped_allreduce_segments_cover_the_stack_in_completion_order():
    """The stack backward finishes blocks from the last to the first; the overlapped all-reduce
    (updaters._OverlappedReduce) reduces contiguous groups of blocks as they complete."""
    from chainer_vq_vae_b200.updaters import block_segments
    assert block_segments(20, 4) == [(15, 20), (10, 15), (5, 10), (0, 5)]
    assert block_segments(3, 8) == [(2, 3), (1, 2), (0, 1)]
    assert block_segments(30, 1) == [(0, 30)]
    for n, s in ((20, 4), (30, 4), (7, 3), (40, 6)):
        segs = block_segments(n, s)
        covered = sorted(b for a, e in segs for b in range(a, e))
        assert covered == list(range(n))                       # every block exactly once
        assert all(x[0] >= y[1] for x, y in zip(segs, segs[1:]))   # highest blocks first
