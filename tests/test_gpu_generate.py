"""Persistent generation kernel against the oracle's restatement of generate.py:104-145 with
the reference's concat-shift queues."""
import numpy as np
import pytest
import torch

import chainer_vq_vae_b200 as V
from chainer_vq_vae_b200.generate import generate_utterance
from oracle import vqvae_oracle as O
from helpers import TOL, build_model, rel_err

pytestmark = pytest.mark.gpu


def _case(cfg, length):
    cfg.length = length
    cfg.batch = 1
    params = O.make_params(cfg)
    inp = O.make_inputs(cfg)
    return params, inp


@pytest.mark.parametrize("fs,n_loop", [(3, 1), (2, 2)])
def test_generate_matches_oracle(fs, n_loop):
    cfg = O.config_cpu()
    cfg.filter_size, cfg.n_loop = fs, n_loop
    params, inp = _case(cfg, 256)
    steps = 120
    u = np.random.default_rng(0).uniform(size=steps)
    out_o, logits_o = O.generate_loop(params, cfg, inp["x_enc"], inp["speaker"], u, n_steps=steps,
                                      return_logits=True)
    model = build_model(cfg, params).eval()
    with torch.no_grad():
        x_enc = torch.from_numpy(inp["x_enc"]).cuda()
        z = model.encoder(x_enc)
        e = model.vq(z)
        cond = model.condition_embed(e, torch.from_numpy(inp["speaker"]).cuda())
        out_g, logits_g = generate_utterance(model.decoder, cond, u, n_steps=steps,
                                             return_logits=True)
    out_g = out_g.cpu().numpy()
    assert out_g.shape == out_o.shape and out_g[-1] == 0          # generate.py:110-112
    assert np.array_equal(out_g[:steps], out_o[:steps]), "sampled indices must be identical"
    assert rel_err(logits_g, logits_o) < TOL


def test_generate_teacher_forced_and_stepwise_api():
    cfg = O.config_cpu()
    params, inp = _case(cfg, 192)
    model = build_model(cfg, params).eval()
    dec = model.decoder
    steps = 40
    with torch.no_grad():
        x_enc = torch.from_numpy(inp["x_enc"]).cuda()
        cond = model.condition_embed(model.vq(model.encoder(x_enc)),
                                     torch.from_numpy(inp["speaker"]).cuda())
        q = inp["quantized"][0]
        # teacher forcing: feed the ground-truth sample after every step
        forced = q[1:steps + 1].astype(np.int32)
        u = np.full(steps, 0.5)
        _, logits = generate_utterance(dec, cond, u, n_steps=steps, forced=forced, return_logits=True)
        # full-sequence forward on the same inputs: column i sees x[i-1], x[i] with x[0]=zeros...
        # the stepwise API reproduces WaveNet.generate (modules.py:245-255) exactly:
        dec.initialize(1)
        x = torch.zeros(1, cfg.input_dim, 1, 1, device="cuda")
        outs = []
        for i in range(steps):
            o = dec.generate(x, cond[:, :, i:i + 1])
            outs.append(o[0, :, 0, 0])
            x = torch.zeros(1, cfg.input_dim, 1, 1, device="cuda")
            x[0, int(forced[i])] = 1
        outs = torch.stack(outs)
    assert rel_err(outs, logits) < 1e-6
    # and both equal the oracle's queue-based generator
    gen = O.WaveNetGenerator(O.sub(params, "decoder/"), cfg, 1)
    xo = torch.zeros(1, cfg.input_dim, 1, 1)
    ref = []
    with torch.no_grad():
        for i in range(steps):
            o = gen.generate(xo, cond[:, :, i:i + 1].cpu())
            ref.append(o[0, :, 0, 0])
            xo = torch.zeros(1, cfg.input_dim, 1, 1)
            xo[0, int(forced[i])] = 1
    assert rel_err(logits, torch.stack(ref)) < TOL


def test_generate_b200_channels_smoke():
    """512/512/256 channels, n_loop=2 x n_layer=10: a few steps against the oracle."""
    cfg = O.config_b200()
    params, inp = _case(cfg, 128)
    steps = 6
    u = np.random.default_rng(1).uniform(size=steps)
    out_o, logits_o = O.generate_loop(params, cfg, inp["x_enc"], inp["speaker"], u, n_steps=steps,
                                      return_logits=True)
    model = build_model(cfg, params).eval()
    with torch.no_grad():
        cond = model.condition_embed(model.vq(model.encoder(torch.from_numpy(inp["x_enc"]).cuda())),
                                     torch.from_numpy(inp["speaker"]).cuda())
        out_g, logits_g = generate_utterance(model.decoder, cond, u, n_steps=steps, return_logits=True)
    assert rel_err(logits_g, logits_o) < TOL
    assert np.array_equal(out_g.cpu().numpy()[:steps], out_o[:steps])


def _mol_case(length):
    cfg = O.config_cpu()
    cfg.use_logistic, cfg.input_dim = True, 1
    params, inp = _case(cfg, length)
    model = build_model(cfg, params).eval()
    with torch.no_grad():
        cond = model.condition_embed(model.vq(model.encoder(torch.from_numpy(inp["x_enc"]).cuda())),
                                     torch.from_numpy(inp["speaker"]).cuda())
    return cfg, params, inp, model, cond


def test_generate_mixture_of_logistics_matches_oracle():
    """generate.py:116-137 in the persistent kernel: scalar input, one logistic draw from every
    mixture component mixed by the softmax weights, / 127.5, clip.  Free running from the same
    uniforms (the fed-back value is continuous, so differences of the last float32 bit are
    carried forward: values within 1e-3 of their [-1, 1] range, logits within 1e-3)."""
    cfg, params, inp, model, cond = _mol_case(192)
    steps, nr = 60, cfg.n_mixture // 3
    u = np.random.default_rng(3).uniform(0.02, 0.98, size=(steps, nr))
    out_o, logits_o = O.generate_loop(params, cfg, inp["x_enc"], inp["speaker"], u, n_steps=steps,
                                      return_logits=True)
    with torch.no_grad():
        out_g, logits_g = generate_utterance(model.decoder, cond, u, n_steps=steps, return_logits=True)
    out_g = out_g.cpu().numpy()
    assert out_g.shape == out_o.shape and out_g[-1] == 0
    assert np.abs(out_g[:steps]).max() <= 1.0 and np.abs(out_g[:steps]).max() > 0
    err_v = np.abs(out_g[:steps] - out_o[:steps]).max()
    print(f"MoL generation: max value diff {err_v:.2e}, logits rel err {rel_err(logits_g, logits_o):.2e}")
    assert err_v < 1e-3
    assert rel_err(logits_g, logits_o) < TOL


def test_generate_mixture_of_logistics_teacher_forced_and_stepwise():
    cfg, params, inp, model, cond = _mol_case(160)
    dec = model.decoder
    steps, nr = 32, cfg.n_mixture // 3
    forced = inp["x_enc"][0, 0, 1:steps + 1, 0].astype(np.float32)      # raw waveform values
    u = np.full((steps, nr), 0.5)
    with torch.no_grad():
        _, logits = generate_utterance(dec, cond, u, n_steps=steps, forced=forced, return_logits=True)
        dec.initialize(1)
        x = torch.zeros(1, 1, 1, 1, device="cuda")
        outs = []
        for i in range(steps):
            outs.append(dec.generate(x, cond[:, :, i:i + 1])[0, :, 0, 0])
            x = torch.full((1, 1, 1, 1), float(forced[i]), device="cuda")
        outs = torch.stack(outs)
    assert rel_err(outs, logits) < 1e-6
    gen = O.WaveNetGenerator(O.sub(params, "decoder/"), cfg, 1)
    xo = torch.zeros(1, 1, 1, 1)
    ref = []
    with torch.no_grad():
        for i in range(steps):
            ref.append(gen.generate(xo, cond[:, :, i:i + 1].cpu())[0, :, 0, 0])
            xo = torch.full((1, 1, 1, 1), float(forced[i]))
    assert rel_err(logits, torch.stack(ref)) < TOL


# ---------------------------------------------------------------------------------------
# BASELINE.json configs[4] at its own size: n_loop=4 x n_layer=10, 512/512/256, fs=3
# ---------------------------------------------------------------------------------------
def _gen_case(length):
    cfg = O.config_gen()
    params, inp = _case(cfg, length)
    model = build_model(cfg, params).eval()
    with torch.no_grad():
        cond = model.condition_embed(model.vq(model.encoder(torch.from_numpy(inp["x_enc"]).cuda())),
                                     torch.from_numpy(inp["speaker"]).cuda())
    return cfg, params, inp, model, cond


def test_generate_config4_teacher_forced_past_the_dilation_512_ring_wrap():
    """40 blocks, dilations 1..512 four times.  The dilation-512 queue holds 1025 columns
    (modules.py:59-62), so the kernel's ring buffers wrap for the first time after step 1025:
    1200 teacher-forced steps against the oracle's concat-shift generator (modules.py:58-74,
    98-110, 232-255), every step's logits within 1e-3."""
    steps = 1200
    cfg, params, inp, model, cond = _gen_case(1280)
    forced = inp["quantized"][0, 1:steps + 1].astype(np.int32)
    u = np.full(steps, 0.5)
    with torch.no_grad():
        _, logits = generate_utterance(model.decoder, cond, u, n_steps=steps, forced=forced,
                                       return_logits=True)
    gen = O.WaveNetGenerator(O.sub(params, "decoder/"), cfg, 1)
    cond_c = cond.cpu()
    xo = torch.zeros(1, cfg.input_dim, 1, 1)
    ref = []
    torch.set_num_threads(max(1, __import__("os").cpu_count() or 2))
    with torch.no_grad():
        for i in range(steps):
            ref.append(gen.generate(xo, cond_c[:, :, i:i + 1])[0, :, 0, 0].clone())
            xo = torch.zeros(1, cfg.input_dim, 1, 1)
            xo[0, int(forced[i])] = 1
    ref = torch.stack(ref)
    worst = max(rel_err(logits[a:b], ref[a:b]) for a, b in ((0, 400), (400, 1000), (1000, steps)))
    tail = rel_err(logits[1030:], ref[1030:])       # steps that read wrapped ring slots
    print(f"configs[4] teacher-forced {steps} steps: worst rel err {worst:.2e}, after the wrap {tail:.2e}")
    assert worst < TOL and tail < TOL


def test_generate_config4_free_running_identical_indices():
    """300 free-running categorical steps at the configs[4] size: the sampled mu-law indices are
    identical to the oracle's generate.py loop (numpy.random.choice rule from the same uniforms)."""
    steps = 300
    cfg, params, inp, model, cond = _gen_case(320)
    u = np.random.default_rng(4).uniform(size=steps)
    out_o, logits_o = O.generate_loop(params, cfg, inp["x_enc"], inp["speaker"], u, n_steps=steps,
                                      return_logits=True)
    with torch.no_grad():
        out_g, logits_g = generate_utterance(model.decoder, cond, u, n_steps=steps, return_logits=True)
    assert np.array_equal(out_g.cpu().numpy()[:steps], out_o[:steps])
    assert rel_err(logits_g, logits_o) < TOL


def test_vae_generate_entry_point_and_wav(tmp_path):
    """VAE.generate(raw, speaker, use_ema) (the convenience entry of models.py:74-105) draws the
    same utterance as the oracle's generate.py loop; use_ema selects the decoder copy like
    generate.py:70-76; the result goes through MuLaw.itransform into a float WAVE file."""
    from scipy.io import wavfile
    cfg = O.config_cpu()
    params, inp = _case(cfg, 256)
    model = build_model(cfg, params, ema_decay=0.999).eval()
    with torch.no_grad():                                   # make the two copies differ
        for p in model.decoder.target.parameters():
            p.mul_(1.01)
    steps = 100
    u = np.random.default_rng(0).uniform(size=steps)
    out_o = O.generate_loop(params, cfg, inp["x_enc"], inp["speaker"], u, n_steps=steps)
    raw = torch.from_numpy(inp["x_enc"]).cuda()
    spk = torch.from_numpy(inp["speaker"]).cuda()
    out_ema = model.generate(raw, spk, use_ema=True, uniforms=u, n_steps=steps)
    out_tgt = model.generate(raw, spk, use_ema=False, uniforms=u, n_steps=steps)
    assert np.array_equal(out_ema.cpu().numpy()[:steps], out_o[:steps])
    assert not np.array_equal(out_tgt.cpu().numpy()[:steps], out_o[:steps])
    V.write_wav(tmp_path / "gen.wav", out_ema, sr=16000, quantize=cfg.quantize)
    sr, wave = wavfile.read(tmp_path / "gen.wav")
    assert sr == 16000 and np.array_equal(wave, O.MuLaw(256).itransform(out_o))
