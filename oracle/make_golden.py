"""Generate tests/golden/*.npz by running the REFERENCE'S OWN model code.

Run in the build container (needs /root/reference):   python oracle/make_golden.py

`net.py`, `utils.py` and `WaveNet/modules.py` are imported unmodified from /root/reference on
top of `oracle/chainer_shim` (a NumPy restatement of the Chainer primitives they call; Chainer
itself is not installable here).  What is frozen is therefore the reference's composition --
layer wiring, padding/slicing, split order, the VQ arithmetic of utils.py:189-211, the
three-loss definition, the MoL loss, the queue logic of initialize()/generate() -- evaluated
with seeded weights and inputs from oracle/vqvae_oracle.py.  tests/test_oracle_golden.py then
pins the oracle (torch-CPU restatement) to these vectors, and the GPU tests pin the CUDA path
to the oracle.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("VQW_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "chainer_shim"))
sys.path.insert(0, REF)

import chainer  # noqa: E402  (the shim)
import chainer.functions as F  # noqa: E402
import net as ref_net  # noqa: E402  (reference)
from WaveNet import WaveNet as RefWaveNet  # noqa: E402  (reference)
from utils import ExponentialMovingAverage as RefEMA, MuLaw as RefMuLaw  # noqa: E402
from oracle import vqvae_oracle as O  # noqa: E402


def build_reference(cfg, params, ema=None):
    """train.py:76-98 wiring."""
    encoder = ref_net.Encoder(cfg.d)
    wavenet = RefWaveNet(cfg.n_loop, cfg.n_layer, cfg.filter_size, cfg.input_dim,
                         cfg.residual_channels, cfg.dilated_channels, cfg.skip_channels,
                         cfg.quantize, cfg.use_logistic, cfg.n_mixture, cfg.log_scale_min,
                         cfg.condition_dim, 0)
    cond = ref_net.ConditionEmbed(cfg.n_speaker, cfg.global_condition_dim,
                                  cfg.local_condition_dim)
    decoder = RefEMA(wavenet, ema) if ema else wavenet
    loss_fun = wavenet.calculate_logistic_loss if cfg.use_logistic else F.softmax_cross_entropy
    model = ref_net.VAE(encoder, decoder, cond, cfg.d, cfg.k, cfg.beta, loss_fun)
    # lazy-shaped links (net.py:34-43 use in_channels=None): materialise, then load the weights
    for i, cin in zip(range(1, 6), [cfg.d] + [cfg.local_condition_dim] * 4):
        getattr(cond, "local_embed%d" % i)._initialize_params(cin)
    own = dict(model.namedparams())
    for name, val in params.items():
        key = "/" + name
        if ema and name.startswith("decoder/"):
            for sub in ("target", "ema"):
                own["/decoder/" + sub + "/" + name[len("decoder/"):]].array = val.numpy().copy()
        else:
            assert own[key].array is None or own[key].array.shape == tuple(val.shape), key
            own[key].array = val.numpy().copy()
    return model, wavenet


def run_case(name, cfg, gen_steps=24, ema=None):
    params = O.make_params(cfg)
    inp = O.make_inputs(cfg)
    model, wavenet = build_reference(cfg, params, ema)
    x_enc, x_dec, spk, t = inp["x_enc"], inp["x_dec"], inp["speaker"], inp["t"]
    out = {}
    # the pieces, in net.py:81-86 order
    z = model.encoder(x_enc)
    e = model.vq(z)
    # utils.py:202-203 on the same z, W (StraightThrough keeps indexes only on the node)
    W = model.vq.W.array
    xs = np.broadcast_to(np.expand_dims(z.array, 1), (z.shape[0], W.shape[0]) + z.shape[1:])
    Wb = np.broadcast_to(np.reshape(W, (1,) + W.shape + (1, 1)), xs.shape)
    out["indexes"] = np.argmin(np.sum((xs - Wb) ** 2, axis=2), axis=1).astype(np.int32)
    condition = model.condition_embed(e, spk)
    y = wavenet(chainer.Variable(x_dec), condition)
    out.update(z=z.array, e=e.array, condition=condition.array, y=y.array)
    # the whole VAE.__call__ (net.py:79-96)
    l1, l2, l3 = model(x_enc, x_dec, spk, t)
    out["losses"] = np.array([float(l1.array), float(l2.array), float(l3.array)], np.float64)
    if ema:
        out["ema_embed_W"] = dict(model.namedparams())["/decoder/ema/embed/W"].array.copy()
    # initialize()/generate() queue path (modules.py:232-255), teacher-forced with x_dec
    wavenet.initialize(1)
    gen = []
    with chainer.using_config("train", False):
        for i in range(gen_steps):
            o = wavenet.generate(chainer.Variable(x_dec[:1, :, i:i + 1]),
                                 chainer.Variable(condition.array[:1, :, i:i + 1]))
            gen.append(o.array[0, :, 0, 0].copy())
    out["generate"] = np.stack(gen)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"),
                        **{k: np.asarray(v) for k, v in out.items()})
    print(name, "losses", out["losses"], "y", out["y"].shape)


def main():
    np.random.seed(0)
    cfg = O.config_cpu()
    cfg.length = 256                       # keeps the fixture small (y is 2x256x256 f32)
    run_case("ref_cpu_config_T256", cfg)
    cfg2 = O.config_cpu()
    cfg2.length = 256
    cfg2.filter_size = 2                   # the reference's default filter_size (params.py:31)
    cfg2.n_loop = 2
    run_case("ref_fs2_nloop2_T256", cfg2, ema=0.9999)
    cfg3 = O.config_cpu()
    cfg3.length = 256
    cfg3.use_logistic, cfg3.input_dim = True, 1      # BASELINE.json configs[3] in miniature
    run_case("ref_mol_T256", cfg3)
    # mu-law (utils.py:12-29) and numpy.random.choice's uniform->index map
    x = np.linspace(-1, 1, 4001).astype(np.float32)
    m = RefMuLaw(256)
    q = m.transform(x)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_mulaw.npz"), x=x, q=q,
                        inv=m.itransform(np.arange(256)))
    print("mulaw", q.min(), q.max())


if __name__ == "__main__":
    main()
