"""GPU parity tests of the individual C-ABI kernels against the oracle (run on the B200)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import chainer_vq_vae_b200 as V
from oracle import vqvae_oracle as O
from helpers import TOL, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


# ---------------------------------------------------------------- VQ -----------------
@pytest.mark.parametrize("B,d,T,k", [(2, 64, 16, 128), (16, 64, 120, 512), (1, 64, 375, 512),
                                      (3, 7, 5, 33), (2, 512, 9, 128)])
def test_vq_forward_bit_exact(B, d, T, k):
    rng = np.random.default_rng(B * 1000 + T)
    z = rng.normal(size=(B, d, T, 1)).astype(np.float32)
    W = rng.normal(0, 1 / np.sqrt(d), size=(k, d)).astype(np.float32)
    # planted ties: duplicated codebook rows (first occurrence must win) and exact hits
    W[k // 2] = W[3]
    W[k - 1] = W[0]
    z[0, :, 0, 0] = W[3]
    z[-1, :, T - 1, 0] = W[0]
    ref = O.vq_indexes(z, W)
    e, idx, count, zsum, sqerr = V.vq_lookup(torch.from_numpy(z).to(DEV), torch.from_numpy(W).to(DEV),
                                             stats=True)
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (B, T, 1)
    assert np.array_equal(idx.cpu().numpy(), ref), "VQ indices must be bit-identical"
    assert idx[0, 0, 0] == 3 and idx[-1, T - 1, 0] == 0
    e_ref = np.transpose(W[ref], (0, 3, 1, 2))
    assert np.array_equal(e.cpu().numpy(), e_ref)
    assert np.array_equal(count.cpu().numpy(), np.bincount(ref.ravel(), minlength=k).astype(np.float32))
    zs = np.zeros((k, d), np.float64)
    np.add.at(zs, ref.ravel(), np.transpose(z[..., 0], (0, 2, 1)).reshape(-1, d))
    assert rel_err(zsum, zs) < 1e-5
    assert abs(float(sqerr) - float(((z.astype(np.float64) - e_ref) ** 2).sum())) < 1e-6 * z.size


def test_vq_small_literal_numpy_and_3d():
    rng = np.random.default_rng(5)
    z = rng.normal(size=(2, 8, 6)).astype(np.float32)          # 3-D input (utils.py:170-171)
    W = rng.normal(size=(16, 8)).astype(np.float32)
    ref = O.vq_indexes_numpy(z, W)
    e, idx, *_ = V.vq_lookup(torch.from_numpy(z).to(DEV), torch.from_numpy(W).to(DEV))
    assert np.array_equal(idx.cpu().numpy(), ref)
    assert np.array_equal(e.cpu().numpy(), np.transpose(W[ref], (0, 2, 1)))


def test_vq_empty_and_errors():
    W = torch.zeros(4, 8, device=DEV)
    e, idx, *_ = V.vq_lookup(torch.zeros(0, 8, 5, 1, device=DEV), W)
    assert e.shape == (0, 8, 5, 1) and idx.shape == (0, 5, 1)
    with pytest.raises(ValueError):
        V.straight_through(torch.zeros(2, 7, 5, 1, device=DEV), W)       # channel mismatch
    with pytest.raises(ValueError):
        V.straight_through(torch.zeros(2, 8, device=DEV), W)             # ndim 2
    with pytest.raises(TypeError):
        V.straight_through(torch.zeros(2, 8, 5, 1, device=DEV, dtype=torch.int32), W)


def test_vq_straight_through_gradients():
    rng = np.random.default_rng(11)
    B, d, T, k = 4, 64, 30, 128
    z = torch.from_numpy(rng.normal(size=(B, d, T, 1)).astype(np.float32))
    W = torch.from_numpy(rng.normal(0, 0.125, size=(k, d)).astype(np.float32))
    gy = torch.from_numpy(rng.normal(size=(B, d, T, 1)).astype(np.float32))
    zo = z.clone().requires_grad_(True)
    Wo = W.clone().requires_grad_(True)
    O.straight_through(zo, Wo).backward(gy)
    zg = z.to(DEV).requires_grad_(True)
    Wg = W.to(DEV).requires_grad_(True)
    V.straight_through(zg, Wg).backward(gy.to(DEV))
    assert torch.equal(zg.grad.cpu(), gy), "gx must equal gy bitwise (utils.py:218-219)"
    assert torch.equal(Wg.grad.cpu(), Wo.grad), "float64-accumulated gW must match exactly"


# ---------------------------------------------------------------- generic conv -------
@pytest.mark.parametrize("cin,cout,k,stride,pad,dil,T,relu,relu_in", [
    (1, 64, 4, 2, 1, 1, 1025, True, False),     # Encoder conv1 (net.py:12)
    (64, 64, 4, 2, 1, 1, 512, False, False),    # Encoder conv6
    (64, 64, 3, 1, 4, 4, 16, True, False),      # ConditionEmbed local_embed3 (net.py:38-39)
    (64, 64, 3, 1, 16, 16, 16, True, False),    # dilation wider than the signal
    (32, 256, 1, 1, 0, 1, 300, False, True),    # proj (modules.py:158-159)
    (256, 32, 2, 1, 1, 1, 77, False, False),    # embed (modules.py:127-128)
    (5, 3, 3, 2, 2, 3, 41, True, True),         # ragged everything
])
def test_conv_forward_backward(cin, cout, k, stride, pad, dil, T, relu, relu_in):
    rng = np.random.default_rng(cin * 7 + T)
    B = 3
    x = torch.from_numpy(rng.normal(size=(B, cin, T, 1)).astype(np.float32))
    W = torch.from_numpy(rng.normal(0, 1 / np.sqrt(cin * k), size=(cout, cin, k, 1)).astype(np.float32))
    b = torch.from_numpy(rng.normal(0, 0.1, size=(cout,)).astype(np.float32))
    xo, Wo, bo = (t.double().clone().requires_grad_(True) for t in (x, W, b))
    yo = O.conv2d(F.relu(xo) if relu_in else xo, Wo, bo, stride=stride, pad=pad, dilate=dil)
    if relu:
        yo = F.relu(yo)
    gy = torch.from_numpy(rng.normal(size=tuple(yo.shape)).astype(np.float32))
    yo.backward(gy.double())
    xg, Wg, bg = (t.to(DEV).requires_grad_(True) for t in (x, W, b))
    yg = V.conv(xg, Wg, bg, stride, pad, dil, relu, None, relu_in)
    assert tuple(yg.shape) == tuple(yo.shape)
    yg.backward(gy.to(DEV))
    assert rel_err(yg, yo) < 1e-5
    assert rel_err(xg.grad, xo.grad) < 1e-5
    assert rel_err(Wg.grad, Wo.grad) < 1e-4
    assert rel_err(bg.grad, bo.grad) < 1e-4


# ---------------------------------------------------------------- embed --------------
def test_embed_gather_matches_one_hot_conv():
    rng = np.random.default_rng(3)
    B, T, Cr, Q = 3, 301, 32, 256
    q = rng.integers(0, Q, size=(B, T)).astype(np.int32)
    W = torch.from_numpy(rng.normal(0, 0.05, size=(Cr, Q, 2, 1)).astype(np.float32))
    b = torch.from_numpy(rng.normal(0, 0.01, size=(Cr,)).astype(np.float32))
    onehot = torch.from_numpy(np.identity(Q, dtype=np.float32)[q].transpose(0, 2, 1)[..., None].copy())
    Wo, bo = W.double().clone().requires_grad_(True), b.double().clone().requires_grad_(True)
    yo = O.conv2d(onehot.double(), Wo, bo, pad=1)[:, :, :T]                 # modules.py:151-152
    gy = torch.from_numpy(rng.normal(size=tuple(yo.shape)).astype(np.float32))
    yo.backward(gy.double())
    Wg, bg = W.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    yg = V.embed_gather(torch.from_numpy(q).to(DEV), Wg, bg)
    yg.backward(gy.to(DEV))
    assert rel_err(yg, yo) < 1e-6
    assert rel_err(Wg.grad, Wo.grad) < 1e-5
    assert rel_err(bg.grad, bo.grad) < 1e-5


# ---------------------------------------------------------------- residual block -----
def _block_params(rng, Cr, Cd, Cs, Cc, fs):
    def w(*shape):
        fan = int(np.prod(shape[1:]))
        return torch.from_numpy(rng.normal(0, 1 / np.sqrt(fan), size=shape).astype(np.float32))

    def bias(n):
        return torch.from_numpy(rng.normal(0, 0.05, size=(n,)).astype(np.float32))
    return {"conv/W": w(Cd, Cr, fs, 1), "conv/b": bias(Cd), "condition_proj/W": w(Cd, Cc, 1, 1),
            "condition_proj/b": bias(Cd), "res/W": w(Cr, Cd // 2, 1, 1), "res/b": bias(Cr),
            "skip/W": w(Cs, Cd // 2, 1, 1), "skip/b": bias(Cs)}


ORDER = ["conv/W", "conv/b", "condition_proj/W", "condition_proj/b", "res/W", "res/b", "skip/W",
         "skip/b"]


@pytest.mark.parametrize("Cr,Cd,Cs,Cc,fs,dil,T", [
    (32, 32, 32, 192, 3, 1, 1024),     # CPU config block (BASELINE.json configs[0])
    (32, 32, 32, 192, 3, 8, 100),      # ragged T, dilation reaching past the start
    (32, 32, 32, 192, 2, 4, 64),       # reference default filter_size (params.py:31)
    (24, 40, 12, 10, 3, 2, 50),        # nothing a multiple of anything
    (64, 128, 48, 192, 3, 16, 96),     # PPW=8
    (128, 256, 64, 192, 3, 2, 64),     # PPW=16
    (512, 512, 256, 192, 3, 4, 64),    # the B200 config's channel counts (fp32 mode)
])
def test_resblock_forward_backward_fp32(Cr, Cd, Cs, Cc, fs, dil, T):
    rng = np.random.default_rng(Cr + T)
    B = 2
    p = _block_params(rng, Cr, Cd, Cs, Cc, fs)
    x = torch.from_numpy(rng.normal(size=(B, Cr, T, 1)).astype(np.float32))
    c = torch.from_numpy(rng.normal(size=(B, Cc, T, 1)).astype(np.float32))
    po = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    xo, co = x.double().clone().requires_grad_(True), c.double().clone().requires_grad_(True)
    ro, so = O.residual_block_forward(po, xo, co, fs, dil)
    gr = torch.from_numpy(rng.normal(size=tuple(ro.shape)).astype(np.float32))
    gs = torch.from_numpy(rng.normal(size=tuple(so.shape)).astype(np.float32))
    (ro * gr.double()).sum().backward(retain_graph=True)
    (so * gs.double()).sum().backward()

    blk = V.ResidualBlock(fs, dil, Cr, Cd, Cs, Cc, 0).to(DEV)
    with torch.no_grad():
        for name, t in zip(ORDER, blk.weights()):
            t.copy_(p[name])
    xg, cg = x.to(DEV).requires_grad_(True), c.to(DEV).requires_grad_(True)
    rg, sg = blk(xg, cg)
    assert rel_err(rg, ro) < 1e-5 and rel_err(sg, so) < 1e-5
    ((rg * gr.to(DEV)).sum() + (sg * gs.to(DEV)).sum()).backward()
    assert rel_err(xg.grad, xo.grad) < 1e-4
    assert rel_err(cg.grad, co.grad) < 1e-4
    for name, t in zip(ORDER, blk.weights()):
        assert rel_err(t.grad, po[name].grad) < 2e-4, name


def test_resblock_rejects_bad_arguments():
    blk = V.ResidualBlock(3, 1, 32, 32, 32, 16, 0).to(DEV)
    x = torch.zeros(2, 32, 64, 1, device=DEV)
    with pytest.raises(ValueError):
        blk(x, torch.zeros(2, 16, 63, 1, device=DEV))           # length mismatch
    with pytest.raises(V.VqwError):
        blk(x.cpu(), torch.zeros(2, 16, 64, 1))                  # CPU tensors: no fallback
    with pytest.raises(NotImplementedError):
        V.ResidualBlock(3, 1, 32, 32, 32, 16, 0.05)


def test_mol_loss_kernel_matches_literal_and_autograd():
    """vqw_mol_loss against (a) the NumPy-literal float32 evaluation of modules.py:169-230 on a
    well-conditioned case (wide logistics: no cdf cancellation) at 1e-5, including the +-0.999
    edge branches and the log_scale floor, and (b) torch autograd of the float64 oracle formula
    for the gradient."""
    from chainer_vq_vae_b200.losses import logistic_loss
    cfg = O.config_cpu()
    cfg.use_logistic, cfg.input_dim = True, 1
    rng = np.random.default_rng(5)
    B, nr, T = 3, 10, 333
    y = np.concatenate([rng.normal(0, 1, (B, nr, T, 1)), rng.normal(0, 40, (B, nr, T, 1)),
                        rng.normal(3.5, 0.3, (B, nr, T, 1))], axis=1).astype(np.float32)
    y[:, 2 * nr:, :40] = -50.0                                   # below log_scale_min = -40
    t = rng.uniform(-1, 1, (B, 1, T, 1)).astype(np.float32)
    t[:, :, :7] = -1.0                                           # x < -0.999 * 127.5
    t[:, :, 7:15] = 1.0                                          # x >  0.999 * 127.5
    want = O.calculate_logistic_loss_numpy(cfg, y, t)
    yg = torch.from_numpy(y).cuda().requires_grad_(True)
    loss = logistic_loss(yg, torch.from_numpy(t).cuda(), cfg.quantize, cfg.log_scale_min)
    assert abs(float(loss) - want) <= 1e-5 * abs(want), (float(loss), want)
    loss.backward()
    y64 = torch.from_numpy(y).double().requires_grad_(True)
    O.calculate_logistic_loss(cfg, y64, torch.from_numpy(t).double()).backward()
    assert rel_err(yg.grad, y64.grad) < 1e-4
    # every group of channels on its own scale (float32 cdf_plus - cdf_min keeps ~1e-3 here)
    for k in range(3):
        assert rel_err(yg.grad[:, k * nr:(k + 1) * nr], y64.grad[:, k * nr:(k + 1) * nr]) < 2e-3


@pytest.mark.parametrize("B,Cl,Cg,H,f", [(2, 64, 128, 16, 64), (1, 5, 3, 375, 64), (3, 4, 2, 7, 3),
                                         (2, 6, 0, 9, 64)])
def test_upsample_concat_matches_resize_images(B, Cl, Cg, H, f):
    """net.py:58-63 (F.resize_images x64 + speaker broadcast + concat) in one kernel against the
    oracle's float64-coordinate restatement; backward against autograd of it."""
    from chainer_vq_vae_b200 import functions as Fn
    rng = np.random.default_rng(H)
    local = torch.from_numpy(rng.normal(size=(B, Cl, H, 1)).astype(np.float32))
    glob = torch.from_numpy(rng.normal(size=(B, Cg)).astype(np.float32))
    T = H * f
    lo = local.double().requires_grad_(True)
    go = glob.double().requires_grad_(True)
    parts = [O.resize_images_h(lo, T)]
    if Cg:
        parts.append(O.resize_images_h(go.reshape(B, Cg, 1, 1), T))
    want = torch.cat(parts, dim=1)
    lg, gg = local.cuda().requires_grad_(True), glob.cuda().requires_grad_(True)
    got = Fn.upsample_concat(lg, gg, T)
    assert got.shape == want.shape and rel_err(got, want) < 1e-6
    g = torch.from_numpy(rng.normal(size=tuple(want.shape)).astype(np.float32))
    (want * g.double()).sum().backward()
    (got * g.cuda()).sum().backward()
    assert rel_err(lg.grad, lo.grad) < 1e-5
    if Cg:
        assert rel_err(gg.grad, go.grad) < 1e-5


def test_vq_full_size_bit_exact_and_properties():
    """BASELINE configs[1]/[2] VQ size (B=16, d=64, T_z=120, k=512) and a 64x longer one: indices
    bit-identical to the NumPy formulation of utils.py:189-203, idempotence (quantising the
    codebook vectors returns the same indices and vectors), first-minimum ties."""
    rng = np.random.default_rng(2)
    for B, d, T, k in ((16, 64, 120, 512), (4, 64, 7680, 512)):
        W = rng.normal(0, 1 / np.sqrt(d), size=(k, d)).astype(np.float32)
        W[37] = W[5]                                             # duplicate rows: tie -> lowest k
        z = rng.normal(size=(B, d, T, 1)).astype(np.float32)
        z[0, :, 3, 0] = W[37]
        e, idx, cnt, zsum, sq = Fn_vq(torch.from_numpy(z).cuda(), torch.from_numpy(W).cuda())
        want = O.vq_indexes(z, W)                 # sequential-over-d restatement (any size)
        if T <= 120:
            assert np.array_equal(want, O.vq_indexes_numpy(z, W))   # == the literal (B,k,d,T,1) form
        got = idx.cpu().numpy()
        assert np.array_equal(got, want)
        assert got[0, 3, 0] == 5
        assert int(cnt.sum()) == B * T and np.array_equal(
            cnt.cpu().numpy().astype(np.int64), np.bincount(want.ravel(), minlength=k))
        # gather: e = W[idx] exactly, and quantising e again is a fixed point
        assert np.array_equal(e.cpu().numpy()[:, :, :, 0], W[want[:, :, 0]].transpose(0, 2, 1))
        e2, idx2, _, _, _ = Fn_vq(e, torch.from_numpy(W).cuda())
        assert torch.equal(idx2, idx) and torch.equal(e2, e)


def Fn_vq(z, W):
    from chainer_vq_vae_b200 import functions as Fn
    return Fn.vq_lookup(z, W, stats=True)
