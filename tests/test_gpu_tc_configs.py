"""tcgen05 residual stack at the configurations round 1 left untested:
filter_size=2 (odd tap delays, K1 = 2*Cr + Cc), Cc=640 (the reference default, params.py:40-41),
Cs=512, ragged T -- forward AND backward, each compared DIRECTLY with the float64 oracle stack
(`torch.autograd` of oracle/vqvae_oracle.py::residual_block_forward, modules.py:30-56).  The
stack has no ReLU, so the comparison is kink-free: every gradient is held to the north star's
1e-3 in the max norm (measured ~1e-5)."""
import numpy as np
import pytest
import torch

import chainer_vq_vae_b200 as V
from chainer_vq_vae_b200 import _lib as L
from oracle import vqvae_oracle as O
from helpers import TOL, rel_err
from test_gpu_tc import ORDER, _stack_case

pytestmark = pytest.mark.gpu
DEV = "cuda"

# (fs, Cc, Cs, T, dilations)
CONFIGS = [(2, 192, 256, 200, [1, 2, 4]),       # fs=2: odd delays, ragged last tile
           (2, 640, 256, 256, [1, 3, 512]),     # the reference default channel counts
           (3, 640, 256, 384, [2, 64, 1]),
           (3, 192, 512, 256, [4, 1, 16])]      # Cs = 512: two skip N-chunks


def _oracle_grads(cfg, p, x, c, dil, g_skip, g_res):
    """float64 forward + autograd of the oracle stack; returns outputs and every gradient."""
    leaf = {k: v.double().requires_grad_(True) for k, v in p.items()}
    xx = x.double().requires_grad_(True)
    cc = c.double().requires_grad_(True)
    h, skip = xx, None
    for i, d in enumerate(dil):
        h, s = O.residual_block_forward(O.sub(leaf, f"resnet/{i}/"), h, cc, cfg.filter_size, d)
        skip = s if skip is None else skip + s
    loss = (skip * g_skip.double()).sum()
    if g_res is not None:
        loss = loss + (h * g_res.double()).sum()
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return skip.detach(), h.detach(), xx.grad, cc.grad, grads


def _tc_grads(cfg, p, x, c, dil, g_skip, g_res, mode="bf16x3"):
    weights = []
    for i in range(len(dil)):
        weights += [p[f"resnet/{i}/{n}"].to(DEV).requires_grad_(True) for n in ORDER]
    xg = x.to(DEV).requires_grad_(True)
    cg = c.to(DEV).requires_grad_(True)
    out = V.residual_stack(xg, cg, dil, cfg.filter_size, weights, L.MODES[mode],
                           keep_last_residual=g_res is not None)
    if g_res is not None:
        skip, res = out
        ((skip * g_skip.to(DEV)).sum() + (res * g_res.to(DEV)).sum()).backward()
    else:
        skip, res = out, None
        (skip * g_skip.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    grads = {f"resnet/{i}/{n}": weights[8 * i + j].grad for i in range(len(dil))
             for j, n in enumerate(ORDER)}
    return skip.detach(), res, xg.grad, cg.grad, grads


def _compare(cfg, p, x, c, dil, keep_last, tol_fwd, tol_bwd, seed=5, mode="bf16x3", tol_w=None):
    rng = np.random.default_rng(seed)
    B, T = x.shape[0], x.shape[2]
    Cs, Cr = p["resnet/0/skip/W"].shape[0], p["resnet/0/res/W"].shape[0]
    g_skip = torch.from_numpy(rng.normal(size=(B, Cs, T, 1)).astype(np.float32))
    g_res = torch.from_numpy(rng.normal(size=(B, Cr, T, 1)).astype(np.float32)) if keep_last else None
    so, ro, gxo, gco, go = _oracle_grads(cfg, p, x, c, dil, g_skip, g_res)
    sg, rg, gxg, gcg, gg = _tc_grads(cfg, p, x, c, dil, g_skip, g_res, mode=mode)
    worst = {"skip": rel_err(sg, so), "gx": rel_err(gxg, gxo), "gcond": rel_err(gcg, gco)}
    if keep_last:
        worst["residual"] = rel_err(rg, ro)
    assert worst["skip"] < tol_fwd, worst
    if keep_last:
        assert worst["residual"] < tol_fwd, worst
    for k, g in go.items():
        if float(g.abs().max()) == 0.0:     # last block's unused res branch (modules.py:52,91)
            assert float(gg[k].abs().max()) == 0.0, k
            continue
        worst[k] = rel_err(gg[k], g)
    # tol_w: separate bound for the weight / bias gradients (fp16x3 contracts them on the hi planes)
    data = ("skip", "residual", "gx", "gcond")
    bad = {k: v for k, v in worst.items()
           if v >= (tol_bwd if (k in data or tol_w is None) else tol_w)}
    assert not bad, bad
    return worst


@pytest.mark.parametrize("fs,Cc,Cs,T,dil", CONFIGS)
@pytest.mark.parametrize("keep_last", [False, True])
def test_tc_forward_backward_configs_vs_fp64_oracle(fs, Cc, Cs, T, dil, keep_last):
    B = 2
    cfg, p, x, c = _stack_case(dil, B, T, fs=fs, Cs=Cs, Cc=Cc, seed=fs * 1000 + Cc + Cs)
    worst = _compare(cfg, p, x, c, dil, keep_last, 1e-4, 2e-4)
    print(f"fs={fs} Cc={Cc} Cs={Cs} T={T}: worst {max(worst, key=worst.get)} = "
          f"{max(worst.values()):.2e}")


def test_tc_backward_full_depth_vs_fp64_oracle_kink_free():
    """20 blocks (dilations 1..512 twice, fs=3, 512/512/256, Cc=192), T = 1152 so that the
    dilation-512 taps are live, g_skip fed directly (no head, hence no ReLU kink): every
    gradient of the tcgen05 backward within 1e-3 of float64 autograd in the max norm."""
    dil = [2 ** i for i in range(10)] * 2
    cfg, p, x, c = _stack_case(dil, 1, 1152, seed=77)
    worst = _compare(cfg, p, x, c, dil, False, 1e-4, TOL)
    name = max(worst, key=worst.get)
    print(f"full depth kink-free: worst {name} = {worst[name]:.2e}")


@pytest.mark.parametrize("fs,Cc,Cs,T,dil", CONFIGS)
def test_fp16x3_forward_backward_configs_vs_fp64_oracle(fs, Cc, Cs, T, dil):
    """VQW_MODE_FP16X3: fp16 hi/lo planes (22 significant bits), 3 MMAs per product in the forward
    and in the data-gradient GEMMs -- held to the same bounds as bf16x3 -- and the weight-gradient
    GEMMs on the hi planes (one contraction over time per weight: 2^-12 per operand, once) -- held
    to the north-star bound of 1e-3."""
    cfg, p, x, c = _stack_case(dil, 2, T, fs=fs, Cs=Cs, Cc=Cc, seed=fs * 1000 + Cc + Cs)
    worst = _compare(cfg, p, x, c, dil, True, 1e-4, 2e-4, mode="fp16x3", tol_w=TOL)
    wk = {k: v for k, v in worst.items() if k.startswith("resnet/")}
    print(f"fp16x3 fs={fs} Cc={Cc} Cs={Cs} T={T}: data worst "
          f"{max(v for k, v in worst.items() if not k.startswith('resnet/')):.2e}, "
          f"weight-gradient worst {max(wk, key=wk.get)} = {max(wk.values()):.2e}")


def test_fp16x3_backward_full_depth_vs_fp64_oracle_kink_free():
    """The 20-block kink-free case of the bf16x3 test in VQW_MODE_FP16X3: every gradient within
    1e-3 of float64 autograd in the max norm, data gradients within 2e-4."""
    dil = [2 ** i for i in range(10)] * 2
    cfg, p, x, c = _stack_case(dil, 1, 1152, seed=77)
    worst = _compare(cfg, p, x, c, dil, False, 1e-4, 2e-4, mode="fp16x3", tol_w=TOL)
    wk = {k: v for k, v in worst.items() if k.startswith("resnet/")}
    print(f"fp16x3 full depth: skip {worst['skip']:.2e} gx {worst['gx']:.2e} gcond {worst['gcond']:.2e}, "
          f"weight-gradient worst {max(wk, key=wk.get)} = {max(wk.values()):.2e}, "
          f"median {sorted(wk.values())[len(wk) // 2]:.2e}")


def test_tc_weight_gradient_atomics_run_to_run_bound():
    """The grouped weight-gradient GEMM accumulates its per-item partial tiles with fp32
    atomicAdd (tc_gemm.cu, EPI_WGRAD) and the bias gradients likewise: the summation ORDER is
    not fixed, so two runs may differ in the last bits.  Bound it: every gradient of two
    identical runs agrees to 1e-6 of its max norm (data gradients are bit-identical)."""
    dil = [1, 2, 4, 8]
    cfg, p, x, c = _stack_case(dil, 4, 512, seed=21)
    rng = np.random.default_rng(1)
    g_skip = torch.from_numpy(rng.normal(size=(4, 256, 512, 1)).astype(np.float32))
    a = _tc_grads(cfg, p, x, c, dil, g_skip, None)
    b = _tc_grads(cfg, p, x, c, dil, g_skip, None)
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]), "forward / gx must be deterministic"
    worst = 0.0
    for k in a[4]:
        if float(a[4][k].abs().max()) == 0.0:
            continue
        worst = max(worst, rel_err(a[4][k], b[4][k]))
    print(f"atomic wgrad run-to-run max diff {worst:.2e} (relative to each gradient's max)")
    assert worst < 1e-6


@pytest.mark.parametrize("mode", ["bf16x3", "fp16x3"])
def test_reference_default_model_step_matches_oracle(mode):
    """The reference's own default model (params.py:24-41: filter_size=2, n_loop=3, n_layer=10,
    d=512, k=128, local_condition_dim=512 => Cc=640, length 7680) on the tensor-core path: one
    full forward + three-loss backward of two items against the oracle.  VQ indices bit-exact
    (d=512), logits and losses within 1e-3, gradients by direction / L2 (ReLU kinks in the head,
    see test_full_depth_step_matches_oracle)."""
    import os
    from helpers import build_model, grads_by_name, to_dev
    cfg = O.Config(batch=2, length=7680, n_loop=3, n_layer=10, filter_size=2,
                   residual_channels=512, dilated_channels=512, skip_channels=256,
                   d=512, k=128, local_condition_dim=512, global_condition_dim=128)
    params = O.make_params(cfg)
    inp = O.make_inputs(cfg)
    args = [torch.from_numpy(inp[k]) for k in ("x_enc", "x_dec", "speaker", "t")]
    torch.set_num_threads(max(1, os.cpu_count() or 2))
    losses, grads, inter = O.three_loss_grads(params, cfg, *args)
    model = build_model(cfg, params, mode=mode)
    opt = V.Adam(2e-4).setup(model)
    upd = V.VQVAE_StandardUpdater(None, opt)
    l1, l2, l3 = model(*to_dev(inp, cfg, indices=True))
    upd.backward_three(model, l1, l2, l3)
    torch.cuda.synchronize()
    assert np.array_equal(model.vq.indexes.cpu().numpy(), inter["indexes"])
    e_y = rel_err(model.y, inter["y"])
    print(f"reference-default model ({mode}): logits max-norm rel err {e_y:.2e}")
    assert e_y < 1e-4
    for got, want in zip((l1, l2, l3), losses):
        assert abs(float(got.detach()) - float(want)) <= TOL * abs(float(want))
    got = grads_by_name(model)
    worst_cos, worst2 = 1.0, 0.0
    for name, g in grads.items():
        if float(g.abs().max()) == 0.0:
            continue
        a, b = got[name].flatten().double(), g.flatten().double()
        worst_cos = min(worst_cos, float(torch.dot(a, b) / (a.norm() * b.norm())))
        worst2 = max(worst2, float((a - b).norm() / b.norm()))
    print(f"reference-default model ({mode}): gradients worst cosine {worst_cos:.6f}, worst L2 {worst2:.2e}")
    assert worst_cos > 0.9999 and worst2 < 1e-2


@pytest.mark.parametrize("mode", ["bf16x3", "fp16x3"])
@pytest.mark.parametrize("fs,Cc,Cg,T,dil", [(3, 192, 128, 384, [1, 2, 512]), (2, 640, 128, 256, [4, 1]),
                                            (3, 160, 96, 200, [2, 8])])
def test_tc_hoisted_global_condition_vs_fp64_oracle(fs, Cc, Cg, T, dil, mode):
    """the projection of the time-constant condition channels (the speaker
    embedding that net.py:60-61 broadcasts over time) is hoisted out of the per-step contraction
    into one gate-bias vector per (item, block).  Same function as modules.py:44 on the
    concatenated condition: forward, gx, the gradient of the time-varying channels, the gradient
    of the global vector (= time sum of the oracle's gradient of those channels) and every
    weight gradient -- including the global columns of condition_proj.W -- against float64."""
    B, Cl = 2, Cc - Cg
    cfg, p, x, c = _stack_case(dil, B, T, fs=fs, Cc=Cc, seed=Cc + Cg)
    rng = np.random.default_rng(3)
    glob = torch.from_numpy(rng.normal(size=(B, Cg)).astype(np.float32))
    c = c.clone()
    c[:, Cl:] = glob[:, :, None, None]                       # constant over time
    g_skip = torch.from_numpy(rng.normal(size=(B, 256, T, 1)).astype(np.float32))
    g_res = torch.from_numpy(rng.normal(size=(B, 512, T, 1)).astype(np.float32))
    so, ro, gxo, gco, go = _oracle_grads(cfg, p, x, c, dil, g_skip, g_res)
    weights = []
    for i in range(len(dil)):
        weights += [p[f"resnet/{i}/{n}"].to(DEV).requires_grad_(True) for n in ORDER]
    xg = x.to(DEV).requires_grad_(True)
    cl = c[:, :Cl].contiguous().to(DEV).requires_grad_(True)
    gg = glob.to(DEV).requires_grad_(True)
    skip, res = V.residual_stack(xg, cl, dil, fs, weights, L.MODES[mode], keep_last_residual=True,
                                 cond_global=gg)
    ((skip * g_skip.to(DEV)).sum() + (res * g_res.to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    errs = {"skip": rel_err(skip, so), "res": rel_err(res, ro), "gx": rel_err(xg.grad, gxo),
            "gcond_local": rel_err(cl.grad, gco[:, :Cl]),
            "g_global": rel_err(gg.grad, gco[:, Cl:].sum(dim=(2, 3)))}
    for i in range(len(dil)):
        for j, n in enumerate(ORDER):
            g = go[f"resnet/{i}/{n}"]
            if float(g.abs().max()) > 0:
                errs[f"{i}/{n}"] = rel_err(weights[8 * i + j].grad, g)
    # fp16x3: the weight / bias gradients are contracted on the hi planes (north-star bound)
    data = ("skip", "res", "gx", "gcond_local")
    bad = {k: v for k, v in errs.items() if v >= (2e-4 if (mode == "bf16x3" or k in data) else TOL)}
    print(f"hoisted global condition ({mode}) fs={fs} Cc={Cc} Cg={Cg}: worst {max(errs, key=errs.get)} "
          f"= {max(errs.values()):.2e}")
    assert not bad, bad
