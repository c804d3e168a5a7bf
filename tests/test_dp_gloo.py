"""Data-parallel host logic with world_size=2 on the gloo backend (CPU): the flat gradient
bucket is all-reduced with SUM and every rank applies the identical Adam step with
alpha = lr/n -- equivalent to the reference's addgrads / copyparams scheme
(updaters.py:36-38,71-77; train.py:101)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import chainer_vq_vae_b200 as V
    from oracle import vqvae_oracle as O
    torch.set_num_threads(1)
    cfg = O.config_cpu()
    cfg.length, cfg.batch = 128, 4
    params = O.make_params(cfg)
    inp = O.make_inputs(cfg)

    class Holder(torch.nn.Module):           # the parameter tree, without the CUDA forward
        def __init__(self):
            super().__init__()
            self.p = torch.nn.ParameterDict({k.replace("/", "__"): torch.nn.Parameter(v.clone())
                                             for k, v in params.items()})
            self.vq = torch.nn.Module()
    model = Holder()
    opt = V.Adam(2e-4 / world).setup(model)                               # train.py:101
    # this rank's replica gradient: oracle on batch[rank::world] (updaters.py:36-38)
    sl = slice(rank, None, world)
    args = [torch.from_numpy(inp[k][sl]) for k in ("x_enc", "x_dec", "speaker", "t")]
    _, g, _ = O.three_loss_grads(params, cfg, *args)
    opt.bucket.zero()
    with torch.no_grad():
        for k, p in model.p.items():
            p.grad.copy_(g[k.replace("__", "/")])
    opt.bucket.allreduce()                                                # addgrads, :71-72
    summed = {k.replace("__", "/"): p.grad.clone() for k, p in model.p.items()}
    opt.update()
    q.put((rank, {k: v.numpy() for k, v in summed.items()},
           {k.replace("__", "/"): p.detach().numpy().copy() for k, p in model.p.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_allreduce_equals_reference_parallel_updater():
    import numpy as np
    from oracle import vqvae_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    cfg = O.config_cpu()
    cfg.length, cfg.batch = 128, 4
    params = O.make_params(cfg)
    _, total = O.parallel_update_grads(params, cfg, O.make_inputs(cfg), 2)
    for k, want in total.items():
        for r in res:
            assert np.allclose(r[1][k], want.numpy(), rtol=1e-5, atol=1e-7), k
    # every rank ends with identical parameters (no broadcast needed, updaters.py:76-77)
    for k in total:
        assert np.array_equal(res[0][2][k], res[1][2][k]), k
    # and they equal one Adam step on the summed gradient with alpha = lr/2
    for k in ("decoder/proj2/W", "vq/W"):
        p = params[k].clone()
        O.adam_step(p, total[k], torch.zeros_like(p), torch.zeros_like(p), 1, 1e-4)
        assert np.allclose(res[0][2][k], p.numpy(), atol=1e-7), k


def _sync_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import chainer_vq_vae_b200 as V
    torch.set_num_threads(1)
    torch.manual_seed(100 + rank)                       # replicas built from DIFFERENT seeds

    class Dec(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.randn(7, 5))

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Parameter(torch.randn(33))
            self.decoder = V.ExponentialMovingAverage(Dec(), 0.999)
            self.vq = torch.nn.Module()
    model = Holder()
    with torch.no_grad():
        model.decoder.ema.w.add_(rank + 1.0)
    opt = V.Adam(1e-3).setup(model)
    opt.flat_m.normal_()
    opt.flat_v.uniform_()
    opt.t = 10 + rank
    upd = V.VQVAE_ParallelUpdater(None, opt)
    upd.sync_replicas()
    q.put((rank, opt.flat_p.clone().numpy(), opt.flat_m.clone().numpy(), opt.flat_v.clone().numpy(),
           model.decoder.ema.w.detach().numpy().copy(), opt.t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_parallel_updater_broadcasts_rank0_state_once():
    """copyparams (updaters.py:76-77) replaced by one broadcast before the first step: replicas
    built from different seeds / resume states end up with rank 0's parameters, Adam moments,
    step count and EMA copy."""
    import numpy as np
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i in range(1, 5):
        assert np.array_equal(res[0][i], res[1][i]), i
    assert res[0][5] == res[1][5] == 10
