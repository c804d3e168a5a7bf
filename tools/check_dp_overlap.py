"""2+ GPUs under torchrun: the overlapped gradient all-reduce (updaters._OverlappedReduce) gives the
same reduced gradients as one all-reduce over the whole bucket, and reports the step time of both.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_dp_overlap.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
import chainer_vq_vae_b200 as V

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(bench.CFG)
cfg["batch"] = 4
out = {}
for overlap in (False, True):
    model = bench.build_model(cfg, dev, os.environ.get("VQW_BENCH_MODE", "fp16x3"))
    model.train()
    opt = V.Adam(cfg["lr"] / world).setup(model)
    batch = bench.synthetic_examples(cfg["batch"], cfg["length"], 71 + rank)

    class It:
        def next(self):
            o = [None] * (cfg["batch"] * world)
            o[rank::world] = batch
            return o
    upd = V.VQVAE_ParallelUpdater(It(), opt, device=dev)
    upd.overlap_allreduce = overlap
    # one step by hand from the common initial state: the reduced gradient bucket before Adam
    arrays = upd.converter(batch, dev)
    l1, l2, l3 = model(*arrays)
    upd.backward_three(model, l1, l2, l3)
    upd._reduce(opt)
    torch.cuda.synchronize()
    flat = opt.bucket.flat.clone()
    losses = []
    for _ in range(3):
        losses.append([float(v) for v in upd.update()])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    for _ in range(5):
        upd.update()
    e1.record()
    torch.cuda.synchronize()
    out[overlap] = (flat, losses, e0.elapsed_time(e1) / 5)
    used = upd._ovl is not None
    del upd, opt, model
    torch.cuda.empty_cache()
    if rank == 0:
        print(f"overlap={overlap} (observer used: {used}): {out[overlap][2]:.2f} ms/step, losses {losses[-1]}",
              flush=True)
diff = float((out[True][0] - out[False][0]).abs().max())
ref = float(out[False][0].abs().max())
# replicas must agree with each other, and the two reduction schedules with one another
p = out[True][0].clone()
dist.broadcast(p, src=0)
rep = float((p - out[True][0]).abs().max())
if rank == 0:
    print(f"max |grad(overlap) - grad(single all-reduce)| = {diff:.3e} (max |grad| {ref:.3e}: the "
          f"weight-gradient atomics alone differ by ~1e-6 run to run); max replica difference {rep:.3e}",
          flush=True)
assert diff <= 1e-5 * ref and rep == 0.0
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("OK")
