"""In-step kernel times of the training step (CUPTI activity records through torch.profiler).

ncu's launch list serialises the kernels, flushes the caches before each one and runs them at
idle clocks -- fine for shares of the big kernels, misleading for the ~10 us ones.  This tool
runs the same step as bench.py, device resident, under torch.profiler and prints, per kernel
name, the time inside the running step, plus the idle time of the stream between kernels.

    python tools/step_trace.py [--mol] [--steps 3] > profiles/rN_step_trace.txt
"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mol", action="store_true")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--mode", default="bf16x3")
    args = ap.parse_args()
    import chainer_vq_vae_b200 as V
    dev = torch.device("cuda", 0)
    cfg = dict(bench.CFG)
    if args.mol:
        cfg.update(use_logistic=True, input_dim=1, n_mixture=30)
    model = bench.build_model(cfg, dev, args.mode)
    model.train()
    opt = V.Adam(cfg["lr"]).setup(model)
    ex = bench.synthetic_examples(cfg["batch"], cfg["length"], 71, mol=args.mol)

    class It:
        def next(self):
            return ex
    upd = V.VQVAE_ParallelUpdater(It(), opt, device=dev)
    batch = V.updaters.concat_examples(ex, dev)
    for _ in range(4):
        upd.update_from_arrays(batch)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            upd.update_from_arrays(batch)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda x: x[0])
    if not ks:
        print("no CUDA activity records")
        return
    span = (max(k[1] for k in ks) - ks[0][0]) / 1e3 / args.steps
    tot = collections.defaultdict(lambda: [0.0, 0])
    busy, gap, last_end = 0.0, 0.0, ks[0][0]
    big_gaps = []
    for s, e, n in ks:
        tot[n][0] += (e - s)
        tot[n][1] += 1
        if s > last_end:
            gap += s - last_end
            if s - last_end > 20:
                big_gaps.append((s - last_end, n))
        busy += max(0, e - max(s, last_end))
        last_end = max(last_end, e)
    print(f"{args.steps} steps: span {span:.3f} ms/step, busy {busy / 1e3 / args.steps:.3f} ms/step, "
          f"idle between kernels {gap / 1e3 / args.steps:.3f} ms/step ({len(ks) // args.steps} records/step)")
    print("| ms / step | launches / step | us / launch | kernel |\n|---:|---:|---:|---|")
    for n, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"| {t / 1e3 / args.steps:.3f} | {c / args.steps:.1f} | {t / c:.1f} | `{n[:90]}` |")
    print("\nlargest idle gaps (us, kernel that followed):")
    for g, n in sorted(big_gaps, reverse=True)[:15]:
        print(f"  {g:8.1f}  {n[:80]}")


if __name__ == "__main__":
    main()
