"""Summaries of the ncu artefacts kept under profiles/ (run here, no GPU needed):
   python tools/ncu_summary.py full   <rep.ncu-rep>       key metrics of every kernel of a --set full capture
   python tools/ncu_summary.py launch <launches.csv[.gz]> [steps]   per-kernel totals of a launch list
"""
import collections
import csv
import gzip
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("| kernel | " + " | ".join(w.split(".")[0] for w in WANT if w in hdr) + " |")
    print("|---|" + "---|" * sum(w in hdr for w in WANT))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])
        vals = []
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                vals.append(f"{r[i]} {units[i]}".strip())
        print(f"| `{name}` | " + " | ".join(vals) + " |")


def launch(path, steps=1.0):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in r:
        try:
            v = float(row[vi].replace(",", ""))
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", row[ki])[:64]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time over {steps:g} step(s)\n")
    print("| ms / step | share | launches / step | us / launch | kernel |")
    print("|---:|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v[1] / 1e6 / steps:.3f} | {100 * v[1] / tot:.1f} % | {v[0] / steps:g} | {v[1] / v[0] / 1e3:.1f} | `{k}` |")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2])
    else:
        launch(sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
