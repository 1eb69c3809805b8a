#!/usr/bin/env python
"""Summarise ncu artefacts brought back from the GPU box into small, committed text files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/rNN_launches_X.md
    python tools/ncu_summary.py full     gpurun_out/prof_X.ncu-rep profiles/rNN_ncu_full_X.md
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
]


def launches(src, dst):
    rows = list(csv.reader(open(src, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[mu], v)
        name = re.sub(r"\(.*", "", r[kn])[:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src}); per-launch times are cold-cache and serialised -- compare SHARES\n\n")
        f.write("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.3f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |\n")
        f.write(f"\ntotal {tot / 1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n")
        for r in rows[2:]:
            f.write(f"\n## {r[h.index('Kernel Name')][:110]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for m in METRICS:
                if m in h:
                    f.write(f"| {m} | {r[h.index(m)]} | {units[h.index(m)]} |\n")
            try:
                tr = float(r[h.index("dram__bytes_read.sum")]) + float(r[h.index("dram__bytes_write.sum")])
                f.write(f"| traffic = dram read + write | {tr:.4f} | {units[h.index('dram__bytes_read.sum')]} |\n")
            except Exception:
                pass


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
