"""Summarise ncu outputs into small text files for profiles/.
    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.md
    python tools/summarize_ncu.py report   gpurun_out/prof_x.ncu-rep   > profiles/r1_x.md
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.sum"]


def short(name):
    name = name.replace("void ", "").replace("pcm::", "")
    return name.split("(")[0][:70]


def launches(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
            rows.append((short(r["Kernel Name"]), us, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for n, us, g, b in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list summary ({path}) — {len(rows)} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {us:.1f} | {100 * us / tot:.1f}% |")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary ({path})\n")
    for vals in rows[2:]:
        kn = vals[hdr.index("Kernel Name")]
        print(f"## `{short(kn)}`  grid {vals[hdr.index('Grid Size')]} block {vals[hdr.index('Block Size')]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                print(f"| {h} | {v} | {u} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
