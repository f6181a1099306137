"""Aggregate the warp-stall samples of one kernel of an .ncu-rep (source page, SASS view) by stall reason, and list the
SASS instructions with the most samples.   python tools/ncu_stalls.py rep.ncu-rep <kernel regex> [launch index]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout.splitlines()
print(out[0][:160])
rows = list(csv.reader(out[1:]))
hdr = rows[0]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = {hdr[i]: 0 for i in stall_cols}
items = []
ninst = 0
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    if not (r[iN] or "0").replace(",", "").isdigit():
        continue                      # repeated header rows (one per source view)
    n = int(r[iN] or 0)
    ninst += int(r[iI] or 0)
    for i in stall_cols:
        tot[hdr[i]] += int(r[i] or 0)
    top = max(stall_cols, key=lambda i: int(r[i] or 0))
    items.append((n, r[iS].strip(), hdr[top], int(r[iI] or 0)))
S = sum(tot.values())
print(f"warp instructions executed: {ninst}; stall samples: {S}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {k:28s} {v:8d}  {100 * v / max(S, 1):5.1f}%")
print("top instructions by samples:")
for n, src, top, ie in sorted(items, key=lambda t: -t[0])[:28]:
    print(f"  {n:6d}  {100 * n / max(S, 1):4.1f}%  {top:22s} x{ie:<8d} {src[:90]}")
