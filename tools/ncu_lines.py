"""Warp-stall samples of one kernel of an .ncu-rep aggregated per CUDA SOURCE LINE: the SASS view of the report is joined
with `nvdisasm -g` line info of the same cubin by instruction offset.
    python tools/ncu_lines.py rep.ncu-rep <kernel regex> <cubin> <mangled-name substring> [top N]"""
import csv
import re
import subprocess
import sys

rep, pat, cubin, fn = sys.argv[1:5]
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-count", "1"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:]))
hdr = rows[0]
iA, iN, iI = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
recs = []
for r in rows[1:]:
    if len(r) < len(hdr) or not r[iA].startswith("0x"):
        continue
    recs.append((int(r[iA], 16), int(r[iN] or 0), int(r[iI] or 0)))
base = recs[0][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, infn = {}, None, False
for l in dis:
    if l.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", l):
        infn = fn in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
agg = {}
tot = 0
for a, n, ie in recs:
    k = line_of.get(a - base, ("?", 0))
    v = agg.setdefault(k, [0, 0])
    v[0] += n
    v[1] += ie
    tot += n
print(f"total samples {tot}")
src_cache = {}
def src(f, ln):
    if f not in src_cache:
        try:
            import glob
            path = glob.glob("physics-based-climate-model_b200/csrc/" + f)[0]
            src_cache[f] = open(path).read().splitlines()
        except Exception:
            src_cache[f] = []
    L = src_cache[f]
    return L[ln - 1].strip()[:100] if 0 < ln <= len(L) else ""
print("--- by line (file order) ---")
for (f, ln), (n, ie) in sorted(agg.items()):
    if n >= max(1, tot // 200):
        print(f"{f}:{ln:5d} {n:6d} {100 * n / tot:5.1f}%  x{ie:<9d} {src(f, ln)}")
