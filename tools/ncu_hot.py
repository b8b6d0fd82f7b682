"""Per-loop dynamic instruction counts and stall samples from an `ncu --page source --csv --print-source sass` export.
    python tools/ncu_hot.py f_proc_src.csv [tiles]
Splits the kernel at backward branches (loops), sums `Instructions Executed` and stall samples per loop body (innermost
attribution), and prints the hottest regions with their top stall reasons."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
addr = [int(r[0], 16) for r in data]
base = addr[0]
src = [r[col["Source"]].strip() for r in data]
ex = [float(r[col["Instructions Executed"]] or 0) for r in data]
smp = [float(r[col["# Samples"]] or 0) for r in data]
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
idx = {a - base: i for i, a in enumerate(addr)}
loops = []
for i, s in enumerate(src):
    m = re.search(r"\bBRA\S*\s+(?:!?U?P\d+,?\s*)?0x([0-9a-f]+)", s)
    if m:
        t = int(m.group(1), 16) - base
        if t in idx and idx[t] <= i:
            loops.append((idx[t], i))
loops.sort(key=lambda p: (p[1] - p[0]))
owner = [-1] * len(src)
for li, (s, e) in enumerate(loops):  # innermost first
    for i in range(s, e + 1):
        if owner[i] < 0:
            owner[i] = li
tot_ex, tot_s = sum(ex), sum(smp)
print(f"total warp-instr {tot_ex:.0f} ({tot_ex / tiles:.0f}/tile), samples {tot_s:.0f}")
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter(), collections.Counter()])
for i in range(len(src)):
    a = agg[owner[i]]
    a[0] += ex[i]
    a[1] += smp[i]
    for sc in stall_cols:
        v = float(data[i][col[sc]] or 0)
        if v:
            a[2][sc] += v
    op = re.sub(r"^@!?U?P\d+\s+", "", src[i]).split(" ")[0].split(".")[0]
    a[3][op] += ex[i]
for li, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    rng = f"{loops[li][0]}..{loops[li][1]} ({loops[li][1] - loops[li][0] + 1} static)" if li >= 0 else "outside loops"
    tags = [k for k in ("LDTM", "UTCHMMA", "STS", "LDG", "STG", "SYNCS", "SHFL", "UBLKCP", "NANOSLEEP", "CS2R") if a[3].get(k)]
    print(f"{rng:34s} exec {a[0] / tot_ex * 100:5.1f}% ({a[0] / tiles:8.0f}/tile) samples {a[1] / tot_s * 100:5.1f}%  [{' '.join(tags)}]")
    print("      stalls: " + " ".join(f"{k[6:]}:{v / a[1] * 100:.0f}%" for k, v in a[2].most_common(5)))
    print("      ops: " + " ".join(f"{k}:{v / a[0] * 100:.0f}%" for k, v in a[3].most_common(10)))
