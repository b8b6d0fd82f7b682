"""Static view of a kernel's loops from `cuobjdump -sass`: for every backward branch, the instruction count and opcode mix of
the loop body [target, branch].  Used to count the instructions a worker warp issues per 64-column chunk without GPU time.
    cuobjdump -sass -fun <mangled> gw_tc3.o | python tools/sass_loops.py [min_len]"""
import collections
import re
import sys

ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
min_len = int(sys.argv[1]) if len(sys.argv) > 1 else 20
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:[!U]*P\d+,?\s*)?(?:`\(\.\w+\)|0x([0-9a-f]+))", t)
    if m and m.group(1):
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            loops.append((addr_index[tgt], i))
def op(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0].split(".")[0]
for s, e in sorted(loops):
    n = e - s + 1
    if n < min_len:
        continue
    c = collections.Counter(op(t) for _, t in ins[s : e + 1])
    tags = [k for k in ("LDTM", "UTCHMMA", "STS", "LDG", "STG", "SYNCS", "NANOSLEEP", "SHFL", "UBLKCP") if c.get(k)]
    print(f"loop {ins[s][0]:#07x}..{ins[e][0]:#07x}  {n:5d} instr  " + " ".join(f"{k}={c[k]}" for k in tags))
    print("    " + " ".join(f"{k}:{v}" for k, v in c.most_common(14)))
