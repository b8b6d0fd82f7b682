"""Turns ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/ncu_digest.py launches gpurun_out/launches.csv  > profiles/rNN_launch_summary.txt
    python tools/ncu_digest.py report   gpurun_out/x.ncu-rep "<command line>" > profiles/rNN_ncu_x.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]  # fmt: skip


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn])
        tot[name] += float(r[mv].replace(",", ""))
        cnt[name] += 1
    s = sum(tot.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'share':>7s}")
    for k, v in tot.most_common():
        print(f"{k:60s} {cnt[k]:8d} {v / 1000:12.1f} {v / s:7.3f}")


def report(path, cmd):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, v = rows[0], rows[1], rows[-1]
    print(cmd)
    for m in METRICS:
        if m in h:
            i = h.index(m)
            print(f"{m} = {v[i]} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    h = rows[hi]
    iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    ops, samp = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
        op = m.group(2) if m else "?"
        ops[op] += int(r[iE])
        samp[op] += int(r[iN])
    tot, ts = sum(ops.values()), max(1, sum(samp.values()))
    print(f"\nwarp instructions executed {tot}; opcode mix (share of instructions | share of stall samples):")
    for op, c in ops.most_common(24):
        print(f"  {op:10s} {100 * c / tot:5.1f} % | {100 * samp[op] / ts:5.1f} %")


def traffic(workload, pairs):
    """profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch of the captured kernel classes, tagged with
    the hash of the CUDA sources and the workload they were measured on (bench.py quotes the figure only when both match).
        python tools/ncu_digest.py traffic <workload> proc_edge=gpurun_out/a.ncu-rep dec_edge=gpurun_out/b.ncu-rep > profiles/traffic.json"""
    import json
    import os

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    out = {"source_hash": bench.source_hash(), "workload": workload, "unit": "bytes per launch (dram read + write)"}
    for pair in pairs:
        tag, path = pair.split("=", 1)
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        h, units, v = rows[0], rows[1], rows[-1]
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = h.index(m)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
            tot += float(v[i].replace(",", "")) * scale
        out[tag] = tot
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3:])
    else:
        report(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
