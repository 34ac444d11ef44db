#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (CSV from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv`): launches, summed duration, summed DRAM bytes, share of the step.  The list covers
`steps` timed steps + warm-up of bench.py; pass --per N to divide by the number of forward passes in the capture."""
import csv, sys, collections, re, argparse

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--per", type=float, default=1.0, help="divide totals by this many passes")
ap.add_argument("--top", type=int, default=25)
args = ap.parse_args()
rows = []
with open(args.csv, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {n: i for i, n in enumerate(hdr)}
agg = collections.defaultdict(lambda: {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
ids = set()
for r in rd:
    if len(r) != len(hdr):
        continue
    name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]])
    name = name.replace("void ", "").replace("chg::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    m, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    a = agg[name]
    if m == "gpu__time_duration.sum":
        a["n"] += 1
        a["ns"] += val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        a["rd" if "read" in m else "wr"] += val * scale
tot_ns = sum(a["ns"] for a in agg.values())
tot_b = sum(a["rd"] + a["wr"] for a in agg.values())
print(f"launches {sum(a['n'] for a in agg.values())}, kernel time {tot_ns / 1e6 / args.per:.3f} ms, DRAM bytes {tot_b / 1e9 / args.per:.3f} GB (per pass; capture / {args.per:g})")
print("| kernel | launches | ms | share | DRAM read GB | DRAM write GB | GB/s |")
print("|---|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"])[: args.top]:
    gbs = (a["rd"] + a["wr"]) / a["ns"] if a["ns"] else 0.0
    print(f"| `{name[:70]}` | {a['n'] / args.per:g} | {a['ns'] / 1e6 / args.per:.3f} | {a['ns'] / tot_ns:.3f} | {a['rd'] / 1e9 / args.per:.3f} | {a['wr'] / 1e9 / args.per:.3f} | {gbs:.0f} |")
