#!/usr/bin/env python
"""profiles/scatter_traffic.json <- dram__bytes_read + dram__bytes_write per launch of the scatter kernel from an
`ncu --set full` capture of `bench.py --workload <wl> --scatter-only`:  update_scatter_traffic.py <rep> <c3|c4>"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from summarize_r2 import bytes_of, raw_table

rep, wl = sys.argv[1], sys.argv[2]
tab = [d for d in raw_table(rep) if "segment_sum_kernel<128" in d["kernel"] or "segment_sum_kernel<(int)128" in d["kernel"]]
b = [bytes_of(d["dram__bytes_read.sum"]) + bytes_of(d["dram__bytes_write.sum"]) for d in tab]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "scatter_traffic.json")
old = json.load(open(path))
old[f"{wl}_w128"] = {"dram_bytes_per_launch": sum(b) / len(b), "launches": len(b), "kernel": tab[0]["kernel"],
                     "source": f"ncu --set full, bench.py --workload {wl} --scatter-only (r2, final kernel: two lane-groups per row)"}
json.dump(old, open(path, "w"), indent=1)
print(old[f"{wl}_w128"])
