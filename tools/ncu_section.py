#!/usr/bin/env python
"""Markdown section (metric table + warp-state samples) for one .ncu-rep: `ncu_section.py <rep> "<title>"`."""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from summarize_r2 import KEYS, fmt, raw_table, stall_summary

rep, title = sys.argv[1], sys.argv[2]
tab = raw_table(rep)
print(f"## ncu --set full: {title}\n")
print("| kernel | " + " | ".join(KEYS.values()) + " |")
print("|---|" + "---|" * len(KEYS))
for d in tab:
    print(f"| `{d['kernel']}` | " + " | ".join(fmt(d[k]) if k in d else "" for k in KEYS) + " |")
print()
for i in range(len(tab)):
    ss = stall_summary(rep, i)
    if ss:
        print(f"* warp-state samples of `{re.sub(r'[(].*', '', ss[0])[-60:]}`: " + ", ".join(f"{n} {v:.0%}" for n, v in ss[1]))
print()
