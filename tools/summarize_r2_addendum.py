#!/usr/bin/env python
"""profiles/SUMMARY_r2_addendum.md: the evidence gathered after SUMMARY_r2.md was generated (launch lists with DRAM
bytes, ncu captures of the wgrad / bwd2 kernels, the staging-memory probes, the wire-format A/B)."""
import glob, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def sh(*args):
    return subprocess.run([sys.executable, *args], capture_output=True, text=True, cwd=ROOT).stdout


lines = ["# Profiles r2 - addendum", "",
         "Everything here was measured on one B200 of the pool after `SUMMARY_r2.md` was written (tools/gpu_r16..r2x.sh). "
         "ncu per-launch times are cold-cache and serialised: compare SHARES and bytes; throughput numbers come from bench.py.", ""]
for wl in ("c3", "c4"):
    p = os.path.join(PROF, f"launches_{wl}_r2.csv")
    if os.path.exists(p):
        lines += [f"## Launch list with DRAM bytes, workload {wl} (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                  "dram__bytes_write.sum`, first 1200 launches = 7 forward+reverse passes incl. warm-up, parity and e2e legs)", "",
                  sh("tools/launch_list_summary.py", p, "--per", "7", "--top", "22")]
for rep, title in (("prof_wgrad_tc_r2.ncu-rep", "wgrad_tc_kernel, 64-row-stage version, c5 step (the 32-row / two-CTA version that followed is 1.2x faster: 6.4 vs 7.8 ms per step)"),
                   ("prof_bwd2_r2.ncu-rep", "gated_bwd2_kernel (second-order reverse of the message kernels), c5 step")):
    rp = os.path.join(OUT, rep)
    if os.path.exists(rp):
        lines.append(sh("tools/ncu_section.py", rp, title))
for name, title in (("r17_h2d_probe.log", "H2D / D2H rates of torch pinned buffers; build_batch before the wire format"),
                    ("r18_probe_w1.log", "build_batch with the wire format (ordinary pinned staging)"),
                    ("r19_probe2.log", "H2D of FRESHLY WRITTEN staging memory: ordinary pinned vs write-combined, 1 vs 8 writer threads"),
                    ("r20_probe.log", "build_batch with the wire format and write-combined staging")):
    p = os.path.join(OUT, name)
    if os.path.exists(p):
        lines += [f"## {title} (`{name}`)", "", "```"] + [l.rstrip() for l in open(p) if "Warning" not in l] + ["```", ""]
rows = []
for p in sorted(glob.glob(os.path.join(OUT, "r18_bench_*_w*.json")) + glob.glob(os.path.join(OUT, "r2[0-9]_bench_c[23].json"))):
    try:
        d = json.loads([l for l in open(p) if l.startswith("{")][0])
    except Exception:  # noqa: BLE001
        continue
    c4 = d.get("c4") or {}
    rows.append(f"| `{os.path.basename(p)}` | {d['ms_per_step']:.2f} | {d['e2e']['ms_per_step']:.2f} | {d['e2e']['h2d_bytes_per_step']} | "
                f"{(d['e2e'].get('breakdown') or {}).get('pack_h2d_csr_ms', 0):.2f} | {c4.get('ms_per_step', '')} | {(c4.get('e2e') or {}).get('ms_per_step', '')} |")
if rows:
    lines += ["## Wire format A/B (`w1` compact wire format, `w0` full format; r2x = wire format + write-combined staging)", "",
              "| run | kernel path ms | e2e ms | H2D bytes | pack+H2D+CSR ms | c4 kernel ms | c4 e2e ms |", "|---|---|---|---|---|---|---|"] + rows + [""]
open(os.path.join(PROF, "SUMMARY_r2_addendum.md"), "w").write("\n".join(lines) + "\n")
print("wrote profiles/SUMMARY_r2_addendum.md", len(lines))
