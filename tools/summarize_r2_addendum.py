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
# ---- validation of the final tree (tools/gpu_r24.sh / gpu_r25.sh): bench lines, test / sanitizer / smoke verdicts ----
lines += ["## Final tree: bench lines (`profiles/r2_bench_*.json`), tests, sanitizer", "",
          "| record | workload | kernel path ms/step | value | e2e ms/step | e2e value | roofline frac | parity ok | launches |", "|---|---|---|---|---|---|---|---|---|"]
for name in ("r2_bench_c3_default.json", "r2_bench_c2.json", "r2_bench_c1.json", "r2_bench_c5.json", "r2_bench_c3_2gpu.json"):
    p = os.path.join(PROF, name)
    if not os.path.exists(p):
        continue
    try:
        d = json.load(open(p))
    except ValueError:
        d = json.loads([l for l in open(p) if l.startswith("{")][0])
    lines.append(f"| `{name}` | {d['config']['workload'][:40]} ({d['n_gpus']} GPU) | {d['ms_per_step']:.3f} | {d['value']:.1f} {d['unit']} | "
                 f"{d['e2e']['ms_per_step']:.3f} | {d['e2e']['value']:.1f} | {(d.get('roofline') or {}).get('frac')} | "
                 f"{(d.get('parity') or {}).get('ok')} | {d.get('gpu_launches')} |")
    c4 = d.get("c4") or {}
    if c4:
        md = c4.get("md") or {}
        lines.append(f"| &nbsp;&nbsp;`c4` key of the same line | 10 000-atom cell | {c4.get('ms_per_step')} | {c4.get('atoms_per_s')} atoms/s | "
                     f"{(c4.get('e2e') or {}).get('ms_per_step')} (graph) / {(c4.get('e2e_from_structure') or {}).get('ms_per_step')} (structure) | | "
                     f"{(c4.get('roofline') or {}).get('frac')} | | MD ms/step: device {(md.get('device_driver') or {}).get('ms_per_step')}, "
                     f"device+skin {(md.get('device_driver_skin') or {}).get('ms_per_step')}, host loop {(md.get('host_driver') or {}).get('ms_per_step')} |")
lines.append("")
for name, what in (("r2_pytest_gpu.log", "`pytest tests -m gpu`"), ("r2_smoke.log", "`__graft_entry__.smoke()`"), ("sanitizer_memcheck_r2.log", "`compute-sanitizer --tool memcheck` (tools/gpu_sanitize_r2.sh)"),
                   ("r2_predict_structure_list.log", "structures -> graphs -> E/F/sigma for a list of 256 structures (tools/time_convert_many.py)")):
    p = os.path.join(PROF, name)
    if os.path.exists(p):
        tail = [l.rstrip() for l in open(p) if l.strip() and "Warning" not in l and "Consider using" not in l][-6:]
        lines += [f"* {what} -> `profiles/{name}`:", "", "```"] + [t[:220] for t in tail] + ["```", ""]
p = os.path.join(OUT, "md_small_r2.json")
if os.path.exists(p):
    import shutil
    shutil.copy(p, os.path.join(PROF, "md_small_r2.json"))
    lines += ["## MD steps/s of small cells (tools/time_md_small.py; NVE, 2 fs, 300 K, LiMnO2 supercells)", "",
              "| atoms | device driver, 0.5 A skin, CUDA-graph replay | device driver, lists rebuilt every step | host calculator loop |", "|---|---|---|---|"]
    for r in json.load(open(p)):
        f = lambda k: f"{r[k]['ms_per_step']} ms ({r[k]['steps_per_s']} steps/s)"
        lines.append(f"| {r['atoms']} | {f('device_skin0.5_cuda_graph')} | {f('device_skin0_rebuild_every_step')} | {f('host_calculator_loop')} |")
    lines.append("")
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
lines += ["## Kernel A/B experiments of the last hours (bench.py c3 unless noted; CUDA events, one B200)", "",
          "| experiment | result | kept? |", "|---|---|---|",
          "| `segment_sum<128>`: rows in flight per lane-group 4 / 8, lane-groups per row 1 / 2 / 4 (tools/gpu_r23.sh) | c3 by-centre sum 67.5 us (4 rows, 1 group) -> 63.9 (8, 1) / 64.3 (4, 2) / 63.5 (8, 4); c4 75.7 -> 76.3 / 71.8 / 73.7 | 2 groups per row for segments >= 32 rows: roofline 0.82 -> 0.87 (c3), 0.88 -> 0.93 (c4) |",
          "| `wgrad_tc_kernel`: 64-row stages, 4 loads in flight per thread -> all 48 loads of a stage up front -> 32-row stages, 2 CTAs / SM | 23.9 -> 7.8 -> 6.4 ms per c5 step (FFMA kernel: 8.7) | yes |",
          "| `gated_ws_bwd_kernel`: load loops unrolled 4 instead of 2 (tools/gpu_r28.sh) | atom 2.14 -> 2.12 ms, bond 2.05 -> 2.09 ms per step: no change | no |",
          "| `gated_ws_fwd_kernel`: setmaxnreg register split, 24 instead of 12 gathers in flight per producer thread (tools/gpu_r29.sh) | AtomConv 1.53 -> 1.47 ms, BondConv 1.55 -> 1.51 ms per step; 31 GPU tests pass | no (`CHG_WS_REGSPLIT=0`): the gathers' parallelism is not the limiter |",
          "| write-combined pinned staging for the batch packer (tools/gpu_r21.sh) | packing 1.1 -> 10.6 ms (write-combining is very slow on these hosts) although the copy itself would run at 54 GB/s | no (`CHGNET_B200_STAGING=wc` selects it) |",
          "| `chg_forward` replayed as one CUDA graph (tools/gpu_r27.sh) | c1 1.204 -> 0.960 ms, c2 7.69 -> 7.37, c3 16.18 -> 15.87 | as `NativeForward.replay` / `CHGNet.static_evaluator` / bench key `graph_replay`; the headline `value` stays the eager path |", ""]
open(os.path.join(PROF, "SUMMARY_r2_addendum.md"), "w").write("\n".join(lines) + "\n")
print("wrote profiles/SUMMARY_r2_addendum.md", len(lines))
