#!/usr/bin/env python
"""bench.py — throughput of the CHGNet hot path (E + F + sigma) on B200.

Contract (see task statement): ``python bench.py --gpus N --steps K --warmup W`` prints ONE
JSON line.  A "step" is one pass of the hot path (forward + the force/stress reverse pass)
over one batch of synthetic CrystalGraphs.

Workloads (BASELINE.json `configs`):
  c2  (default) batch = 64 random periodic cells, 40..60 atoms, cutoffs 6 A / 3 A   [configs[1]]
  c3  batch = 256 random cells, 20..40 atoms                                       [configs[2]]
  c4  one 10,000-atom LiMnO2 supercell (10x5x25), sigma = 0.02 A displacements      [configs[3]]
  c1  the 8-atom LiMnO2 cell                                                       [configs[0]]

value      structures/s of the kernel path, batch descriptor already resident in HBM
e2e        the same through ``CHGNet.predict_graph`` from host CrystalGraphs (host packing,
           H2D, CSR build, kernels, D2H numpy) — the user-facing call
roofline   AtomConv scatter-reduce kernel (chg_segment_sum over center-sorted messages),
           timed alone with CUDA events at this workload's size, L2 flushed between launches
cpu_baseline / --impl reference
           the oracle port of the reference's torch CPU path (oracle/chgnet_oracle.py) on
           the host cores, on a bounded sample of the same workload

Multi-GPU: one process per GPU (torchrun), graphs sharded by rank, no device-path
collective (inference); time = max over ranks; weak scaling (each rank owns one batch).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WEIGHTS = os.path.join(ROOT, "tests", "golden", "chgnet_0.3.0_weights.npz")
L2_FLUSH_BYTES = 256 << 20


def make_workload(name: str, rank: int):
    from chgnet_b200 import graphgen

    if name == "c1":
        z, frac, lat = graphgen.limno2_structure()
        return [graphgen.make_crystal_graph(z, frac, lat, graph_id="mp-18767")], "LiMnO2 mp-18767, 8 atoms"
    if name == "c2":
        return graphgen.random_graphs(64, 40, 60, 1000 + 100 * rank), "batch=64 random periodic cells, 40..60 atoms, rho=0.10/A^3, cutoffs 6/3 A"
    if name == "c3":
        return graphgen.random_graphs(256, 20, 40, 2000 + 1000 * rank), "batch=256 random periodic cells, 20..40 atoms, rho=0.10/A^3, cutoffs 6/3 A"
    if name == "c4":
        z, frac, lat = graphgen.limno2_structure((10, 5, 25), 0.02, 4000 + rank)
        return [graphgen.make_crystal_graph(z, frac, lat, graph_id="LiMnO2-10x5x25")], "LiMnO2 10x5x25 supercell, 10,000 atoms, sigma=0.02 A"
    if name == "c5":
        return graphgen.random_graphs(128, 20, 40, 5000 + 1000 * rank), "fine-tune batch=128 random periodic cells, 20..40 atoms, rho=0.10/A^3, cutoffs 6/3 A; targets 'efsm' (MSE, ratios 1/1/0.1/0.1), Adam lr 1e-3"
    raise SystemExit(f"unknown workload {name}")


def train_labels(preds, seed: int):
    """labels = prediction + uniform noise (+-0.1 eV/atom, +-0.01 eV/A, +-0.05 GPa, +-0.03 muB)"""
    gen = torch.Generator().manual_seed(seed)

    def noisy(v, amp):
        v = torch.as_tensor(np.asarray(v), dtype=torch.float32)
        return v + (torch.rand(v.shape, generator=gen) - 0.5) * 2 * amp

    return {"e": noisy([float(p["e"]) for p in preds], 0.1), "f": [noisy(p["f"], 0.01) for p in preds],
            "s": [noisy(p["s"], 0.05) for p in preds], "m": [noisy(p["m"], 0.03) for p in preds]}


def run_reference_train(args) -> None:
    """--impl reference --workload c5: one reference training step (trainer.py:398-411) on the host
    cores: oracle forward (train mode) -> CombinedLoss('em', MSE) -> backward -> torch Adam."""
    from oracle import chgnet_oracle as orc

    graphs, desc = make_workload("c5", 0)
    sample = graphs[: max(1, min(len(graphs), args.cpu_sample))]
    w = orc.load_weights_npz(WEIGHTS)
    base = orc.predict_graph(w, sample, "efsm", batch_size=len(sample))
    lab = train_labels(base, 5)
    P = {k: torch.as_tensor(np.asarray(v)).float().requires_grad_(k != "composition_model.fc.weight") for k, v in w.items()}
    opt = torch.optim.Adam([v for v in P.values() if v.requires_grad], lr=1e-3)
    crit = torch.nn.MSELoss()

    def step():
        opt.zero_grad()
        out = orc.forward(P, sample, "efsm", train=True)
        loss = (crit(lab["e"], out["e"]) + crit(torch.cat(lab["f"]), torch.cat(out["f"]))
                + 0.1 * crit(torch.stack(lab["s"]), torch.stack(out["s"])) + 0.1 * crit(torch.cat(lab["m"]), torch.cat(out["m"])))
        loss.backward()
        opt.step()

    threads = pick_cpu_threads(step)
    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = len(sample) / dt
    sdesc = f"first {len(sample)} graphs of the batch per step; {threads} of {os.cpu_count()} host threads (fastest of a 4..all sweep)"
    print(json.dumps({
        "impl": "reference", "metric": "train_structures_per_sec_EFSM", "value": value, "unit": "structures/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"c5: {desc}", "task": "train efsm"},
        "cpu_baseline": {"value": value, "unit": "structures/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sdesc},
        "e2e": {"value": value, "unit": "structures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_train(args, rank: int, world: int, local_rank: int) -> None:
    """--workload c5: one fine-tuning step per 'step' (forward, CombinedLoss, parameter gradients,
    one gradient all-reduce over NCCL, fused Adam, weight re-pack)."""
    import contextlib
    import io

    import torch.distributed as dist

    from chgnet_b200.batch import build_batch
    from chgnet_b200.model import CHGNet
    from chgnet_b200.trainer import Trainer, loss_and_grads
    from chgnet_b200.weights import unpack_grads

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    with contextlib.redirect_stdout(io.StringIO()):
        model = CHGNet.from_file(WEIGHTS, version="0.3.0").to(dev)
    graphs, desc = make_workload("c5", rank)
    c = counts(graphs)
    base = model.predict_graph(graphs, task="efsm", batch_size=len(graphs))
    lab = train_labels(base, 5 + rank)
    trainer = Trainer(model, targets="efsm", criterion="MSE", learning_rate=1e-3)
    flush = L2Flush(dev)
    K = model._get_engine().K

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch = build_batch(graphs, dev, with_reverse=True)
    tg_dev = trainer._targets(lab, batch.atoms_per_graph, dev)

    def step_resident():
        engine = model._get_engine()  # re-packs the weights the previous Adam step changed
        report_, G = loss_and_grads(engine, batch, trainer.cfg, tg_dev, model.is_intensive, None)
        fg = trainer.flatten_grads(unpack_grads(G, model.state_dict()))
        if world > 1:
            dist.all_reduce(fg)
        trainer.step_count += 1
        K.adam_step(trainer.flat, fg, trainer.exp_avg, trainer.exp_avg_sq, trainer.lr, 0.9, 0.999, 1e-8, 0.0, trainer.step_count)
        model.mark_params_updated()
        return report_

    for _ in range(max(args.warmup, 3)):
        flush()
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = K.launches
    elapsed_ms = 0.0
    for _ in range(args.steps):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rep_last = step_resident()
        e.record()
        e.synchronize()
        elapsed_ms += s.elapsed_time(e)
    barrier()
    launches = K.launches - launches0
    clocks = sampler.stop()
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps

    # end to end: Trainer.train_step from host graphs + host labels, report read back every step
    targets = lab
    for _ in range(2):
        trainer.train_step(graphs, targets)
    h2d = int(build_batch(graphs, dev, with_reverse=True).h2d_bytes) + 4 * sum(int(v.numel()) for v in tg_dev.values())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush()
        trainer.train_step(graphs, targets)
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / args.steps
    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    # breakdown of one resident step (synchronised, outside the timed loops)
    from chgnet_b200.engine import EV_A3_TO_GPA
    from chgnet_b200.trainer import loss_and_seeds

    def tick():
        torch.cuda.synchronize()
        return time.perf_counter()

    t0 = tick()
    engine = model._get_engine()
    t1 = tick()
    o = engine.run(batch, need_grad=True, need_magmom=True, train=True)
    t2 = tick()
    engine.input_grads(o, record=True)
    t3 = tick()
    n_dev = torch.tensor(batch.atoms_per_graph, device=dev, dtype=torch.float64)
    preds = {"e": ((o.energy + o.e_ref) / n_dev).float(), "m": o.magmom, "f": o.force.float(),
             "s": (o.virial.view(-1, 3, 3) * (EV_A3_TO_GPA / batch.volume.double())[:, None, None]).float()}
    rep_, seeds = loss_and_seeds(K, trainer.cfg, preds, tg_dev, False)  # rank 0 only: no collective here
    t4 = tick()
    G = engine.param_grads(o, (seeds["e"] / n_dev.float()).contiguous(), seeds["m"], seeds["f"], seeds["s"])
    t5 = tick()
    fg = trainer.flatten_grads(unpack_grads(G, model.state_dict()))
    t6 = tick()
    breakdown = {"repack_weights_ms": (t1 - t0) * 1e3, "forward_ms": (t2 - t1) * 1e3, "force_pass_ms": (t3 - t2) * 1e3,
                 "loss_ms": (t4 - t3) * 1e3, "second_order_and_wgrads_ms": (t5 - t4) * 1e3,
                 "unpack_flatten_grads_ms": (t6 - t5) * 1e3}
    peaks, peak_kind = measured_peaks()
    sc_ms, sc_bytes = time_scatter_kernel(K, batch)
    achieved = sc_bytes / (sc_ms * 1e-3) / 1e9
    roofline = {"kernel": "segment_sum_kernel<64> (AtomConv scatter-reduce)", "bound": "hbm", "achieved": round(achieved, 1),
                "peak": peaks["hbm_gbs"], "peak_kind": f"{peak_kind} copy bandwidth", "unit": "GB/s",
                "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": None, "us_per_launch": round(sc_ms * 1e3, 2),
                "algorithmic_bytes": sc_bytes, "bytes_formula": "256*E_d + 256*N + 4*(N+1)"}
    ek = EventKernels(K)
    from chgnet_b200.engine import Engine

    eng2 = Engine(model._get_engine().pw, ek)
    out2 = eng2.run(batch, need_grad=True, need_magmom=True, train=True)
    eng2.input_grads(out2, record=True)
    eng2.param_grads(out2, torch.ones(c["graphs"], device=dev), torch.ones(c["atoms"], device=dev),
                     torch.ones(c["atoms"], 3, device=dev), torch.ones(c["graphs"], 3, 3, device=dev))
    shares = ek.table()
    total = c["graphs"] * world
    print(json.dumps({
        "metric": "train_structures_per_sec_EFSM", "value": total / (ms_per_step * 1e-3), "unit": "structures/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"c5: {desc}", "task": "train efsm", "per_gpu": c, "weights": "CHGNet 0.3.0",
                   "l2": "256 MiB buffer written, then 256 MiB read (clean lines), between timed iterations",
                   "parallelism": f"graph-sharded x{world}, one all-reduce of the flat gradient buffer per step",
                   "second_order": "tangent pass + reverse over (primal, tangent) for the force / stress loss terms"},
        "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "structures/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 48, "api": "Trainer.train_step(list[CrystalGraph] on host, labels on host)"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": None,
        "last_report": rep_last, "breakdown": breakdown, "kernel_shares": shares}), flush=True)
    if world > 1:
        dist.barrier()


def counts(graphs):
    n = sum(int(g.atomic_number.shape[0]) for g in graphs)
    ed = sum(int(g.atom_graph.reshape(-1, 2).shape[0]) for g in graphs)
    a = sum(int(g.bond_graph.reshape(-1, 5).shape[0]) for g in graphs)
    return {"graphs": len(graphs), "atoms": n, "directed_edges": ed, "undirected_bonds": ed // 2, "angles": a}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int) -> None:
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.sm_max = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self) -> None:
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(0.02)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def pick_cpu_threads(fn) -> int:
    """The torch CPU path does not scale to every core of a big host (tiny ops, OpenMP
    fork/join): time `fn` once per candidate thread count and keep the fastest, so the CPU
    baseline is the reference at its best, not at `os.cpu_count()`."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        fn()  # warm
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int) -> None:
    """--impl reference: the reference's CPU path (oracle port), bounded sample per step."""
    if rank != 0:
        return
    from oracle import chgnet_oracle as orc

    graphs, desc = make_workload(args.workload, 0)
    sample = graphs[: max(1, min(len(graphs), args.cpu_sample))]
    if args.workload == "c4":
        from chgnet_b200 import graphgen

        z, frac, lat = graphgen.limno2_structure((5, 4, 3), 0.02, 4000)
        sample = [graphgen.make_crystal_graph(z, frac, lat)]
        desc_s = "LiMnO2 5x4x3 supercell (480 atoms) — largest cell timed on the CPU; value scaled by atoms"
    else:
        desc_s = f"first {len(sample)} graphs of the batch per step"
    w = orc.load_weights_npz(WEIGHTS)
    probe = sample[:2]
    threads = pick_cpu_threads(lambda: orc.predict_graph(w, probe, "efs", batch_size=len(probe)))
    desc_s += f"; {threads} of {os.cpu_count()} host threads (fastest of a 4..all sweep)"
    for _ in range(args.warmup):
        orc.predict_graph(w, sample, "efs", batch_size=len(sample))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.predict_graph(w, sample, "efs", batch_size=len(sample))
    dt = (time.perf_counter() - t0) / args.steps
    c = counts(sample)
    value = c["graphs"] / dt
    if args.workload == "c4":
        value = (c["atoms"] / dt) / counts(graphs)["atoms"]  # 10k-atom structures/s at the same atoms/s
    line = {
        "impl": "reference", "metric": "structures_per_sec_EFS", "value": value, "unit": "structures/s",
        "atoms_per_s": c["atoms"] / dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"{args.workload}: {desc}", "task": "efs"},
        "cpu_baseline": {"value": value, "unit": "structures/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": desc_s},
        "e2e": {"value": value, "unit": "structures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
class L2Flush:
    """Evicts everything of ours from the 126 MB L2 between timed iterations: writes a 256 MiB
    buffer (the rule of the task statement), then streams a second 256 MiB buffer through with a
    read so that the cache is left holding CLEAN lines — otherwise the timed kernel also pays
    for the write-back of the flush buffer's dirty lines."""

    def __init__(self, dev) -> None:
        self.w = torch.empty(L2_FLUSH_BYTES // 4, device=dev)
        self.r = torch.zeros(L2_FLUSH_BYTES // 4, device=dev)
        self.sink = torch.zeros((), device=dev)

    def __call__(self) -> None:
        self.w.zero_()
        self.sink.copy_(self.r.sum())


def time_scatter_kernel(K, batch, n_iter: int = 20):
    """AtomConv scatter-reduce alone: CUDA events on the launching stream, L2 flushed."""
    dev = batch.z.device
    msg = torch.randn(batch.n_edges, 64, device=dev)
    out = torch.empty(batch.n_atoms, 64, device=dev)
    flush = L2Flush(dev)
    for _ in range(3):
        K.segment_sum(msg, None, batch.ptr_c, 0, out)
    total = 0.0
    for _ in range(n_iter):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        K.segment_sum(msg, None, batch.ptr_c, 0, out)
        e.record()
        e.synchronize()
        total += s.elapsed_time(e)
    ms = total / n_iter
    alg_bytes = 256 * batch.n_edges + 256 * batch.n_atoms + 4 * (batch.n_atoms + 1)
    return ms, alg_bytes


class EventKernels:
    """Wraps the kernel binding with per-call CUDA events (used OUTSIDE the timed region)."""

    def __init__(self, inner) -> None:
        self._inner, self.records = inner, []

    def __getattr__(self, name):
        attr = getattr(self._inner, name)
        if not callable(attr) or name.startswith("_"):
            return attr

        def timed(*a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            attr(*a)
            e.record()
            self.records.append((name, s, e))

        return timed

    def table(self):
        torch.cuda.synchronize()
        agg: dict[str, list] = {}
        for name, s, e in self.records:
            agg.setdefault(name, [0.0, 0])
            agg[name][0] += s.elapsed_time(e)
            agg[name][1] += 1
        tot = sum(v[0] for v in agg.values()) or 1.0
        return {k: {"ms": round(v[0], 4), "calls": v[1], "share": round(v[0] / tot, 4)}
                for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}


def run_ours(args, rank: int, world: int, local_rank: int) -> None:
    import torch.distributed as dist

    from chgnet_b200.batch import build_batch
    from chgnet_b200.engine import EV_A3_TO_GPA, Engine
    from chgnet_b200.model import CHGNet

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        model = CHGNet.from_file(WEIGHTS, version="0.3.0").to(dev).eval()
    graphs, desc = make_workload(args.workload, rank)
    c = counts(graphs)
    engine = model._get_engine()
    K = engine.K
    flush = L2Flush(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernel path, inputs resident in HBM ----------------
    batch = build_batch(graphs, dev, with_reverse=True)
    if args.scatter_only:  # ncu capture target: only the AtomConv scatter-reduce launches
        ms, nbytes = time_scatter_kernel(K, batch, n_iter=5)
        print(json.dumps({"scatter_only": True, "us_per_launch": ms * 1e3, "algorithmic_bytes": nbytes}))
        return

    native = model._get_native()  # the product's inference path: ONE chg_forward call per step

    def step_resident():
        out = native(batch, need_grad=True)
        scale = EV_A3_TO_GPA / batch.volume.to(torch.float64)
        stress = (out["virial"].view(-1, 3, 3) * scale[:, None, None]).to(torch.float32)
        return out["energy"], out["force"].to(torch.float32), stress

    for _ in range(max(args.warmup, 3)):
        flush()
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = K.launches
    elapsed_ms = 0.0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step_resident()
        e.record()
        e.synchronize()
        elapsed_ms += s.elapsed_time(e)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = K.launches - launches0
    clocks = sampler.stop()
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps

    # ---------------- end to end through the public API ----------------
    def step_e2e():
        return model.predict_graph(graphs, task="efs", batch_size=len(graphs))

    for _ in range(2):
        preds = step_e2e()
    d2h = sum(int(v.nbytes) for p in preds for v in p.values())
    h2d = int(model.last_batch.h2d_bytes)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush()
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(t.item()) / args.steps

    if rank != 0:
        if world > 1:
            dist.barrier()  # wait for rank 0's extra legs (roofline, shares, CPU baseline)
        return
    # e2e breakdown (one synchronised pass, outside the timed loops)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bb = build_batch(graphs, dev, with_reverse=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    oo = native(bb, need_grad=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    _ = (oo["energy"].cpu(), oo["force"].to(torch.float32).cpu(), oo["virial"].cpu())
    t3 = time.perf_counter()
    breakdown = {"pack_h2d_csr_ms": (t1 - t0) * 1e3, "kernels_ms": (t2 - t1) * 1e3, "d2h_ms": (t3 - t2) * 1e3}
    # ---------------- roofline of the AtomConv scatter kernel ----------------
    peaks, peak_kind = measured_peaks()
    sc_ms, sc_bytes = time_scatter_kernel(K, batch)
    achieved = sc_bytes / (sc_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "scatter_traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
    roofline = {"kernel": "segment_sum_kernel<64> (AtomConv scatter-reduce)", "bound": "hbm",
                "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "peak_kind": f"{peak_kind} copy bandwidth",
                "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": traffic,
                "us_per_launch": round(sc_ms * 1e3, 2), "algorithmic_bytes": sc_bytes,
                "bytes_formula": "256*E_d + 256*N + 4*(N+1)"}
    # the same kernel at the AtomConv size of the 10,000-atom cell (84 edges per atom), on
    # synthetic uniform segments — the size the north-star's >= 50 % target is quoted for
    if args.workload != "c4":
        class _B:  # minimal stand-in carrying the three fields time_scatter_kernel reads
            z = batch.z
            n_atoms, n_edges = 10000, 840000
            ptr_c = (torch.arange(10001, device=dev, dtype=torch.int32) * 84).contiguous()
        ms10, b10 = time_scatter_kernel(K, _B)
        roofline["at_10k_atoms"] = {"us_per_launch": round(ms10 * 1e3, 2), "algorithmic_bytes": b10,
                                    "achieved": round(b10 / (ms10 * 1e-3) / 1e9, 1),
                                    "frac": round(b10 / (ms10 * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                                    "input": "synthetic: 10,000 segments x 84 rows x 256 B"}
    # per-kernel shares (own events, outside the timed region)
    ek = EventKernels(K)
    Engine(engine.pw, ek).run(batch, need_grad=True)
    shares = ek.table()

    # ---------------- CPU baseline: oracle port on the host cores ----------------
    cpu = None
    torch_cuda = None
    if not args.no_cpu_baseline:
        from oracle import chgnet_oracle as orc

        w = orc.load_weights_npz(WEIGHTS)
        if args.workload == "c4":
            from chgnet_b200 import graphgen

            z, frac, lat = graphgen.limno2_structure((5, 4, 3), 0.02, 4000)
            sample = [graphgen.make_crystal_graph(z, frac, lat)]
            sdesc = "LiMnO2 5x4x3 (480 atoms), 1 warm-up + 2 timed; structures/s scaled by atoms to the 10,000-atom cell"
        else:
            sample = graphs[: min(len(graphs), args.cpu_sample)]
            sdesc = f"first {len(sample)} graphs of the batch, 1 warm-up + 2 timed predict_graph(task='efs') calls"
        probe = sample[:2]
        threads = pick_cpu_threads(lambda: orc.predict_graph(w, probe, "efs", batch_size=len(probe)))
        sdesc += f"; {threads} of {os.cpu_count()} host threads (fastest of a 4..all sweep)"
        orc.predict_graph(w, sample, "efs", batch_size=len(sample))
        t0 = time.perf_counter()
        for _ in range(2):
            orc.predict_graph(w, sample, "efs", batch_size=len(sample))
        dt = (time.perf_counter() - t0) / 2
        cs = counts(sample)
        v = cs["graphs"] / dt if args.workload != "c4" else (cs["atoms"] / dt) / c["atoms"]
        cpu = {"value": v, "unit": "structures/s", "atoms_per_s": cs["atoms"] / dt, "cores": torch.get_num_threads(),
               "kind": "port", "sample": sdesc}
        # the realistic incumbent: the reference's torch ops on the SAME B200 (stock PyTorch CUDA)
        if args.workload != "c4":
            try:
                for _ in range(2):
                    orc.predict_graph(w, graphs, "efs", batch_size=len(graphs), device=dev)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(3):
                    orc.predict_graph(w, graphs, "efs", batch_size=len(graphs), device=dev)
                torch.cuda.synchronize()
                dtc = (time.perf_counter() - t0) / 3
                torch_cuda = {"value": c["graphs"] / dtc, "unit": "structures/s", "ms_per_step": dtc * 1e3,
                              "what": "oracle port = the reference's torch ops and per-graph batching loop, stock PyTorch "
                                      "CUDA on the same B200, fp32, host graphs in / numpy out (compare with e2e)"}
            except Exception as exc:  # reported, never fatal for the bench line
                torch_cuda = {"unavailable": repr(exc)[:200]}

    total_graphs = c["graphs"] * world
    value = total_graphs / (ms_per_step * 1e-3)
    line = {
        "metric": "structures_per_sec_EFS", "value": value, "unit": "structures/s",
        "atoms_per_s": c["atoms"] * world / (ms_per_step * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "task": "efs", "per_gpu": c, "weights": "CHGNet 0.3.0",
                   "l2": "256 MiB buffer written, then 256 MiB read (clean lines), between timed iterations", "parallelism": f"graph-sharded x{world}",
                   "engine": "native chg_forward (one C call per step); kernel_shares via the Python schedule of the same kernels"},
        "e2e": {"value": total_graphs / (e2e_ms_per_step * 1e-3), "unit": "structures/s",
                "ms_per_step": e2e_ms_per_step, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "CHGNet.predict_graph(list[CrystalGraph] on host, task='efs')", "breakdown": breakdown},
        "gpu_launches": int(launches), "wall_ms_timed_region": wall_ms,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "torch_cuda_baseline": torch_cuda,
        "kernel_shares": shares,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CHGNET_BENCH_WORKLOAD", "c2"), choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--cpu-sample", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scatter-only", action="store_true", help="run only the AtomConv scatter kernel timing (ncu target)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        if args.workload == "c5":
            if rank == 0:
                run_reference_train(args)
            return
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the hot path has no CPU implementation")
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        (run_train if args.workload == "c5" else run_ours)(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
