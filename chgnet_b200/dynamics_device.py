"""MD / relaxation with the whole step on the device.

The reference's MD loop is ``ase.md`` driving ``CHGNetCalculator.calculate`` (reference chgnet/model/dynamics.py:
129-181): every step converts the ASE atoms to a pymatgen ``Structure``, rebuilds the ``CrystalGraph`` on the CPU
(dynamics.py:156-157), copies it to the GPU, runs the model and copies energy / forces / stress back.  Here positions,
velocities and forces never leave the device:

* ``DeviceGraphBuilder`` (csrc/graph_device.cu) rebuilds the neighbour list / bond graph on the GPU.  Two policies:
  ``skin = 0`` - rebuild every step between the drift and the model call, exact cutoffs (the right choice for large
  cells: the build costs 0.5 ms for 10,000 atoms while a skin inflates the model's work, angles grow like (3 + skin)^6);
  ``skin > 0`` - a Verlet SKIN: the lists are built with cutoffs ``r + skin`` and reused until some atom has moved more
  than ``skin / 2`` since the last build, which lets the whole step be ONE CUDA graph replay (the right choice for small
  cells, where the ~140 launches of a step cost more than the kernels).  ``skin=None`` picks by the number of atoms.  Pairs beyond the model's cutoffs contribute exactly zero - CHGNet's polynomial envelope and hence the
  bond weights w_ag / w_bg vanish for d >= r_c (reference basis.py:184-205, layers.py:118-126, 245-254) - so the skin
  changes no result (tests/test_dynamics_device_gpu.py compares against rebuilding every step);
* between rebuilds the step = [half kick + drift] -> ``chg_forward`` -> [half kick] is ONE CUDA graph replay
  (captured after every rebuild): ~140 kernel launches become one;
* the host reads one double per step (the skin test) and, when asked, the energies.

``DeviceMD`` = velocity Verlet (NVE, what ``ase.md.verlet.VelocityVerlet`` does); ``DeviceFIRE`` = FIRE relaxation at
fixed cell with its state on the device.  The host drivers of ``chgnet_b200.dynamics`` are the checkers.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from chgnet_b200._lib import ChgnetB200Error, load_library
from chgnet_b200.dynamics import ATOMIC_MASSES, FS, KB
from chgnet_b200.graph_device import DeviceGraphBuilder

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        lib = load_library()
        P, I, D = ctypes.c_void_p, ctypes.c_int32, ctypes.c_double
        lib.chg_md_kick_drift.restype = I
        lib.chg_md_kick_drift.argtypes = [P, P, P, P, I, D, P, P, P, P, P, P]
        lib.chg_md_kick.restype = I
        lib.chg_md_kick.argtypes = [P, P, P, I, D, P, P]
        lib.chg_fire_step.restype = I
        lib.chg_fire_step.argtypes = [P, P, P, I, P, P, P, P, D, D, P]
        _LIB = lib
    return _LIB


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise ChgnetB200Error(f"{what} failed ({rc}): {_lib().chg_last_error().decode()}")


class _DeviceSystem:
    """Positions / velocities / forces on the device + the model evaluation with a skin-managed graph."""

    def __init__(self, model, numbers, positions, cell, *, skin: float | None = 0.5, use_cuda_graph: bool = True) -> None:
        self.model = model.eval()
        self.dev = model.device
        if self.dev.type != "cuda":
            raise ChgnetB200Error("device MD needs the model on a CUDA device")
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self.numbers = np.asarray(numbers, dtype=np.int64).reshape(-1)
        self.n = len(self.numbers)
        self.cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
        self.inv_cell = np.ascontiguousarray(np.linalg.inv(self.cell))
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.x = torch.as_tensor(np.asarray(positions, dtype=np.float64).reshape(-1, 3)).to(self.dev).contiguous()
        self.v = torch.zeros(self.n, 3, **f64)
        self.f = torch.zeros(self.n, 3, **f64)
        self.inv_mass = torch.as_tensor(1.0 / ATOMIC_MASSES[self.numbers - 1]).to(self.dev)
        self.frac64 = torch.empty(self.n, 3, **f64)
        self.frac32 = torch.empty(self.n, 3, dtype=torch.float32, device=self.dev)
        self.x_ref = self.x.clone()
        self.max_disp2 = torch.zeros(1, **f64)
        self.e_kin = torch.zeros(1, **f64)
        if skin is None:  # large cells: exact lists every step; small cells: skin + one CUDA graph per step
            skin = 0.0 if self.n >= 2000 else 0.5
        self.skin = float(skin)
        if self.skin < 0.0:
            raise ValueError("skin must be >= 0")
        self.every_step = self.skin == 0.0
        gc = model.graph_converter
        self.builder = DeviceGraphBuilder(self.dev, float(gc.atom_graph_cutoff) + self.skin, float(gc.bond_graph_cutoff) + self.skin)
        self.compact = not model._arch.get("mlp_out_bias", False)
        self.use_cuda_graph = use_cuda_graph and not self.every_step
        self.batch = None
        self.graph = None
        self.out: dict | None = None
        self.n_builds = 0
        self.n_steps = 0
        self.n_captures = 0
        self.t_rebuild = self.t_capture = 0.0  # host seconds spent rebuilding lists / capturing step graphs
        self._update_frac()
        self._rebuild()
        self._forward()
        self._old_graph = None

    # ---- pieces ---------------------------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _update_frac(self) -> None:
        fr = self.x @ torch.as_tensor(self.inv_cell).to(self.dev)
        self.frac64.copy_(fr)
        self.frac32.copy_(fr.to(torch.float32))

    def _rebuild(self) -> None:
        """New neighbour lists (cutoffs + skin) from the current positions; the captured step graph is dropped."""
        batch = self.builder.build_batch(self.numbers, self.frac64, self.cell, with_reverse=True, compact_bonds=self.compact)
        batch.frac = self.frac32  # the model reads the positions the integrator writes
        self.batch = batch
        self.x_ref.copy_(self.x)
        self.max_disp2.zero_()
        self._old_graph, self.graph = self.graph, None  # keep the pool's last user alive until the next capture
        self.n_builds += 1

    def _forward(self) -> None:
        nat = self.model._get_native()
        self.out = nat(self.batch, need_grad=True)
        self.f.copy_(self.out["force"])

    def needs_rebuild(self) -> bool:
        """One double from the device: has some atom moved more than skin / 2 since the lists were built?"""
        return float(self.max_disp2.item()) > (0.5 * self.skin) ** 2

    @property
    def potential_energy(self) -> float:
        return float((self.out["energy"] + self.out["e_ref"])[0].item())

    def positions(self) -> np.ndarray:
        return self.x.cpu().numpy()


class DeviceMD(_DeviceSystem):
    """Velocity-Verlet NVE with positions, velocities and forces resident on the device; ``timestep`` in fs."""

    def __init__(self, model, numbers, positions, cell, *, timestep: float = 2.0, velocities=None, skin: float | None = None,
                 use_cuda_graph: bool = True) -> None:
        super().__init__(model, numbers, positions, cell, skin=skin, use_cuda_graph=use_cuda_graph)
        self.dt = float(timestep) * FS
        if velocities is not None:
            self.v.copy_(torch.as_tensor(np.asarray(velocities, dtype=np.float64)).to(self.dev))

    def set_temperature(self, kelvin: float, seed: int = 0) -> None:
        rng = np.random.default_rng(seed)
        m = ATOMIC_MASSES[self.numbers - 1][:, None]
        v = rng.normal(size=(self.n, 3)) * np.sqrt(KB * kelvin / m)
        v -= (m * v).sum(axis=0) / m.sum()
        self.v.copy_(torch.as_tensor(v).to(self.dev))

    def _step_body(self) -> None:
        lib, st = _lib(), self._stream()
        with torch.cuda.device(self.dev):
            _check(lib.chg_md_kick_drift(self.x.data_ptr(), self.v.data_ptr(), self.f.data_ptr(), self.inv_mass.data_ptr(), self.n, self.dt,
                                         self.inv_cell.ctypes.data, self.frac64.data_ptr(), self.frac32.data_ptr(),
                                         self.x_ref.data_ptr(), self.max_disp2.data_ptr(), st), "chg_md_kick_drift")
        if self.every_step:  # exact neighbour lists for the new positions (0.5 ms for 10,000 atoms)
            self._rebuild()
        self._forward()
        self.e_kin.zero_()
        with torch.cuda.device(self.dev):
            _check(lib.chg_md_kick(self.v.data_ptr(), self.f.data_ptr(), self.inv_mass.data_ptr(), self.n, self.dt, self.e_kin.data_ptr(),
                                   self._stream()), "chg_md_kick")

    def step(self) -> None:
        if self.use_cuda_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._step_body()
        self.n_steps += 1
        if not self.every_step and self.needs_rebuild():
            import time

            t0 = time.perf_counter()
            self._rebuild()
            self.t_rebuild += time.perf_counter() - t0

    def _capture(self) -> None:
        """Capture [kick + drift] -> chg_forward -> [kick] for the current topology.  Raw capture_begin / capture_end
        on a side stream with one memory pool shared by all captures of this run (``torch.cuda.graph`` would run the
        garbage collector and empty the allocator cache at every capture)."""
        import time

        t0 = time.perf_counter()
        nat = self.model._get_native()
        nat.reserve(self.batch, need_grad=True)  # never reallocate the workspace inside a capture
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(self.dev)
            self._pool = torch.cuda.graph_pool_handle()
        s = self._side
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            if not getattr(self, "_warmed", False):
                # once: a real step on the side stream (lazily set kernel attributes, allocator warm-up), then undone
                state = (self.x.clone(), self.v.clone(), self.f.clone(), self.max_disp2.clone())
                self._step_body()
                s.synchronize()
                self.x.copy_(state[0]), self.v.copy_(state[1]), self.f.copy_(state[2]), self.max_disp2.copy_(state[3])
                self._warmed = True
            # the previous capture stays alive until the new one exists: the shared pool must never drop to zero users
            g = torch.cuda.CUDAGraph()
            g.capture_begin(pool=self._pool, capture_error_mode="thread_local")  # other threads (NCCL watchdog) may call CUDA
            try:
                self._step_body()
            finally:
                g.capture_end()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        self._old_graph = None
        self.graph = g  # capturing executed nothing: the state is still the pre-step state
        self.n_captures += 1
        self.t_capture += time.perf_counter() - t0

    @property
    def kinetic_energy(self) -> float:
        return float(self.e_kin.item())

    def temperature(self) -> float:
        return 2.0 * self.kinetic_energy / (3.0 * self.n * KB)

    def run(self, steps: int, log_every: int = 1) -> list[dict]:
        log = []
        for i in range(steps):
            self.step()
            if log_every and (i + 1) % log_every == 0:
                log.append({"step": self.n_steps, "e_pot": self.potential_energy, "e_kin": self.kinetic_energy,
                            "temperature": self.temperature()})
        return log


class DeviceFIRE(_DeviceSystem):
    """FIRE relaxation at fixed cell (the reference's default optimizer, dynamics.py:190-204) with the optimizer state
    on the device; the host looks at the largest force every ``check_every`` steps."""

    def __init__(self, model, numbers, positions, cell, *, dt: float = 0.1, dt_max: float = 1.0, max_step: float = 0.2,
                 skin: float | None = None) -> None:
        super().__init__(model, numbers, positions, cell, skin=skin, use_cuda_graph=False)
        self.dt_max, self.max_step = float(dt_max), float(max_step)
        self.state = torch.zeros(12, dtype=torch.float64, device=self.dev)
        self.state[0], self.state[1] = dt, 0.1

    def run(self, fmax: float = 0.1, steps: int = 500, check_every: int = 5) -> dict:
        lib = _lib()
        energies = []
        fm = float(self.f.pow(2).sum(dim=1).max().sqrt().item())
        it = 0
        while it < steps and fm >= fmax:
            with torch.cuda.device(self.dev):
                _check(lib.chg_fire_step(self.x.data_ptr(), self.v.data_ptr(), self.f.data_ptr(), self.n, self.state.data_ptr(),
                                         self.inv_cell.ctypes.data, self.frac64.data_ptr(), self.frac32.data_ptr(), self.dt_max,
                                         self.max_step, self._stream()), "chg_fire_step")
            # skin test on the host side of the loop (positions moved by at most max_step)
            if self.every_step:
                self._rebuild()
            else:
                d2 = float(((self.x - self.x_ref) ** 2).sum(dim=1).max().item())
                if d2 > (0.5 * self.skin) ** 2:
                    self._rebuild()
            self._forward()
            it += 1
            if it % check_every == 0 or it == steps:
                fm = float(self.f.pow(2).sum(dim=1).max().sqrt().item())
                energies.append(self.potential_energy)
        return {"energies": energies, "fmax": fm, "steps": it, "converged": fm < fmax, "graph_builds": self.n_builds}


__all__ = ["DeviceMD", "DeviceFIRE"]
