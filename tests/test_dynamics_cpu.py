"""CPU: the calculator shim and the host MD / relaxation drivers (reference chgnet/model/dynamics.py:58-181)
with the torch kernel specifications injected in place of the CUDA library."""
import os

import numpy as np
import pytest

from chgnet_b200 import graphgen
from chgnet_b200.batch import build_batch  # noqa: F401
from chgnet_b200.dynamics import GPA, Atoms, CHGNetCalculator, VelocityVerlet, fire_relax
from chgnet_b200.engine import Engine
from chgnet_b200.weights import pack_weights
from oracle.kernel_specs import SpecKernels


@pytest.fixture()
def calc(monkeypatch):
    from chgnet_b200.model import CHGNet

    path = os.path.join(os.path.dirname(__file__), "golden", "chgnet_0.3.0_weights.npz")
    model = CHGNet.from_file(path, version="0.3.0")
    eng = Engine(pack_weights(model.state_dict(), model.model_args, device="cpu"), SpecKernels())
    monkeypatch.setattr(model, "_get_engine", lambda: eng)
    monkeypatch.setenv("CHGNET_B200_ENGINE", "python")
    return CHGNetCalculator(model=model, return_site_energies=True)


def _limno2(displacement=0.0, seed=0):
    z, frac, lat = graphgen.limno2_structure((1, 1, 1), displacement, seed)
    return Atoms(z, frac @ lat, lat)


def test_calculator_results_follow_the_reference_conventions(calc, golden):
    atoms = _limno2()
    calc.calculate(atoms)
    r = calc.results
    assert r["energy"] == pytest.approx(-7.36769 * 8, abs=1e-3) and r["free_energy"] == r["energy"]  # tests/test_model.py:68, extensive
    assert r["forces"].shape == (8, 3) and r["stress"].shape == (3, 3) and r["magmoms"].shape == (8,)
    assert np.allclose(r["stress"], golden["limno2.ref32.s"] * GPA, atol=2e-3 * GPA)  # GPa -> eV/A^3 (dynamics.py:69)
    assert np.allclose(r["forces"], golden["limno2.ref32.f"], atol=1e-3)
    assert r["energies"].shape == (8,) and r["crystal_fea"].shape == (64,)
    assert calc.n_params == 412525 and calc.version == "0.3.0"
    calc.calculate(atoms, task="e")
    assert "energy" in calc.results


def test_relaxation_and_nve_dynamics(calc):
    atoms = _limno2(0.04, seed=3)
    calc.calculate(atoms, task="ef")
    e0, f0 = float(calc.results["energy"]), float(np.abs(calc.results["forces"]).max())
    out = fire_relax(atoms, calc, fmax=0.02, steps=12)
    assert out["energies"][-1] < e0 - 1e-4 and out["fmax"] < f0  # downhill
    md_atoms = _limno2(0.02, seed=5)
    md = VelocityVerlet(md_atoms, calc, timestep=1.0)
    md.set_temperature(300.0, seed=1)
    e_start = md.potential_energy() + md.kinetic_energy()
    log = md.run(6)
    e_end = log[-1]["e_pot"] + log[-1]["e_kin"]
    assert abs(e_end - e_start) < 5e-3, (e_start, e_end)  # eV for 8 atoms over 6 fs
    assert 50 < log[-1]["temperature"] < 1000 and md.nsteps == 6


def test_isolated_atoms_policy(calc):
    atoms = Atoms([3, 8], [[0.0, 0, 0], [10.0, 10, 10]], np.eye(3) * 20.0)  # both atoms isolated: no edges at all
    calc.calculate(atoms, task="e")  # nothing to warn about when the graph has no edges (model.py:841-843)
    far = Atoms([3, 8, 8], [[0.0, 0, 0], [1.5, 0, 0], [10.0, 10, 10]], np.eye(3) * 20.0)
    with pytest.warns(UserWarning, match="isolated atoms"):
        calc.calculate(far, task="e")
    calc.on_isolated_atoms = "error"
    with pytest.raises(ValueError):
        calc.calculate(far, task="e")
