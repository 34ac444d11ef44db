"""TEST INFRASTRUCTURE — container-only import shim for the real reference.

Imports the unmodified reference (`/root/reference/chgnet`) on a machine that has
no pymatgen / ase / monty, following the recipe of SURVEY.md §8c:

* a stub ``pymatgen.core`` that only provides the names the hot path uses for
  ``isinstance`` checks (reference chgnet/model/model.py:10, 581),
* bare package objects for ``chgnet.model`` and ``chgnet.utils`` so their
  ``__init__`` files (which pull in ase / monty) are skipped.

Nothing here is product code and nothing here runs on the GPU box
(`/root/reference` does not exist there).  It is used by
``oracle/make_golden.py`` to (a) pin ``oracle/chgnet_oracle.py`` against the live
reference and (b) write the fixtures under ``tests/golden/``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CHGNET_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "chgnet", "model"))


def _install_pymatgen_stub() -> None:
    if "pymatgen" in sys.modules:
        return
    pmg = types.ModuleType("pymatgen")
    core = types.ModuleType("pymatgen.core")
    structure = types.ModuleType("pymatgen.core.structure")

    class Structure:  # only ever used for isinstance()
        pass

    class Molecule:
        pass

    class Lattice:
        pass

    core.Structure = Structure
    core.Lattice = Lattice
    core.Molecule = Molecule
    structure.Structure = Structure
    structure.Molecule = Molecule
    pmg.core = core
    core.structure = structure
    sys.modules["pymatgen"] = pmg
    sys.modules["pymatgen.core"] = core
    sys.modules["pymatgen.core.structure"] = structure


def import_reference():
    """Return the reference's ``chgnet.model.model`` module (CPU)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_pymatgen_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import chgnet  # noqa: F401  (top-level __init__ is dependency-free)

    for name in ("chgnet.model", "chgnet.utils"):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(REFERENCE_ROOT, *name.split("."))]
            sys.modules[name] = pkg
    cu = importlib.import_module("chgnet.utils.common_utils")
    utils = sys.modules["chgnet.utils"]
    for attr in dir(cu):
        if not attr.startswith("_"):
            setattr(utils, attr, getattr(cu, attr))
    return importlib.import_module("chgnet.model.model")


def load_reference_model(model_name: str = "0.3.0"):
    mod = import_reference()
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        model = mod.CHGNet.load(model_name=model_name, use_device="cpu", verbose=False)
    return model.eval()
