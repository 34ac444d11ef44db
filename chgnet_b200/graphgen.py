"""Host-side synthetic structures and a vectorised CrystalGraph builder.

This is the data generator for tests and ``bench.py``: it makes
the ``CrystalGraph`` objects that are fed — unchanged — to both the CUDA path
and the oracle.  It restates the *semantics* of the reference converter
(reference chgnet/graph/converter.py:102-190 and graph.py:132-328) with numpy
array operations instead of Python objects:

* periodic neighbour list: every (center, neighbour, image) with
  1e-8 < d <= r, grouped by center (what ``Structure.get_neighbor_list`` yields,
  converter.py:132);
* undirected bonds: a directed edge (c, n, img) and its reverse (n, c, -img)
  share one undirected index, numbered by first appearance (graph.py:132-224);
  ``undirected2directed`` points at the first of the two (graph.py:287);
* bond graph: for every undirected bond with d <= cutoff and each of its two
  ends, one row per *other* directed edge with d < cutoff leaving that end
  (graph.py:283-327).

The row order inside one center differs from the reference's dict-iteration
order; nothing downstream depends on it (sums only).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from chgnet_b200.graph import TORCH_DTYPE, CrystalGraph

# species pool of SURVEY.md §8d (all Z <= 94)
SPECIES_POOL = (3, 8, 11, 12, 13, 14, 15, 22, 25, 26, 27, 28)

# examples/mp-18767-LiMnO2.cif of the reference: P1 orthorhombic, 8 sites.
LIMNO2_ABC = (2.868779, 4.634475, 5.832507)
LIMNO2_Z = (3, 3, 25, 25, 8, 8, 8, 8)
LIMNO2_FRAC = (
    (0.5, 0.5, 0.37975050),
    (0.0, 0.0, 0.62024950),
    (0.5, 0.5, 0.86325250),
    (0.0, 0.0, 0.13674750),
    (0.5, 0.0, 0.36082450),
    (0.0, 0.5, 0.09851350),
    (0.5, 0.0, 0.90148650),
    (0.0, 0.5, 0.63917550),
)


def neighbor_list(frac: np.ndarray, lattice: np.ndarray, r: float, tol: float = 1e-8):
    """All (center, neighbour, image, distance) with tol < d <= r, center-grouped."""
    from scipy.spatial import cKDTree

    frac = np.asarray(frac, dtype=np.float64)
    lattice = np.asarray(lattice, dtype=np.float64)
    n = len(frac)
    vol = abs(np.linalg.det(lattice))
    # distance between lattice planes along each axis -> number of images needed
    heights = [
        vol / np.linalg.norm(np.cross(lattice[(k + 1) % 3], lattice[(k + 2) % 3]))
        for k in range(3)
    ]
    lo = np.floor(frac.min(axis=0)) if n else np.zeros(3)
    hi = np.ceil(frac.max(axis=0)) if n else np.ones(3)
    reps = [math.ceil(r / h) for h in heights]
    ranges = [
        np.arange(-reps[k] - int(hi[k] - lo[k]), reps[k] + int(hi[k] - lo[k]) + 1)
        for k in range(3)
    ]
    images = np.array(np.meshgrid(*ranges, indexing="ij")).reshape(3, -1).T
    cart = frac @ lattice
    shifts = images @ lattice  # [n_img, 3]
    all_pos = (cart[None, :, :] + shifts[:, None, :]).reshape(-1, 3)
    tree = cKDTree(all_pos)
    ctree = cKDTree(cart)
    sm = ctree.sparse_distance_matrix(tree, r * (1 + 1e-12), output_type="coo_matrix")
    center = sm.row.astype(np.int64)
    flat = sm.col.astype(np.int64)
    dist = sm.data
    # scipy drops exact zeros from the sparse matrix; enforce the tolerance anyway
    keep = (dist > tol) & (dist <= r)
    center, flat, dist = center[keep], flat[keep], dist[keep]
    img_idx, neighbor = np.divmod(flat, n)
    image = images[img_idx]
    order = np.lexsort((image[:, 2], image[:, 1], image[:, 0], neighbor, center))
    return center[order], neighbor[order], image[order], dist[order]


def build_graph_arrays(center, neighbor, image, distance, bond_cutoff: float):
    """Edge pairing + line graph, vectorised.  Returns numpy index arrays."""
    center = np.asarray(center, dtype=np.int64)
    neighbor = np.asarray(neighbor, dtype=np.int64)
    image = np.asarray(image, dtype=np.int64).reshape(-1, 3)
    distance = np.asarray(distance, dtype=np.float64)
    n_dir = len(center)
    if n_dir == 0:
        z = np.zeros(0, dtype=np.int64)
        return (np.zeros((0, 2), np.int64), z, z, np.zeros((0, 5), np.int64))
    # canonical orientation of each directed edge
    first_nz = np.where(
        image[:, 0] != 0, image[:, 0], np.where(image[:, 1] != 0, image[:, 1], image[:, 2])
    )
    fwd = (center < neighbor) | ((center == neighbor) & (first_nz > 0))
    a = np.where(fwd, center, neighbor)
    b = np.where(fwd, neighbor, center)
    im = np.where(fwd[:, None], image, -image)
    keys = np.column_stack([a, b, im])
    _, first_idx, inverse, counts = np.unique(
        keys, axis=0, return_index=True, return_inverse=True, return_counts=True
    )
    inverse = inverse.reshape(-1)
    if not np.all(counts == 2):
        raise ValueError(
            "directed edges are not complete: some undirected bond does not have "
            "exactly 2 directed edges"
        )
    rank = np.empty(len(first_idx), dtype=np.int64)
    rank[np.argsort(first_idx, kind="stable")] = np.arange(len(first_idx))
    d2u = rank[inverse]
    u2d = np.sort(first_idx)
    atom_graph = np.column_stack([center, neighbor])

    # ---- line graph ----
    short = np.nonzero(distance < bond_cutoff)[0]
    if len(short) == 0:
        return atom_graph, d2u, u2d, np.zeros((0, 5), np.int64)
    order = np.argsort(center[short], kind="stable")
    short = short[order]
    sc = center[short]
    # group boundaries by center
    change = np.nonzero(np.diff(sc))[0] + 1
    starts = np.concatenate([[0], change])
    sizes = np.diff(np.concatenate([starts, [len(short)]]))
    grp = np.repeat(np.arange(len(starts)), sizes)
    g_start = starts[grp]
    g_size = sizes[grp]
    local = np.arange(len(short)) - g_start
    reps = g_size - 1
    total = int(reps.sum())
    if total == 0:
        return atom_graph, d2u, u2d, np.zeros((0, 5), np.int64)
    i_pos = np.repeat(np.arange(len(short)), reps)
    blk_start = np.concatenate([[0], np.cumsum(reps)[:-1]])
    t = np.arange(total) - np.repeat(blk_start, reps)
    j_local = t + (t >= local[i_pos])
    j_pos = g_start[i_pos] + j_local
    e_i = short[i_pos]
    e_j = short[j_pos]
    # reference keeps bond i only when d_i <= cutoff (graph.py:290) and j when
    # d_j < cutoff (graph.py:316); `short` is the strict set, equal for d != cutoff
    rows = np.column_stack([center[e_i], d2u[e_i], e_i, d2u[e_j], e_j])
    second = (u2d[d2u[e_i]] != e_i).astype(np.int64)
    order = np.argsort(d2u[e_i] * 2 + second, kind="stable")
    return atom_graph, d2u, u2d, rows[order]


def native_graph_arrays(frac: np.ndarray, lattice: np.ndarray, atom_graph_cutoff: float, bond_graph_cutoff: float):
    """The same arrays as ``neighbor_list`` + ``build_graph_arrays`` from the C++ builder of the kernel
    library (``chg_graph_build``, csrc/graph_builder.cu — host code, no GPU needed): atom_graph [Ed,2],
    image [Ed,3] (fp32), d2u, u2d, bond_graph [A,5] (int32)."""
    import ctypes

    from chgnet_b200._lib import ChgnetB200Error, load_library

    lib = load_library()
    if not getattr(lib, "_graph_bound", False):
        lib.chg_graph_build.restype = ctypes.c_int32
        lib.chg_graph_build.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_double, ctypes.c_double,
                                        ctypes.POINTER(ctypes.c_void_p)]
        lib.chg_graph_sizes.restype = None
        lib.chg_graph_sizes.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int64)] * 3
        lib.chg_graph_export.restype = ctypes.c_int32
        lib.chg_graph_export.argtypes = [ctypes.c_void_p] * 6
        lib.chg_graph_free.restype = None
        lib.chg_graph_free.argtypes = [ctypes.c_void_p]
        lib._graph_bound = True
    frac = np.ascontiguousarray(frac, dtype=np.float64).reshape(-1, 3)
    lattice = np.ascontiguousarray(lattice, dtype=np.float64).reshape(3, 3)
    handle = ctypes.c_void_p()
    rc = lib.chg_graph_build(frac.ctypes.data, lattice.ctypes.data, len(frac), float(atom_graph_cutoff),
                             float(bond_graph_cutoff), ctypes.byref(handle))
    try:
        if rc != 0:
            msg = lib.chg_last_error().decode()
            raise (ValueError if "not complete" in msg else ChgnetB200Error)(msg)
        ne, nb, na = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        lib.chg_graph_sizes(handle, ctypes.byref(ne), ctypes.byref(nb), ctypes.byref(na))
        ag = np.empty((ne.value, 2), np.int32)
        img = np.empty((ne.value, 3), np.float32)
        d2u, u2d = np.empty(ne.value, np.int32), np.empty(nb.value, np.int32)
        bg = np.empty((na.value, 5), np.int32)
        lib.chg_graph_export(handle, ag.ctypes.data, img.ctypes.data, d2u.ctypes.data, u2d.ctypes.data, bg.ctypes.data)
    finally:
        if handle:
            lib.chg_graph_free(handle)
    return ag, img, d2u, u2d, bg


def make_crystal_graph(
    atomic_numbers,
    frac,
    lattice,
    *,
    atom_graph_cutoff: float = 6.0,
    bond_graph_cutoff: float = 3.0,
    graph_id: str | None = None,
    backend: str = "native",
) -> CrystalGraph:
    """numpy structure -> CrystalGraph with the reference's dtypes.

    ``backend="native"`` (default): the C++ builder of the kernel library; ``"numpy"``: the vectorised
    restatement above (the builder's own checker, pinned against the reference's ``Graph`` class)."""
    frac = np.asarray(frac, dtype=np.float64)
    lattice = np.asarray(lattice, dtype=np.float64)
    if backend == "native":
        ag, img, d2u, u2d, bg = native_graph_arrays(frac, lattice, atom_graph_cutoff, bond_graph_cutoff)
    elif backend == "numpy":
        c, n, img, d = neighbor_list(frac, lattice, atom_graph_cutoff)
        ag, d2u, u2d, bg = build_graph_arrays(c, n, img, d, bond_graph_cutoff)
    else:
        raise ValueError(f"unknown {backend=}")
    as_i32 = (lambda a: torch.from_numpy(a)) if backend == "native" else (lambda a: torch.tensor(a, dtype=torch.int32))
    return CrystalGraph(
        atomic_number=torch.tensor(np.asarray(atomic_numbers), dtype=torch.int32),
        atom_frac_coord=torch.tensor(frac, dtype=TORCH_DTYPE),
        atom_graph=as_i32(ag).reshape(-1, 2),
        atom_graph_cutoff=atom_graph_cutoff,
        neighbor_image=(torch.from_numpy(img) if backend == "native" else torch.tensor(img, dtype=TORCH_DTYPE)).reshape(-1, 3),
        directed2undirected=as_i32(d2u),
        undirected2directed=as_i32(u2d),
        bond_graph=as_i32(bg).reshape(-1, 5),
        bond_graph_cutoff=bond_graph_cutoff,
        lattice=torch.tensor(lattice, dtype=TORCH_DTYPE),
        graph_id=graph_id,
    )


# --------------------------------------------------------------------------
# synthetic structures
# --------------------------------------------------------------------------
def limno2_structure(supercell=(1, 1, 1), displacement: float = 0.0, seed: int = 0):
    """LiMnO2 mp-18767 (optionally a supercell with Gaussian displacements)."""
    sx, sy, sz = supercell
    base = np.array(LIMNO2_FRAC, dtype=np.float64)
    cells = np.array(np.meshgrid(range(sx), range(sy), range(sz), indexing="ij")).reshape(3, -1).T
    frac = (base[None, :, :] + cells[:, None, :]).reshape(-1, 3) / np.array([sx, sy, sz])
    z = np.tile(np.array(LIMNO2_Z), len(cells))
    lattice = np.diag(np.array(LIMNO2_ABC) * np.array([sx, sy, sz]))
    if displacement > 0:
        rng = np.random.default_rng(seed)
        cart = frac @ lattice + rng.normal(0.0, displacement, size=frac.shape)
        frac = cart @ np.linalg.inv(lattice)
    return z, frac, lattice


def random_structure(n_atoms: int, seed: int, density: float = 0.10, d_min: float = 1.6):
    """Random periodic cell: cubic a=(n/rho)^(1/3), symmetric strain U(-0.1,0.1),
    uniform positions with hard-sphere rejection under PBC."""
    rng = np.random.default_rng(seed)
    a = (n_atoms / density) ** (1.0 / 3.0)
    eps = np.zeros((3, 3))
    iu = np.triu_indices(3)
    eps[iu] = rng.uniform(-0.1, 0.1, size=6)
    eps = eps + eps.T - np.diag(np.diag(eps))
    lattice = a * (np.eye(3) + eps)
    z = rng.choice(np.array(SPECIES_POOL), size=n_atoms)
    shifts = np.array(np.meshgrid(*[(-1, 0, 1)] * 3, indexing="ij")).reshape(3, -1).T
    frac = np.zeros((0, 3))
    tries = 0
    while len(frac) < n_atoms:
        tries += 1
        if tries > 200000:
            raise RuntimeError("hard-sphere packing failed")
        cand = rng.uniform(0.0, 1.0, size=3)
        if len(frac):
            delta = (frac - cand)[None, :, :] + shifts[:, None, :]
            dist = np.linalg.norm(delta @ lattice, axis=-1)
            if dist.min() < d_min:
                continue
        frac = np.vstack([frac, cand])
    return z, frac, lattice


def random_graphs(n_graphs: int, n_lo: int, n_hi: int, seed0: int, **cut) -> list[CrystalGraph]:
    """``n_graphs`` random cells with n_i ~ U{n_lo..n_hi}, seeds seed0 + i."""
    out = []
    for i in range(n_graphs):
        rng = np.random.default_rng(seed0 + i)
        n = int(rng.integers(n_lo, n_hi + 1))
        z, frac, lat = random_structure(n, seed0 + i)
        out.append(make_crystal_graph(z, frac, lat, graph_id=f"rand{seed0 + i}", **cut))
    return out
