// Host-native CrystalGraph builder: periodic neighbour list, undirected-bond pairing, bond (line) graph.
//
// This is the counterpart of the reference's only native component on this path — `create_graph.c` /
// `cygraph.pyx` (chgnet/graph/cygraph.pyx:69-175, create_graph.c:100-107) — plus the neighbour list the
// reference takes from pymatgen (converter.py:132), with the semantics of converter.py:102-190 and
// graph.py:132-328 as restated (and pinned against the reference's `Graph` class) by
// chgnet_b200/graphgen.py:
//   * directed edges: every (center, neighbour, image) with 1e-8 < d <= r_atom, sorted by
//     (center, neighbour, image);
//   * undirected bonds: (c, n, img) and (n, c, -img) share one index, numbered by first appearance;
//     undirected2directed points at the first of the two;
//   * bond graph: for every directed edge i with d < r_bond and every OTHER directed edge j with
//     d < r_bond leaving the same centre, one row (centre, u(i), i, u(j), j); rows sorted by
//     (u(i), whether i is the second edge of its bond), stable.
// Pure host code (no device work): a uniform grid over the replicated images makes the search
// O(N * neighbours), single-threaded (0.5 ms for 50 atoms, ~100 ms for 10 000).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "common.cuh"
#include "worker_pool.h"

struct chg_graph {
  std::vector<int32_t> atom_graph;  // [Ed][2]
  std::vector<float> image;         // [Ed][3]
  std::vector<int32_t> d2u, u2d;    // [Ed], [Eu]
  std::vector<int32_t> bond_graph;  // [A][5]
  int32_t n_isolated = 0;           // atoms without any neighbour inside r_atom
};

namespace chg {
namespace {

void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

}  // namespace
}  // namespace chg

using namespace chg;

extern "C" int chg_graph_build(const double* frac, const double* lattice, int32_t n_atoms, double r_atom, double r_bond,
                               chg_graph** out) {
  CHG_CHECK_ARG(n_atoms >= 0 && r_atom > 0 && r_bond >= 0, "bad size or cutoff");
  CHG_CHECK_ARG(lattice != nullptr && out != nullptr && (frac != nullptr || n_atoms == 0), "null pointer");
  const double tol = 1e-8;
  const double* L = lattice;  // rows = lattice vectors
  const int n = n_atoms;
  chg_graph* G = new chg_graph();
  *out = G;
  if (n == 0) return CHG_OK;

  const bool timing = std::getenv("CHG_GRAPH_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "chg_graph_build %-12s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  // ---- image range: distance between lattice planes -> repetitions (graphgen.neighbor_list) ----
  double c12[3], c20[3], c01[3];
  cross3(L + 3, L + 6, c12);
  cross3(L + 6, L + 0, c20);
  cross3(L + 0, L + 3, c01);
  const double vol = std::fabs(L[0] * c12[0] + L[1] * c12[1] + L[2] * c12[2]);
  CHG_CHECK_ARG(vol > 0, "singular lattice");
  const double* crosses[3] = {c12, c20, c01};
  int lo_i[3], hi_i[3];
  for (int k = 0; k < 3; ++k) {
    const double* c = crosses[k];
    const double height = vol / std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    double fmin = frac[k], fmax = frac[k];
    for (int i = 1; i < n; ++i) {
      fmin = std::min(fmin, frac[3 * i + k]);
      fmax = std::max(fmax, frac[3 * i + k]);
    }
    const int reps = (int)std::ceil(r_atom / height);
    const int span = (int)(std::ceil(fmax) - std::floor(fmin));
    lo_i[k] = -reps - span;
    hi_i[k] = reps + span;
  }
  std::vector<double> cart((size_t)n * 3);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 3; ++j)
      cart[3 * i + j] = frac[3 * i] * L[j] + frac[3 * i + 1] * L[3 + j] + frac[3 * i + 2] * L[6 + j];

  // ---- uniform grid (cell edge r_atom) over the centres' bounding box grown by r_atom; only image
  //      points inside it can be neighbours ----
  double bmin[3], bmax[3];
  for (int j = 0; j < 3; ++j) bmin[j] = bmax[j] = cart[j];
  for (int i = 1; i < n; ++i)
    for (int j = 0; j < 3; ++j) {
      bmin[j] = std::min(bmin[j], cart[3 * i + j]);
      bmax[j] = std::max(bmax[j], cart[3 * i + j]);
    }
  int64_t dims[3];
  for (int j = 0; j < 3; ++j) {
    bmin[j] -= r_atom * (1 + 1e-9);
    bmax[j] += r_atom * (1 + 1e-9);
    dims[j] = std::max<int64_t>(1, (int64_t)std::floor((bmax[j] - bmin[j]) / r_atom));
  }
  const double inv_cell[3] = {dims[0] / (bmax[0] - bmin[0]), dims[1] / (bmax[1] - bmin[1]), dims[2] / (bmax[2] - bmin[2])};
  auto bin_of = [&](const double* p, int64_t* ijk) {
    for (int j = 0; j < 3; ++j) ijk[j] = std::min<int64_t>(dims[j] - 1, std::max<int64_t>(0, (int64_t)((p[j] - bmin[j]) * inv_cell[j])));
  };
  struct Pt {
    double x, y, z;
    int32_t atom, i0, i1, i2;
  };
  std::vector<Pt> pts;
  std::vector<int64_t> pt_bin;
  for (int i0 = lo_i[0]; i0 <= hi_i[0]; ++i0)
    for (int i1 = lo_i[1]; i1 <= hi_i[1]; ++i1)
      for (int i2 = lo_i[2]; i2 <= hi_i[2]; ++i2) {
        const double sh[3] = {i0 * L[0] + i1 * L[3] + i2 * L[6], i0 * L[1] + i1 * L[4] + i2 * L[7],
                              i0 * L[2] + i1 * L[5] + i2 * L[8]};
        for (int a = 0; a < n; ++a) {
          const double p[3] = {cart[3 * a] + sh[0], cart[3 * a + 1] + sh[1], cart[3 * a + 2] + sh[2]};
          if (p[0] < bmin[0] || p[0] > bmax[0] || p[1] < bmin[1] || p[1] > bmax[1] || p[2] < bmin[2] || p[2] > bmax[2]) continue;
          int64_t ijk[3];
          bin_of(p, ijk);
          pts.push_back(Pt{p[0], p[1], p[2], a, i0, i1, i2});
          pt_bin.push_back((ijk[0] * dims[1] + ijk[1]) * dims[2] + ijk[2]);
        }
      }
  const int64_t n_bins = dims[0] * dims[1] * dims[2];
  std::vector<int64_t> bin_start((size_t)n_bins + 1, 0);
  for (int64_t b : pt_bin) ++bin_start[b + 1];
  for (int64_t b = 0; b < n_bins; ++b) bin_start[b + 1] += bin_start[b];
  std::vector<Pt> binned(pts.size());  // the points in bin order: a bin is one contiguous run
  {
    std::vector<int64_t> fill(bin_start.begin(), bin_start.end() - 1);
    for (size_t p = 0; p < pts.size(); ++p) binned[fill[pt_bin[p]]++] = pts[p];
  }
  pts.clear();
  pts.shrink_to_fit();

  lap("grid");
  // ---- neighbour search, centre by centre; a directed edge is keyed by (neighbour, image) packed into
  //      one integer so that sorting and the reverse-edge lookup are plain integer operations ----
  for (int k = 0; k < 3; ++k) CHG_CHECK_ARG(lo_i[k] >= -127 && hi_i[k] <= 127, "more than 127 periodic images along one axis");
  auto pack_key = [](int64_t atom, int i0, int i1, int i2) {
    return ((uint64_t)atom << 24) | ((uint64_t)(i0 + 128) << 16) | ((uint64_t)(i1 + 128) << 8) | (uint64_t)(i2 + 128);
  };
  struct Cand {
    uint64_t key;
    double d;
  };
  std::vector<uint64_t> keys;  // [Ed], sorted within each centre
  std::vector<double> dist;    // [Ed]
  std::vector<int64_t> first_edge((size_t)n + 1, 0);
  keys.reserve((size_t)n * 96);
  dist.reserve((size_t)n * 96);
  std::vector<Cand> cand;
  const double r2_max = r_atom * r_atom * (1 + 1e-12);
  for (int c = 0; c < n; ++c) {
    const double* pc = &cart[3 * c];
    int64_t ijk[3];
    bin_of(pc, ijk);
    cand.clear();
    for (int64_t bx = std::max<int64_t>(0, ijk[0] - 1); bx <= std::min(dims[0] - 1, ijk[0] + 1); ++bx)
      for (int64_t by = std::max<int64_t>(0, ijk[1] - 1); by <= std::min(dims[1] - 1, ijk[1] + 1); ++by)
        for (int64_t bz = std::max<int64_t>(0, ijk[2] - 1); bz <= std::min(dims[2] - 1, ijk[2] + 1); ++bz) {
          const int64_t b = (bx * dims[1] + by) * dims[2] + bz;
          for (int64_t q = bin_start[b]; q < bin_start[b + 1]; ++q) {
            const Pt& p = binned[q];
            const double dx = p.x - pc[0], dy = p.y - pc[1], dz = p.z - pc[2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 > r2_max) continue;
            const double d = std::sqrt(d2);
            if (d > tol && d <= r_atom) cand.push_back(Cand{pack_key(p.atom, p.i0, p.i1, p.i2), d});
          }
        }
    std::sort(cand.begin(), cand.end(), [](const Cand& x, const Cand& y) { return x.key < y.key; });
    for (const Cand& h : cand) {
      keys.push_back(h.key);
      dist.push_back(h.d);
    }
    first_edge[c + 1] = (int64_t)keys.size();
    if (first_edge[c + 1] == first_edge[c]) ++G->n_isolated;
  }
  lap("search");
  const size_t n_dir = keys.size();
  CHG_CHECK_ARG(n_dir < (size_t)INT32_MAX, "too many edges for int32 indices");
  G->atom_graph.resize(n_dir * 2);
  G->image.resize(n_dir * 3);
  G->d2u.resize(n_dir);
  std::vector<int32_t> ctr(n_dir);
  for (int c = 0; c < n; ++c)
    for (int64_t e = first_edge[c]; e < first_edge[c + 1]; ++e) {
      const uint64_t k = keys[e];
      G->atom_graph[2 * e] = c;
      G->atom_graph[2 * e + 1] = (int32_t)(k >> 24);
      G->image[3 * e] = (float)((int)((k >> 16) & 255) - 128);
      G->image[3 * e + 1] = (float)((int)((k >> 8) & 255) - 128);
      G->image[3 * e + 2] = (float)((int)(k & 255) - 128);
      ctr[e] = c;
    }
  lap("emit");
  // ---- undirected bonds, numbered by first appearance ----
  // the reverse of edge (c, n, img) is (n, c, -img): binary search among the (sorted) edges of n
  std::vector<int32_t> rev(n_dir);
  for (size_t e = 0; e < n_dir; ++e) {
    const uint64_t k = keys[e];
    const int64_t nb = (int64_t)(k >> 24);
    const uint64_t want = pack_key(ctr[e], 128 - (int)((k >> 16) & 255), 128 - (int)((k >> 8) & 255), 128 - (int)(k & 255));
    const uint64_t* lo = keys.data() + first_edge[nb];
    const uint64_t* hi = keys.data() + first_edge[nb + 1];
    const uint64_t* it = std::lower_bound(lo, hi, want);
    if (it == hi || *it != want) {
      set_error("chg_graph_build: directed edges are not complete: some undirected bond does not have exactly 2 directed edges");
      return CHG_ERR_ARG;
    }
    rev[e] = (int32_t)(it - keys.data());
  }
  G->u2d.reserve(n_dir / 2);
  for (size_t e = 0; e < n_dir; ++e) {
    if ((size_t)rev[e] > e) {  // first appearance of this bond
      G->d2u[e] = (int32_t)G->u2d.size();
      G->u2d.push_back((int32_t)e);
    } else {
      G->d2u[e] = G->d2u[rev[e]];
    }
  }

  lap("pairing");
  // ---- bond graph: rows are generated centre by centre; the final order (undirected bond of i, then
  //      first / second edge of that bond, stable) is a counting sort over 2 * Eu keys ----
  std::vector<int32_t> short_e;
  for (size_t e = 0; e < n_dir; ++e)
    if (dist[e] < r_bond) short_e.push_back((int32_t)e);
  const size_t n_keys = G->u2d.size() * 2;
  std::vector<int64_t> key_start(n_keys + 1, 0);
  auto key_of = [&](int32_t ei) { return (size_t)G->d2u[ei] * 2 + (G->u2d[G->d2u[ei]] != ei ? 1 : 0); };
  for (size_t s0 = 0; s0 < short_e.size();) {
    size_t t = s0;
    while (t < short_e.size() && ctr[short_e[t]] == ctr[short_e[s0]]) ++t;
    for (size_t i = s0; i < t; ++i) key_start[key_of(short_e[i]) + 1] += (int64_t)(t - s0 - 1);
    s0 = t;
  }
  for (size_t k = 0; k < n_keys; ++k) key_start[k + 1] += key_start[k];
  G->bond_graph.resize((size_t)key_start[n_keys] * 5);
  {
    std::vector<int64_t> fill(key_start.begin(), key_start.end() - 1);
    for (size_t s0 = 0; s0 < short_e.size();) {
      size_t t = s0;
      while (t < short_e.size() && ctr[short_e[t]] == ctr[short_e[s0]]) ++t;
      for (size_t i = s0; i < t; ++i) {
        const int32_t ei = short_e[i];
        const int32_t ui = G->d2u[ei];
        int64_t& at = fill[key_of(ei)];
        for (size_t j = s0; j < t; ++j) {
          if (j == i) continue;
          const int32_t ej = short_e[j];
          int32_t* row = &G->bond_graph[(size_t)at * 5];
          row[0] = ctr[ei]; row[1] = ui; row[2] = ei; row[3] = G->d2u[ej]; row[4] = ej;
          ++at;
        }
      }
      s0 = t;
    }
  }
  lap("bond graph");
  return CHG_OK;
}

extern "C" void chg_graph_sizes(const chg_graph* g, int64_t* n_edges, int64_t* n_bonds, int64_t* n_angles) {
  if (n_edges) *n_edges = g ? (int64_t)g->d2u.size() : 0;
  if (n_bonds) *n_bonds = g ? (int64_t)g->u2d.size() : 0;
  if (n_angles) *n_angles = g ? (int64_t)g->bond_graph.size() / 5 : 0;
}

extern "C" int chg_graph_export(const chg_graph* g, int32_t* atom_graph, float* image, int32_t* d2u, int32_t* u2d,
                                int32_t* bond_graph) {
  CHG_CHECK_ARG(g != nullptr, "null graph");
  auto put = [](void* dst, const void* src, size_t bytes) {
    if (dst != nullptr && bytes > 0) std::memcpy(dst, src, bytes);
  };
  put(atom_graph, g->atom_graph.data(), g->atom_graph.size() * 4);
  put(image, g->image.data(), g->image.size() * 4);
  put(d2u, g->d2u.data(), g->d2u.size() * 4);
  put(u2d, g->u2d.data(), g->u2d.size() * 4);
  put(bond_graph, g->bond_graph.data(), g->bond_graph.size() * 4);
  return CHG_OK;
}

extern "C" void chg_graph_free(chg_graph* g) { delete g; }

// ---- many structures at once (the converter loop of the reference's predict_structure, model.py:578-583) ----------
// chg_graph_build for n structures on the persistent worker threads (one structure per task, dynamic assignment).
// out[i] is always a handle to free (possibly of an empty graph); the return code is the first failure's, and its
// message is re-created on the calling thread.
extern "C" int chg_graph_build_many(int32_t n, const double* const* frac, const double* const* lattice, const int32_t* n_atoms,
                                    double r_atom, double r_bond, chg_graph** out) {
  CHG_CHECK_ARG(n >= 0, "negative size");
  CHG_CHECK_ARG(n == 0 || (frac != nullptr && lattice != nullptr && n_atoms != nullptr && out != nullptr), "null pointer");
  for (int i = 0; i < n; ++i) out[i] = nullptr;
  std::atomic<int> next{0};
  std::atomic<int> first_bad{INT32_MAX};
  WorkerPool& wp = pool();
  wp.run(std::min(wp.size(), std::max(1, n)), [&](int) {
    for (;;) {
      const int i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= n) break;
      if (chg_graph_build(frac[i], lattice[i], n_atoms[i], r_atom, r_bond, &out[i]) != CHG_OK) {
        int cur = first_bad.load();
        while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {
        }
      }
    }
  });
  const int bad = first_bad.load();
  if (bad == INT32_MAX) return CHG_OK;
  chg_graph* again = nullptr;  // the error text is per thread: repeat the failing build here
  const int rc = chg_graph_build(frac[bad], lattice[bad], n_atoms[bad], r_atom, r_bond, &again);
  delete again;
  return rc != CHG_OK ? rc : CHG_ERR_ARG;
}

// counts [n][3] = directed edges, bonds, angles; ptrs [n][5] = atom_graph, image, d2u, u2d, bond_graph (host pointers
// into the handles, valid until they are freed: what chg_pack_batch_wire / chg_pack_batch_host take); n_isolated [n].
extern "C" int chg_graph_views(int32_t n, chg_graph* const* graphs, int64_t* counts, uint64_t* ptrs, int32_t* n_isolated) {
  CHG_CHECK_ARG(n >= 0 && (n == 0 || (graphs != nullptr && counts != nullptr && ptrs != nullptr)), "bad size or null pointer");
  for (int i = 0; i < n; ++i) {
    const chg_graph* g = graphs[i];
    CHG_CHECK_ARG(g != nullptr, "null graph");
    counts[3 * i] = (int64_t)g->d2u.size();
    counts[3 * i + 1] = (int64_t)g->u2d.size();
    counts[3 * i + 2] = (int64_t)g->bond_graph.size() / 5;
    ptrs[5 * i] = (uint64_t)(uintptr_t)g->atom_graph.data();
    ptrs[5 * i + 1] = (uint64_t)(uintptr_t)g->image.data();
    ptrs[5 * i + 2] = (uint64_t)(uintptr_t)g->d2u.data();
    ptrs[5 * i + 3] = (uint64_t)(uintptr_t)g->u2d.data();
    ptrs[5 * i + 4] = (uint64_t)(uintptr_t)g->bond_graph.data();
    if (n_isolated != nullptr) n_isolated[i] = g->n_isolated;
  }
  return CHG_OK;
}

extern "C" void chg_graph_free_many(int32_t n, chg_graph** graphs) {
  for (int i = 0; i < n; ++i) {
    delete graphs[i];
    graphs[i] = nullptr;
  }
}

// =====================================================================================
// Host batch packer: list of CrystalGraphs -> the concatenated, offset-adjusted SoA of
// chgnet_b200/batch.py::DeviceBatch in ONE pass over host memory (what BatchedGraph.from_graphs does
// with ~25 tensor ops per graph, reference model.py:820-899).  Pure host code; the caller ships the
// two staging buffers to the device with one copy each.
//   ibuf layout (int32): z[N] owner[N] center[Ed] nbr[Ed] d2u[Ed] u2d[Eu] ang_atom[A] ang_i[A] ang_di[A]
//                        ang_j[A] ang_dj[A]
//   fbuf layout (fp32) : frac[N*3] image[Ed*3] lattice[B*9]
// flags_out[0] = edges sorted by centre within every graph, flags_out[1] = angles sorted by bond i.
// =====================================================================================
extern "C" int chg_pack_batch_host(int32_t n_graphs, const int64_t* counts /* [B][4]: atoms, edges, bonds, angles */,
                                   const void* const* ptrs /* [B][8]: z, frac, atom_graph, image, d2u, u2d, bond_graph, lattice */,
                                   int32_t* ibuf, float* fbuf, int32_t* flags_out) {
  CHG_CHECK_ARG(n_graphs >= 0, "negative size");
  CHG_CHECK_ARG(counts != nullptr && ptrs != nullptr && ibuf != nullptr && fbuf != nullptr && flags_out != nullptr, "null pointer");
  // per-graph offsets (exclusive prefix sums of the counts)
  std::vector<int64_t> off((size_t)(n_graphs + 1) * 4, 0);
  for (int g = 0; g < n_graphs; ++g)
    for (int k = 0; k < 4; ++k) off[(size_t)(g + 1) * 4 + k] = off[(size_t)g * 4 + k] + counts[4 * g + k];
  const int64_t N = off[(size_t)n_graphs * 4], Ed = off[(size_t)n_graphs * 4 + 1], Eu = off[(size_t)n_graphs * 4 + 2],
                A = off[(size_t)n_graphs * 4 + 3];
  CHG_CHECK_ARG(N < INT32_MAX && Ed < INT32_MAX && A < INT32_MAX, "batch too large for int32 indices");
  int32_t* z = ibuf;
  int32_t* owner = z + N;
  int32_t* center = owner + N;
  int32_t* nbr = center + Ed;
  int32_t* d2u = nbr + Ed;
  int32_t* u2d = d2u + Ed;
  int32_t* ang_atom = u2d + Eu;
  int32_t* ang_i = ang_atom + A;
  int32_t* ang_di = ang_i + A;
  int32_t* ang_j = ang_di + A;
  int32_t* ang_dj = ang_j + A;
  float* frac = fbuf;
  float* image = frac + N * 3;
  float* lattice = image + Ed * 3;
  static thread_local std::vector<uint8_t> in_bond_graph;  // one flag per bond of the batch (bonds never cross graphs)
  in_bond_graph.assign((size_t)Eu, 0);
  uint8_t* bg_flag = in_bond_graph.data();

  struct Partial {
    bool edges_sorted = true, angles_sorted = true;
    int64_t bad_z = -1, n_short = 0;
  };
  // [part, parts): this worker's share of every array of the graphs g0..g1-1 (parts == 1: whole graphs)
  auto pack_range = [&](int g0, int g1, int part, int parts, Partial& res) {
    for (int g = g0; g < g1; ++g) {
      const int64_t n = counts[4 * g], ed = counts[4 * g + 1], eu = counts[4 * g + 2], an = counts[4 * g + 3];
      const int64_t a_off = off[(size_t)g * 4], e_off = off[(size_t)g * 4 + 1], u_off = off[(size_t)g * 4 + 2], g_off = off[(size_t)g * 4 + 3];
      const void* const* p = ptrs + 8 * g;
      auto lo = [&](int64_t len) { return len * part / parts; };
      auto hi = [&](int64_t len) { return len * (part + 1) / parts; };
      if (n > 0) {
        const int64_t i0 = lo(n), i1 = hi(n);
        std::memcpy(z + a_off + i0, static_cast<const int32_t*>(p[0]) + i0, (size_t)(i1 - i0) * 4);
        std::memcpy(frac + (a_off + i0) * 3, static_cast<const float*>(p[1]) + i0 * 3, (size_t)(i1 - i0) * 12);
        for (int64_t i = i0; i < i1; ++i) {
          owner[a_off + i] = g;
          const int32_t zi = z[a_off + i];
          if ((zi < 1 || zi > CHG_MAX_Z) && res.bad_z < 0) res.bad_z = a_off + i;
        }
      }
      const int32_t* ag = static_cast<const int32_t*>(p[2]);
      const int32_t* du = static_cast<const int32_t*>(p[4]);
      const int64_t e0 = lo(ed), e1 = hi(ed);
      for (int64_t e = e0; e < e1; ++e) {
        center[e_off + e] = ag[2 * e] + (int32_t)a_off;
        nbr[e_off + e] = ag[2 * e + 1] + (int32_t)a_off;
        d2u[e_off + e] = du[e] + (int32_t)u_off;
        if (e > 0 && ag[2 * e] < ag[2 * e - 2]) res.edges_sorted = false;
      }
      if (e1 > e0) std::memcpy(image + (e_off + e0) * 3, static_cast<const float*>(p[3]) + e0 * 3, (size_t)(e1 - e0) * 12);
      const int32_t* ud = static_cast<const int32_t*>(p[5]);
      for (int64_t u = lo(eu); u < hi(eu); ++u) u2d[u_off + u] = ud[u] + (int32_t)e_off;
      const int32_t* bg = static_cast<const int32_t*>(p[6]);
      for (int64_t a = lo(an); a < hi(an); ++a) {
        ang_atom[g_off + a] = bg[5 * a] + (int32_t)a_off;
        ang_i[g_off + a] = bg[5 * a + 1] + (int32_t)u_off;
        ang_di[g_off + a] = bg[5 * a + 2] + (int32_t)e_off;
        ang_j[g_off + a] = bg[5 * a + 3] + (int32_t)u_off;
        ang_dj[g_off + a] = bg[5 * a + 4] + (int32_t)e_off;
        if (a > 0 && bg[5 * a + 1] < bg[5 * a - 4]) res.angles_sorted = false;
        for (int which = 1; which <= 3; which += 2) {  // bond i, bond j: this graph's own range of the flag array
          const int64_t ul = bg[5 * a + which];
          // test-and-set: workers that share a graph may meet the same bond
          if (ul >= 0 && ul < eu && __atomic_exchange_n(&bg_flag[u_off + ul], (uint8_t)1, __ATOMIC_RELAXED) == 0) ++res.n_short;
        }
      }
      if (part == 0) std::memcpy(lattice + (size_t)g * 9, p[7], 36);
    }
  };

  // many graphs: contiguous ranges of whole graphs of about equal size per worker; few (large) graphs: every worker takes
  // a slice of every array of every graph
  const int64_t total_items = N * 5 + Ed * 6 + Eu + A * 5;
  int n_thr = 1;
  if (total_items > (1 << 20)) {
    const unsigned hc = std::thread::hardware_concurrency();
    n_thr = (int)std::min<int64_t>(std::min<unsigned>(hc == 0 ? 4 : hc, 16), std::max<int64_t>(1, total_items >> 19));
    if (const char* e = std::getenv("CHG_PACK_THREADS")) n_thr = std::max(1, std::min(std::atoi(e), 64));
  }
  const bool by_graph = n_graphs >= 4 * n_thr;
  std::vector<Partial> parts((size_t)n_thr);
  if (n_thr <= 1) {
    pack_range(0, n_graphs, 0, 1, parts[0]);
  } else if (by_graph) {
    auto weight = [&](int g) { return off[(size_t)g * 4] * 5 + off[(size_t)g * 4 + 1] * 6 + off[(size_t)g * 4 + 2] + off[(size_t)g * 4 + 3] * 5; };
    std::vector<int> cut((size_t)n_thr + 1, n_graphs);
    cut[0] = 0;
    for (int t = 1, g = 0; t < n_thr; ++t) {
      const int64_t target = total_items * t / n_thr;
      while (g < n_graphs && weight(g) < target) ++g;
      cut[t] = g;
    }
    std::vector<std::thread> pool;
    pool.reserve((size_t)n_thr - 1);
    for (int t = 1; t < n_thr; ++t) pool.emplace_back(pack_range, cut[t], cut[t + 1], 0, 1, std::ref(parts[t]));
    pack_range(cut[0], cut[1], 0, 1, parts[0]);
    for (auto& th : pool) th.join();
  } else {
    std::vector<std::thread> pool;
    pool.reserve((size_t)n_thr - 1);
    for (int t = 1; t < n_thr; ++t) pool.emplace_back(pack_range, 0, n_graphs, t, n_thr, std::ref(parts[t]));
    pack_range(0, n_graphs, 0, n_thr, parts[0]);
    for (auto& th : pool) th.join();
  }
  bool edges_sorted = true, angles_sorted = true;
  int64_t bad_z = -1, n_short = 0;
  for (const Partial& r : parts) {
    edges_sorted = edges_sorted && r.edges_sorted;
    angles_sorted = angles_sorted && r.angles_sorted;
    if (r.bad_z >= 0 && (bad_z < 0 || r.bad_z < bad_z)) bad_z = r.bad_z;
    n_short += r.n_short;
  }
  flags_out[0] = edges_sorted ? 1 : 0;
  flags_out[1] = angles_sorted ? 1 : 0;
  flags_out[2] = (int32_t)bad_z;
  flags_out[3] = (int32_t)n_short;
  return CHG_OK;
}
