// Device CrystalGraph builder: periodic neighbour list, undirected-bond pairing and bond (line) graph as CUDA
// kernels that write the batch descriptor arrays (chg_batch: center / nbr / image / d2u / u2d / ang_*) directly.
//
// Replaces, for one structure, what the reference does on the host every MD step
// (chgnet/model/dynamics.py:156-157 -> converter.py:102-190: pymatgen's get_neighbor_list, then
// fast_converter_libraries/create_graph.c:100-219 / cygraph.pyx:69-175, then graph.py:283-327) and what
// csrc/graph_builder.cu does in host C++.  The integer outputs are BIT-IDENTICAL to the host builder's
// (tests/test_graph_device_gpu.py): the same fp64 distance arithmetic (round-to-nearest multiplies and adds, no
// FMA contraction), the same ordering rules:
//   * directed edges: every (center, neighbour, image) with 1e-8 < d <= r_atom, sorted by (center, neighbour, image);
//   * undirected bonds numbered by first appearance; u2d points at the first directed edge of the pair;
//   * bond-graph rows (atom, u_i, e_i, u_j, e_j) for e_i, e_j with d < r_bond leaving the same atom, e_j != e_i,
//     sorted by (u_i, e_i is the second edge of its bond, e_j).
//
// Search: atoms are binned by wrapped fractional coordinate into nb_k = max(1, floor(height_k / r_atom)) bins per axis
// (a bin is at least r_atom thick, or the whole cell); one warp per centre walks the (2R+1)^3 neighbouring bins
// (R_k = ceil(r_atom / bin thickness): > 1 only for cells thinner than the cutoff, where the same bin is visited once
// per periodic image).  Counting pass -> exclusive scan -> fill pass -> per-centre rank sort by the packed
// (neighbour, image) key.  Reverse edges by binary search in the neighbour's sorted segment.  Everything else is
// flags + exclusive scans.  Two host reads of one int32 each (edge count, angle count) size the outputs' used parts;
// the buffers themselves are caller-allocated with a capacity.
#include <cmath>
#include <cstdint>
#include <cstdio>

#include "common.cuh"

namespace chg {

int exclusive_scan_i32(const int32_t* in, int n, int32_t* out, int32_t* sums, cudaStream_t st);  // batch_csr.cu

namespace {

struct Cell {
  double L[9];       // rows = lattice vectors
  double r_atom, r_bond, r2_max, tol;
  int nb[3], R[3];   // bins per axis, search radius in bins
};

__device__ __forceinline__ double dot3_rn(double a0, double b0, double a1, double b1, double a2, double b2) {
  // (a0*b0 + a1*b1) + a2*b2 with separately rounded products and sums: what the host compiler emits (no FMA)
  return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}

__device__ __forceinline__ uint64_t pack_key(int atom, int i0, int i1, int i2) {
  return ((uint64_t)atom << 24) | ((uint64_t)(i0 + 128) << 16) | ((uint64_t)(i1 + 128) << 8) | (uint64_t)(i2 + 128);
}

// cartesian positions, the integer part of the fractional coordinates, and the bin of the wrapped position
__global__ void prepare_atoms_kernel(const double* __restrict__ frac, int n, Cell cell, double* __restrict__ cart,
                                     int32_t* __restrict__ shift, int32_t* __restrict__ bin, int32_t* __restrict__ bin_cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double f0 = frac[3 * i], f1 = frac[3 * i + 1], f2 = frac[3 * i + 2];
#pragma unroll
  for (int j = 0; j < 3; ++j) cart[3 * i + j] = dot3_rn(f0, cell.L[j], f1, cell.L[3 + j], f2, cell.L[6 + j]);
  const double f[3] = {f0, f1, f2};
  int b[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double fl = floor(f[k]);
    shift[3 * i + k] = (int)fl;
    int bk = (int)((f[k] - fl) * cell.nb[k]);
    b[k] = min(max(bk, 0), cell.nb[k] - 1);
  }
  const int id = (b[0] * cell.nb[1] + b[1]) * cell.nb[2] + b[2];
  bin[i] = id;
  atomicAdd(bin_cnt + id, 1);
}

__global__ void bin_fill_kernel(const int32_t* __restrict__ bin, int n, const int32_t* __restrict__ bin_ptr,
                                int32_t* __restrict__ cursor, int32_t* __restrict__ binned) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = bin[i];
  binned[bin_ptr[b] + atomicAdd(cursor + b, 1)] = i;
}

// One warp per centre.  FILL = false: count the neighbours; FILL = true: write (key, distance) at ptr_c[c] + k (any order).
template <bool FILL>
__global__ void neighbour_kernel(const double* __restrict__ cart, const int32_t* __restrict__ shift, const int32_t* __restrict__ bin,
                                 const int32_t* __restrict__ bin_ptr, const int32_t* __restrict__ binned, int n, Cell cell,
                                 int32_t* __restrict__ cnt, const int32_t* __restrict__ ptr_c, uint64_t* __restrict__ keys,
                                 double* __restrict__ dist, int cap_edges, int32_t* __restrict__ err) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n) return;
  const double pc0 = cart[3 * c], pc1 = cart[3 * c + 1], pc2 = cart[3 * c + 2];
  const int sc0 = shift[3 * c], sc1 = shift[3 * c + 1], sc2 = shift[3 * c + 2];
  const int bid = bin[c];
  const int b2 = bid % cell.nb[2], b1 = (bid / cell.nb[2]) % cell.nb[1], b0 = bid / (cell.nb[1] * cell.nb[2]);
  int found = 0;
  const int base = FILL ? ptr_c[c] : 0;
  for (int d0 = -cell.R[0]; d0 <= cell.R[0]; ++d0) {
    const int t0 = b0 + d0;
    const int m0 = (t0 >= 0 ? t0 / cell.nb[0] : -((-t0 + cell.nb[0] - 1) / cell.nb[0]));  // floor division
    const int w0 = t0 - m0 * cell.nb[0];
    for (int d1 = -cell.R[1]; d1 <= cell.R[1]; ++d1) {
      const int t1 = b1 + d1;
      const int m1 = (t1 >= 0 ? t1 / cell.nb[1] : -((-t1 + cell.nb[1] - 1) / cell.nb[1]));
      const int w1 = t1 - m1 * cell.nb[1];
      for (int d2 = -cell.R[2]; d2 <= cell.R[2]; ++d2) {
        const int t2 = b2 + d2;
        const int m2 = (t2 >= 0 ? t2 / cell.nb[2] : -((-t2 + cell.nb[2] - 1) / cell.nb[2]));
        const int w2 = t2 - m2 * cell.nb[2];
        const int nbin = (w0 * cell.nb[1] + w1) * cell.nb[2] + w2;
        const int beg = bin_ptr[nbin], end = bin_ptr[nbin + 1];
        for (int q0 = beg; q0 < end; q0 += 32) {
          const int q = q0 + lane;
          bool hit = false;
          uint64_t key = 0;
          double d = 0.0;
          if (q < end) {
            const int at = binned[q];
            // image of the neighbour relative to its ORIGINAL fractional coordinates
            const int i0 = m0 - shift[3 * at] + sc0, i1 = m1 - shift[3 * at + 1] + sc1, i2 = m2 - shift[3 * at + 2] + sc2;
            const double x0 = (double)i0, x1 = (double)i1, x2 = (double)i2;
            const double p0 = __dadd_rn(cart[3 * at], dot3_rn(x0, cell.L[0], x1, cell.L[3], x2, cell.L[6]));
            const double p1 = __dadd_rn(cart[3 * at + 1], dot3_rn(x0, cell.L[1], x1, cell.L[4], x2, cell.L[7]));
            const double p2 = __dadd_rn(cart[3 * at + 2], dot3_rn(x0, cell.L[2], x1, cell.L[5], x2, cell.L[8]));
            const double dx = __dsub_rn(p0, pc0), dy = __dsub_rn(p1, pc1), dz = __dsub_rn(p2, pc2);
            const double dd = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (!(dd > cell.r2_max)) {
              d = __dsqrt_rn(dd);
              hit = d > cell.tol && d <= cell.r_atom;
              if (hit && (i0 < -127 || i0 > 127 || i1 < -127 || i1 > 127 || i2 < -127 || i2 > 127)) {
                atomicExch(err, 2);  // more than 127 periodic images along one axis
                hit = false;
              }
              key = pack_key(at, i0, i1, i2);
            }
          }
          const unsigned m = __ballot_sync(0xffffffffu, hit);
          if (FILL && hit) {
            const int pos = base + found + __popc(m & ((1u << lane) - 1u));
            if (pos < cap_edges) {
              keys[pos] = key;
              dist[pos] = d;
            }
          }
          found += __popc(m);
        }
      }
    }
  }
  if (!FILL && lane == 0) cnt[c] = found;
}

// ascending key order inside every centre's segment (rank sort, one warp per centre), plus the per-edge outputs
__global__ void sort_emit_kernel(const int32_t* __restrict__ ptr_c, int n, const uint64_t* __restrict__ keys_in,
                                 const double* __restrict__ dist_in, uint64_t* __restrict__ keys, double* __restrict__ dist,
                                 int32_t* __restrict__ center, int32_t* __restrict__ nbr, float* __restrict__ image) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n) return;
  const int beg = ptr_c[c], end = ptr_c[c + 1];
  for (int i = beg + lane; i < end; i += 32) {
    const uint64_t k = keys_in[i];
    int rank = 0;
    for (int j = beg; j < end; ++j) rank += keys_in[j] < k ? 1 : 0;
    const int e = beg + rank;
    keys[e] = k;
    dist[e] = dist_in[i];
    center[e] = c;
    nbr[e] = (int32_t)(k >> 24);
    image[3 * e] = (float)((int)((k >> 16) & 255) - 128);
    image[3 * e + 1] = (float)((int)((k >> 8) & 255) - 128);
    image[3 * e + 2] = (float)((int)(k & 255) - 128);
  }
}

// reverse edge of (c, n, img) = (n, c, -img): binary search in n's sorted segment; first[e] = this is the first of the pair
__global__ void pair_kernel(const int32_t* __restrict__ ptr_c, const uint64_t* __restrict__ keys, const int32_t* __restrict__ center,
                            int n_edges, int32_t* __restrict__ rev, int32_t* __restrict__ first, int32_t* __restrict__ err) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const uint64_t k = keys[e];
  const int nb = (int)(k >> 24);
  // image fields are stored as i + 128: the reverse image -i is 256 - field
  const uint64_t target = ((uint64_t)center[e] << 24) | ((uint64_t)(256 - (int)((k >> 16) & 255)) << 16) |
                          ((uint64_t)(256 - (int)((k >> 8) & 255)) << 8) | (uint64_t)(256 - (int)(k & 255));
  int lo = ptr_c[nb], hi = ptr_c[nb + 1];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  if (lo >= ptr_c[nb + 1] || keys[lo] != target) {
    atomicExch(err, 1);  // directed edges are not complete
    rev[e] = e;
    first[e] = 0;
    return;
  }
  rev[e] = lo;
  first[e] = lo > e ? 1 : 0;
}

__global__ void bond_ids_kernel(const int32_t* __restrict__ rev, const int32_t* __restrict__ first,
                                const int32_t* __restrict__ first_scan, int n_edges, int32_t* __restrict__ d2u,
                                int32_t* __restrict__ u2d) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  if (first[e]) {
    const int u = first_scan[e];
    d2u[e] = u;
    u2d[u] = e;
  } else {
    d2u[e] = first_scan[rev[e]];
  }
}

// per directed edge e (as bond-graph edge i): number of rows = (short edges of its centre) - 1 if e is short, keyed by
// key2 = 2 * d2u[e] + (e is the second edge of its bond)
__global__ void angle_count_kernel(const int32_t* __restrict__ ptr_c, const double* __restrict__ dist, const int32_t* __restrict__ d2u,
                                   const int32_t* __restrict__ u2d, int n, double r_bond, int32_t* __restrict__ cnt2,
                                   int32_t* __restrict__ n_short_at) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n) return;
  const int beg = ptr_c[c], end = ptr_c[c + 1];
  int ns = 0;
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    ns += __popc(__ballot_sync(0xffffffffu, e < end && dist[e] < r_bond));
  }
  if (lane == 0) n_short_at[c] = ns;
  for (int e = beg + lane; e < end; e += 32) {
    const int u = d2u[e];
    cnt2[2 * u + (u2d[u] != e ? 1 : 0)] = dist[e] < r_bond ? ns - 1 : 0;
  }
}

// one warp per directed short edge e_i: rows (atom, u_i, e_i, u_j, e_j) for the other short edges e_j of the centre, ascending
__global__ void angle_fill_kernel(const int32_t* __restrict__ ptr_c, const int32_t* __restrict__ center, const double* __restrict__ dist,
                                  const int32_t* __restrict__ d2u, const int32_t* __restrict__ u2d, const int32_t* __restrict__ row_ptr,
                                  int n_edges, double r_bond, int cap_angles, int32_t* __restrict__ ang_atom,
                                  int32_t* __restrict__ ang_i, int32_t* __restrict__ ang_di, int32_t* __restrict__ ang_j,
                                  int32_t* __restrict__ ang_dj) {
  const int ei = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (ei >= n_edges || !(dist[ei] < r_bond)) return;
  const int c = center[ei], ui = d2u[ei];
  int at = row_ptr[2 * ui + (u2d[ui] != ei ? 1 : 0)];
  const int beg = ptr_c[c], end = ptr_c[c + 1];
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int ej = e0 + lane;
    const bool ok = ej < end && ej != ei && dist[ej] < r_bond;
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const int row = at + __popc(m & ((1u << lane) - 1u));
      if (row < cap_angles) {
        ang_atom[row] = c;
        ang_i[row] = ui;
        ang_di[row] = ei;
        ang_j[row] = d2u[ej];
        ang_dj[row] = ej;
      }
    }
    at += __popc(m);
  }
}

inline unsigned blocks(long long n, int per = 256) { return (unsigned)((n + per - 1) / per); }

}  // namespace
}  // namespace chg

using namespace chg;

extern "C" int64_t chg_graph_device_scratch_bytes(int32_t n_atoms, int32_t cap_edges) {
  if (n_atoms < 0 || cap_edges < 0) return -1;
  const int64_t n = n_atoms, e = cap_edges;
  // cart 24n, shift 12n, bin 4n, binned 4n, bin_cnt / bin_ptr / cursor 3 x 4(n+2) (bins <= atoms + 1 is enforced),
  // cnt 4(n+1), n_short_at 4n, keys 2 x 8e, dist 2 x 8e, rev 4e, first 4e, first_scan 4(e+1), cnt2 4(e+1), row_ptr 4(e+2),
  // scan tile sums 16 KB, error flag + sizes
  return 24 * n + 12 * n + 4 * n + 4 * n + 12 * (n + 2) + 4 * (n + 1) + 4 * n + 32 * e + 8 * e + 4 * (e + 1) + 4 * (e + 1) + 4 * (e + 2) +
         16384 + 4096;
}

// frac [N][3] fp64 (device), lattice [9] fp64 (host, rows = lattice vectors).  Outputs (device, caller-allocated with the
// given capacities): center / nbr / d2u [cap_edges], image [cap_edges][3] fp32, u2d [cap_edges / 2], ptr_c [N + 1],
// ang_atom / ang_i / ang_di / ang_j / ang_dj [cap_angles].  sizes_out (host) = {n_edges, n_bonds, n_angles, 0}.
// Returns CHG_ERR_ARG with "capacity" in the message when a capacity is too small (sizes_out then holds the needed sizes).
extern "C" int chg_graph_build_device(const double* frac, const double* lattice, int32_t n_atoms, double r_atom, double r_bond,
                                      int32_t cap_edges, int32_t cap_angles, int32_t* center, int32_t* nbr, float* image,
                                      int32_t* d2u, int32_t* u2d, int32_t* ptr_c, int32_t* ang_atom, int32_t* ang_i,
                                      int32_t* ang_di, int32_t* ang_j, int32_t* ang_dj, void* scratch, int32_t* sizes_out,
                                      void* stream) {
  CHG_CHECK_ARG(n_atoms >= 0 && r_atom > 0 && r_bond >= 0 && cap_edges >= 0 && cap_angles >= 0, "bad size or cutoff");
  CHG_CHECK_ARG(lattice != nullptr && sizes_out != nullptr && ptr_c != nullptr && scratch != nullptr, "null pointer");
  sizes_out[0] = sizes_out[1] = sizes_out[2] = sizes_out[3] = 0;
  cudaStream_t st = as_stream(stream);
  const int n = n_atoms;
  if (n == 0) {
    CHG_CUDA(cudaMemsetAsync(ptr_c, 0, 4, st));
    return CHG_OK;
  }
  CHG_CHECK_ARG(frac != nullptr, "null pointer");
  Cell cell;
  const double* L = lattice;
  for (int i = 0; i < 9; ++i) cell.L[i] = L[i];
  cell.r_atom = r_atom;
  cell.r_bond = r_bond;
  cell.r2_max = r_atom * r_atom * (1 + 1e-12);
  cell.tol = 1e-8;
  const double c12[3] = {L[4] * L[8] - L[5] * L[7], L[5] * L[6] - L[3] * L[8], L[3] * L[7] - L[4] * L[6]};
  const double c20[3] = {L[7] * L[2] - L[8] * L[1], L[8] * L[0] - L[6] * L[2], L[6] * L[1] - L[7] * L[0]};
  const double c01[3] = {L[1] * L[5] - L[2] * L[4], L[2] * L[3] - L[0] * L[5], L[0] * L[4] - L[1] * L[3]};
  const double vol = std::fabs(L[0] * c12[0] + L[1] * c12[1] + L[2] * c12[2]);
  CHG_CHECK_ARG(vol > 0, "singular lattice");
  const double* crosses[3] = {c12, c20, c01};
  long long n_bins = 1;
  for (int k = 0; k < 3; ++k) {
    const double* c = crosses[k];
    const double height = vol / std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    int nb = (int)std::floor(height / r_atom);
    nb = nb < 1 ? 1 : nb;
    cell.nb[k] = nb;
    n_bins *= nb;
  }
  // no more bins than atoms + 1 (scratch is sized by the atom count): coarsen the finest axis until it fits
  while (n_bins > (long long)n + 1) {
    int k = 0;
    for (int j = 1; j < 3; ++j)
      if (cell.nb[j] > cell.nb[k]) k = j;
    n_bins = n_bins / cell.nb[k];
    cell.nb[k] = (cell.nb[k] + 1) / 2;
    n_bins *= cell.nb[k];
  }
  for (int k = 0; k < 3; ++k) {
    const double* c = crosses[k];
    const double height = vol / std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    cell.R[k] = (int)std::ceil(r_atom / (height / cell.nb[k]) * (1 + 1e-12));
    CHG_CHECK_ARG(cell.R[k] <= 127, "more than 127 periodic images along one axis");
  }
  // ---- scratch carving ----
  char* p = static_cast<char*>(scratch);
  auto take = [&](size_t bytes) {
    char* q = p;
    p += (bytes + 255) / 256 * 256;
    return q;
  };
  double* cart = reinterpret_cast<double*>(take((size_t)n * 24));
  uint64_t* keys_a = reinterpret_cast<uint64_t*>(take((size_t)cap_edges * 8));
  uint64_t* keys = reinterpret_cast<uint64_t*>(take((size_t)cap_edges * 8));
  double* dist_a = reinterpret_cast<double*>(take((size_t)cap_edges * 8));
  double* dist = reinterpret_cast<double*>(take((size_t)cap_edges * 8));
  int32_t* shift = reinterpret_cast<int32_t*>(take((size_t)n * 12));
  int32_t* bin = reinterpret_cast<int32_t*>(take((size_t)n * 4));
  int32_t* binned = reinterpret_cast<int32_t*>(take((size_t)n * 4));
  int32_t* bin_cnt = reinterpret_cast<int32_t*>(take((size_t)(n + 2) * 4));
  int32_t* bin_ptr = reinterpret_cast<int32_t*>(take((size_t)(n + 2) * 4));
  int32_t* cursor = reinterpret_cast<int32_t*>(take((size_t)(n + 2) * 4));
  int32_t* cnt = reinterpret_cast<int32_t*>(take((size_t)(n + 1) * 4));
  int32_t* n_short_at = reinterpret_cast<int32_t*>(take((size_t)n * 4));
  int32_t* rev = reinterpret_cast<int32_t*>(take((size_t)cap_edges * 4));
  int32_t* first = reinterpret_cast<int32_t*>(take((size_t)cap_edges * 4));
  int32_t* first_scan = reinterpret_cast<int32_t*>(take((size_t)(cap_edges + 1) * 4));
  int32_t* cnt2 = reinterpret_cast<int32_t*>(take((size_t)(cap_edges + 1) * 4));
  int32_t* row_ptr = reinterpret_cast<int32_t*>(take((size_t)(cap_edges + 2) * 4));
  int32_t* sums = reinterpret_cast<int32_t*>(take(16384));
  int32_t* err = reinterpret_cast<int32_t*>(take(256));
  int rc;

  // ---- bins ----
  CHG_CUDA(cudaMemsetAsync(bin_cnt, 0, (size_t)(n_bins + 1) * 4, st));
  CHG_CUDA(cudaMemsetAsync(cursor, 0, (size_t)(n_bins + 1) * 4, st));
  CHG_CUDA(cudaMemsetAsync(err, 0, 4, st));
  prepare_atoms_kernel<<<blocks(n), 256, 0, st>>>(frac, n, cell, cart, shift, bin, bin_cnt);
  count_launch();
  if ((rc = exclusive_scan_i32(bin_cnt, (int)n_bins, bin_ptr, sums, st)) != CHG_OK) return rc;
  bin_fill_kernel<<<blocks(n), 256, 0, st>>>(bin, n, bin_ptr, cursor, binned);
  count_launch();
  // ---- neighbour list: count, scan, fill, sort ----
  neighbour_kernel<false><<<blocks((long long)n * 32), 256, 0, st>>>(cart, shift, bin, bin_ptr, binned, n, cell, cnt, nullptr, nullptr,
                                                                      nullptr, 0, err);
  count_launch();
  if ((rc = exclusive_scan_i32(cnt, n, ptr_c, sums, st)) != CHG_OK) return rc;
  int32_t n_edges = 0, h_err = 0;
  CHG_CUDA(cudaMemcpyAsync(&n_edges, ptr_c + n, 4, cudaMemcpyDeviceToHost, st));
  CHG_CUDA(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, st));
  CHG_CUDA(cudaStreamSynchronize(st));
  CHG_CHECK_ARG(h_err != 2, "more than 127 periodic images along one axis");
  sizes_out[0] = n_edges;
  sizes_out[1] = n_edges / 2;
  if (n_edges > cap_edges) {
    set_error("chg_graph_build_device: edge capacity %d too small (%d directed edges)", cap_edges, n_edges);
    return CHG_ERR_ARG;
  }
  if (n_edges == 0) return CHG_OK;
  CHG_CHECK_ARG(center && nbr && image && d2u && u2d, "null pointer");
  neighbour_kernel<true><<<blocks((long long)n * 32), 256, 0, st>>>(cart, shift, bin, bin_ptr, binned, n, cell, nullptr, ptr_c, keys_a,
                                                                     dist_a, cap_edges, err);
  count_launch();
  sort_emit_kernel<<<blocks((long long)n * 32), 256, 0, st>>>(ptr_c, n, keys_a, dist_a, keys, dist, center, nbr, image);
  count_launch();
  // ---- undirected bonds ----
  pair_kernel<<<blocks(n_edges), 256, 0, st>>>(ptr_c, keys, center, n_edges, rev, first, err);
  count_launch();
  if ((rc = exclusive_scan_i32(first, n_edges, first_scan, sums, st)) != CHG_OK) return rc;
  bond_ids_kernel<<<blocks(n_edges), 256, 0, st>>>(rev, first, first_scan, n_edges, d2u, u2d);
  count_launch();
  // ---- bond graph ----
  angle_count_kernel<<<blocks((long long)n * 32), 256, 0, st>>>(ptr_c, dist, d2u, u2d, n, r_bond, cnt2, n_short_at);
  count_launch();
  if ((rc = exclusive_scan_i32(cnt2, n_edges, row_ptr, sums, st)) != CHG_OK) return rc;
  int32_t n_angles = 0;
  CHG_CUDA(cudaMemcpyAsync(&n_angles, row_ptr + n_edges, 4, cudaMemcpyDeviceToHost, st));
  CHG_CUDA(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, st));
  CHG_CUDA(cudaStreamSynchronize(st));
  if (h_err == 1) {
    set_error("chg_graph_build_device: directed edges are not complete: some undirected bond does not have exactly 2 directed edges");
    return CHG_ERR_ARG;
  }
  sizes_out[2] = n_angles;
  if (n_angles > cap_angles) {
    set_error("chg_graph_build_device: angle capacity %d too small (%d angles)", cap_angles, n_angles);
    return CHG_ERR_ARG;
  }
  if (n_angles > 0) {
    CHG_CHECK_ARG(ang_atom && ang_i && ang_di && ang_j && ang_dj, "null pointer");
    angle_fill_kernel<<<blocks((long long)n_edges * 32), 256, 0, st>>>(ptr_c, center, dist, d2u, u2d, row_ptr, n_edges, r_bond, cap_angles,
                                                                       ang_atom, ang_i, ang_di, ang_j, ang_dj);
    count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("chg_graph_build_device: launch failed: %s", cudaGetErrorString(e));
    return CHG_ERR_CUDA;
  }
  return CHG_OK;
}
