// Segment structures of a batch, built on the device:  chg_build_csr
//
// The batch arrives with its directed edges sorted by centre and its angles sorted by bond i (reference
// graph.py:132-328 emits them that way; batch.py reorders otherwise), so the forward CSR pointers are boundary
// searches, and the transposed groupings the reverse pass gathers through (edges by neighbour / by bond, angles
// by bond j / by atom) are counting sorts: histogram -> exclusive scan -> fill -> ascending order inside every
// segment (a rank sort per segment, so the result equals a STABLE sort by key and every later sum has a fixed
// order).  This replaces ~30 torch ops (4 ATen radix sorts, searchsorted, nonzero + its host sync) of round 1's
// batch.py; integer work, a dozen small launches, no host synchronisation (the number of bond-graph bonds is
// counted by the host packer).
//
// Also here: the compact index space of the bonds that appear in the bond graph (short_ids, ang_is / ang_js,
// ptr_is, ptr_js; perm_js == perm_j because slot numbers are monotone in the bond index).
#include "common.cuh"

namespace chg {
namespace {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// ptr[v] = first position p with keys[p] >= v, for sorted keys (v = 0 .. n_rows)
__global__ void csr_ptr_sorted_kernel(const int32_t* __restrict__ keys, int n, int n_rows, int32_t* __restrict__ ptr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  const int lo = p == 0 ? 0 : min(keys[p - 1] + 1, n_rows + 1);  // first row whose pointer is p
  const int hi = p == n ? n_rows + 1 : min(keys[p] + 1, n_rows + 1);
  for (int v = lo; v < hi; ++v) ptr[v] = p;
}

__global__ void histogram_kernel(const int32_t* __restrict__ keys, int n, int32_t* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(cnt + keys[i], 1);
}

// bond-graph membership: mask[u] = 1 if bond u is bond i or bond j of some angle
__global__ void mark_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int n, int32_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mask[a[i]] = 1;
    mask[b[i]] = 1;
  }
}

// ---- exclusive scan of int32 (three phases; n <= SCAN_TILE * SCAN_TILE) -------------------------------------
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) s_warp[32] = wi;
  }
  __syncthreads();
  total = s_warp[32];
  const int res = s_warp[warp] + incl - v;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const int32_t* __restrict__ in, int n, int32_t* __restrict__ sums) {
  __shared__ int s_warp[33];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) v += in[base + k];
  int total;
  block_exclusive_scan(v, s_warp, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums in place (n_tiles <= SCAN_TILE)
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(int32_t* __restrict__ sums, int n_tiles) {
  __shared__ int s_warp[33];
  const int base = threadIdx.x * SCAN_ITEMS;
  int x[SCAN_ITEMS], v = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    x[k] = base + k < n_tiles ? sums[base + k] : 0;
    v += x[k];
  }
  int total;
  int off = block_exclusive_scan(v, s_warp, total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n_tiles) sums[base + k] = off;
    off += x[k];
  }
}

// out[i] = exclusive prefix of in (out has n + 1 entries: out[n] = total); in == out allowed
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int32_t* in, int n, const int32_t* __restrict__ sums,
                                                                  int32_t* out) {
  __shared__ int s_warp[33];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int x[SCAN_ITEMS], v = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    x[k] = base + k < n ? in[base + k] : 0;
    v += x[k];
  }
  int total;
  int off = sums[blockIdx.x] + block_exclusive_scan(v, s_warp, total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = off;
    off += x[k];
    if (base + k == n - 1) out[n] = off;
  }
}

int exclusive_scan(const int32_t* in, int n, int32_t* out, int32_t* sums, cudaStream_t st) {
  if (n == 0) {
    CHG_CUDA(cudaMemsetAsync(out, 0, 4, st));
    return CHG_OK;
  }
  const int n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (n_tiles > SCAN_TILE) {
    set_error("chg_build_csr: more than %d keys in one scan", SCAN_TILE * SCAN_TILE);
    return CHG_ERR_ARG;
  }
  scan_tile_sums_kernel<<<n_tiles, SCAN_THREADS, 0, st>>>(in, n, sums);
  count_launch();
  scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(sums, n_tiles);
  count_launch();
  scan_apply_kernel<<<n_tiles, SCAN_THREADS, 0, st>>>(in, n, sums, out);
  count_launch();
  return CHG_OK;
}

// items into their key's segment, arbitrary order inside a segment (made canonical by segment_sort_kernel)
__global__ void fill_kernel(const int32_t* __restrict__ keys, int n, const int32_t* __restrict__ ptr, int32_t* __restrict__ cursor,
                            int32_t* __restrict__ perm_tmp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = keys[i];
  perm_tmp[ptr[k] + atomicAdd(cursor + k, 1)] = i;
}

// ascending order inside every segment: one warp per segment, rank sort (values are distinct item indices)
__global__ void segment_sort_kernel(const int32_t* __restrict__ ptr, int n_seg, const int32_t* __restrict__ tmp,
                                    int32_t* __restrict__ perm) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_seg) return;
  const int beg = ptr[warp], end = ptr[warp + 1];
  const int len = end - beg;
  if (len <= 32) {
    const int v = lane < len ? tmp[beg + lane] : INT32_MAX;
    int rank = 0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) rank += __shfl_sync(0xffffffffu, v, j) < v ? 1 : 0;
    if (lane < len) perm[beg + rank] = v;
    return;
  }
  for (int i = lane; i < len; i += 32) {
    const int v = tmp[beg + i];
    int rank = 0;
    for (int j = 0; j < len; ++j) rank += tmp[beg + j] < v ? 1 : 0;
    perm[beg + rank] = v;
  }
}

// the two directed edges of every bond, ascending
__global__ void pair_sort_kernel(int32_t* __restrict__ perm_u, int n_bonds) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_bonds) return;
  const int a = perm_u[2 * u], b = perm_u[2 * u + 1];
  if (a > b) {
    perm_u[2 * u] = b;
    perm_u[2 * u + 1] = a;
  }
}

__global__ void iota2_kernel(int32_t* __restrict__ ptr_u, int n_bonds) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u <= n_bonds) ptr_u[u] = 2 * u;
}

// slot[u] = exclusive scan of mask (already in `slot`); compact ids and the slot-space copies of the angle indices
__global__ void compact_ids_kernel(const int32_t* __restrict__ mask, const int32_t* __restrict__ slot, int n_bonds,
                                   int32_t* __restrict__ short_ids) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n_bonds && mask[u]) short_ids[slot[u]] = u;
}
__global__ void remap_kernel(const int32_t* __restrict__ slot, const int32_t* __restrict__ a_in, const int32_t* __restrict__ b_in,
                             int n, int32_t* __restrict__ a_out, int32_t* __restrict__ b_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    a_out[i] = slot[a_in[i]];
    b_out[i] = slot[b_in[i]];
  }
}
// ptr_js[s] = ptr_j[short_ids[s]] (s < n_short), ptr_js[n_short] = n_angles
__global__ void gather_ptr_kernel(const int32_t* __restrict__ ptr_j, const int32_t* __restrict__ short_ids, int n_short,
                                  int n_angles, int32_t* __restrict__ ptr_js) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_short) ptr_js[s] = ptr_j[short_ids[s]];
  if (s == n_short) ptr_js[s] = n_angles;
}

inline unsigned blocks(int n, int per = 256) { return (unsigned)((n + per - 1) / per); }

}  // namespace

// out[0..n] = exclusive prefix sums of in[0..n) (out[n] = total); sums: >= 4096 int32 of scratch.  Shared with
// graph_device.cu.
int exclusive_scan_i32(const int32_t* in, int n, int32_t* out, int32_t* sums, cudaStream_t st) {
  return exclusive_scan(in, n, out, sums, st);
}
}  // namespace chg

using namespace chg;

extern "C" int64_t chg_build_csr_scratch_ints(int32_t n_atoms, int32_t n_edges, int32_t n_bonds, int32_t n_angles) {
  if (n_atoms < 0 || n_edges < 0 || n_bonds < 0 || n_angles < 0) return -1;
  const int64_t keys = (int64_t)(n_atoms > n_bonds ? n_atoms : n_bonds) + 1;
  const int64_t items = n_edges > n_angles ? n_edges : n_angles;
  return 2 * keys + items + 4096 + 64;  // counters / cursors, unsorted permutation, scan tile sums
}

extern "C" int chg_build_csr(const chg_csr_in* in, const chg_csr_out* out, int32_t* scratch, void* stream) {
  CHG_CHECK_ARG(in != nullptr && out != nullptr, "null pointer");
  const int N = in->n_atoms, Ed = in->n_edges, Eu = in->n_bonds, A = in->n_angles, Es = in->n_short;
  CHG_CHECK_ARG(N >= 0 && Ed >= 0 && Eu >= 0 && A >= 0 && Ed == 2 * Eu, "bad sizes (directed edges must be 2 x bonds)");
  CHG_CHECK_ARG(scratch != nullptr && out->ptr_c && out->ptr_i, "null pointer");
  CHG_CHECK_ARG(Ed == 0 || (in->center && in->nbr && in->d2u), "null pointer");
  CHG_CHECK_ARG(A == 0 || (in->ang_atom && in->ang_i && in->ang_j), "null pointer");
  cudaStream_t st = as_stream(stream);
  const int64_t keys = (int64_t)(N > Eu ? N : Eu) + 1;
  int32_t* cnt = scratch;             // [keys]
  int32_t* cursor = cnt + keys;       // [keys]
  int32_t* tmp = cursor + keys;       // [max(Ed, A)]
  int32_t* sums = tmp + (Ed > A ? Ed : A);  // [4096]
  int rc = CHG_OK;

  // forward pointers: inputs are sorted by centre / by bond i
  csr_ptr_sorted_kernel<<<blocks(Ed + 1), 256, 0, st>>>(in->center, Ed, N, out->ptr_c);
  count_launch();
  csr_ptr_sorted_kernel<<<blocks(A + 1), 256, 0, st>>>(in->ang_i, A, Eu, out->ptr_i);
  count_launch();

  // one transposed grouping: histogram -> scan -> fill -> canonical order
  auto group = [&](const int32_t* keyv, int n_items, int n_rows, int32_t* ptr, int32_t* perm) -> int {
    CHG_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(n_rows + 1) * 4, st));
    CHG_CUDA(cudaMemsetAsync(cursor, 0, (size_t)(n_rows + 1) * 4, st));
    if (n_items > 0) {
      histogram_kernel<<<blocks(n_items), 256, 0, st>>>(keyv, n_items, cnt);
      count_launch();
    }
    const int r = exclusive_scan(cnt, n_rows, ptr, sums, st);
    if (r != CHG_OK) return r;
    if (n_items > 0) {
      fill_kernel<<<blocks(n_items), 256, 0, st>>>(keyv, n_items, ptr, cursor, tmp);
      count_launch();
      segment_sort_kernel<<<blocks(n_rows * 32), 256, 0, st>>>(ptr, n_rows, tmp, perm);
      count_launch();
    }
    return CHG_OK;
  };

  if (in->with_reverse) {
    CHG_CHECK_ARG(out->perm_n && out->ptr_n && out->perm_u && out->ptr_u && out->ptr_j && out->ptr_x, "null pointer");
    CHG_CHECK_ARG(A == 0 || (out->perm_j && out->perm_x), "null pointer");
    if ((rc = group(in->nbr, Ed, N, out->ptr_n, out->perm_n)) != CHG_OK) return rc;
    // the two directed edges of every bond (ptr_u = 2 u)
    iota2_kernel<<<blocks(Eu + 1), 256, 0, st>>>(out->ptr_u, Eu);
    count_launch();
    if (Ed > 0) {
      CHG_CUDA(cudaMemsetAsync(cursor, 0, (size_t)(Eu + 1) * 4, st));
      fill_kernel<<<blocks(Ed), 256, 0, st>>>(in->d2u, Ed, out->ptr_u, cursor, out->perm_u);
      count_launch();
      pair_sort_kernel<<<blocks(Eu), 256, 0, st>>>(out->perm_u, Eu);
      count_launch();
    }
    if ((rc = group(in->ang_j, A, Eu, out->ptr_j, out->perm_j)) != CHG_OK) return rc;
    if ((rc = group(in->ang_atom, A, N, out->ptr_x, out->perm_x)) != CHG_OK) return rc;
  }

  // compact index space of the bond-graph bonds
  if (Es >= 0) {
    CHG_CHECK_ARG(out->short_ids && out->ang_is && out->ang_js && out->ptr_is, "null pointer");
    int32_t* mask = cnt;     // [Eu]
    int32_t* slot = cursor;  // [Eu + 1]
    CHG_CUDA(cudaMemsetAsync(mask, 0, (size_t)(Eu + 1) * 4, st));
    if (A > 0) {
      mark_kernel<<<blocks(A), 256, 0, st>>>(in->ang_i, in->ang_j, A, mask);
      count_launch();
    }
    if ((rc = exclusive_scan(mask, Eu, slot, sums, st)) != CHG_OK) return rc;
    if (Eu > 0) {
      compact_ids_kernel<<<blocks(Eu), 256, 0, st>>>(mask, slot, Eu, out->short_ids);
      count_launch();
    }
    if (A > 0) {
      remap_kernel<<<blocks(A), 256, 0, st>>>(slot, in->ang_i, in->ang_j, A, out->ang_is, out->ang_js);
      count_launch();
    }
    csr_ptr_sorted_kernel<<<blocks(A + 1), 256, 0, st>>>(out->ang_is, A, Es, out->ptr_is);
    count_launch();
    if (in->with_reverse) {
      CHG_CHECK_ARG(out->ptr_js, "null pointer");
      gather_ptr_kernel<<<blocks(Es + 1), 256, 0, st>>>(out->ptr_j, out->short_ids, Es, A, out->ptr_js);
      count_launch();
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("chg_build_csr: launch failed: %s", cudaGetErrorString(e));
    return CHG_ERR_CUDA;
  }
  return CHG_OK;
}

// number of distinct bonds that occur as bond i or bond j of an angle (n_short of chg_build_csr) for index arrays that
// live on the device (the device graph builder): mark -> scan -> one int32 to the host (synchronises the stream).
extern "C" int chg_bond_graph_count(const int32_t* ang_i, const int32_t* ang_j, int32_t n_angles, int32_t n_bonds,
                                    int32_t* scratch /* 2 * (n_bonds + 1) + 4096 int32 */, int32_t* count_out, void* stream) {
  CHG_CHECK_ARG(n_angles >= 0 && n_bonds >= 0 && count_out != nullptr, "bad arguments");
  *count_out = 0;
  if (n_angles == 0 || n_bonds == 0) return CHG_OK;
  CHG_CHECK_ARG(ang_i && ang_j && scratch, "null pointer");
  cudaStream_t st = as_stream(stream);
  int32_t* mask = scratch;
  int32_t* slot = mask + (n_bonds + 1);
  int32_t* sums = slot + (n_bonds + 1);
  CHG_CUDA(cudaMemsetAsync(mask, 0, (size_t)(n_bonds + 1) * 4, st));
  mark_kernel<<<blocks(n_angles), 256, 0, st>>>(ang_i, ang_j, n_angles, mask);
  count_launch();
  const int rc = exclusive_scan_i32(mask, n_bonds, slot, sums, st);
  if (rc != CHG_OK) return rc;
  CHG_CUDA(cudaMemcpyAsync(count_out, slot + n_bonds, 4, cudaMemcpyDeviceToHost, st));
  CHG_CUDA(cudaStreamSynchronize(st));
  return CHG_OK;
}
