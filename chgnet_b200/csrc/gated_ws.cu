// AtomConv / BondConv message + aggregation as ONE warp-specialised tcgen05 kernel (sm_100a).
//
//   agg[s] = sum_{rows r of segment s} GatedMLP(pre_r) * w_r          (reference layers.py:113-126, 238-254)
//
// with pre_r = the gather-add of 3 (AtomConv) / 4 (BondConv) first-layer rows, the two
// 64x64 second-layer products on the tensor cores (3xTF32, accumulators in tensor memory), LayerNorm +
// SiLU x sigmoid in the epilogue, and the segmented reduction over the centre-sorted (bond-i-sorted) rows done
// inside the CTA: the [rows, 64] message never goes to HBM.
//
// One persistent CTA per SM, 17 warps (544 threads, <= 96 registers each), three roles connected by mbarriers:
//
//   warps 8-11  producer of the CORE half, warps 12-15 producer of the GATE half of every 128-row tile (two
//               independent groups, so one gathers while the other converts): 16 lanes x float4 per 64-float
//               half-row (coalesced) gather + add the first-layer rows, SiLU, store the [128 x 64] half-tile to
//               shared memory (16-byte chunks XOR-swizzled by row), group barrier, then thread t reads ITS row
//               (conflict-free), splits hi / lo and tcgen05.st's it into the group's A stage of tensor memory;
//               BondConv also writes save_pre
//   warp 16     MMA: one lane issues 8 k-steps x 3 split terms of tcgen05.mma.kind::tf32 (M=128, N=64) per
//               half into one of two D stages, tcgen05.commit -> mbarriers
//   warps 0-3   epilogue of the CORE half, warps 4-7 of the GATE half: thread t owns (half of) row t: two
//               tcgen05.ld sweeps (shifted one-pass mean / variance, then normalise: LayerNorm is in-thread, no
//               shuffles), SiLU (core) / sigmoid (gate) -> two [128 x 64] tiles in shared memory; then each of the
//               8 warps reduces a 16-row strip (core x gate x bond weights, the weights read coalesced) over runs
//               of equal segment id: complete segments are stored, strip-boundary partials go to `parts`
//
// A tiny second kernel (seg_stitch) adds the strip partials of every segment that spans strips, in strip
// order (deterministic, no atomics), and zeroes empty segments.
//
// TMEM: 2 x (64 hi + 64 lo) A + 2 x 128 D = 512 columns.  Shared memory: 4 weight images (64 KB) + 2 half-tiles
// (64 KB) + the core / gate output tiles (64 KB) + indices.
#include "gated_common.cuh"
#include "tc.cuh"

namespace chg {
namespace gated {
namespace {

constexpr int WS_THREADS = 544;  // 17 warps: 8 epilogue, 8 producer, 1 MMA
constexpr int TR = 128;          // rows per tile
constexpr int HALF_BYTES = TR * 64 * 4;
constexpr int IMG_BYTES = 64 * 64 * 4;
constexpr int STRIP = 16;        // rows per reduction strip (one half-warp)

struct WsSmem {
  static constexpr int IMG_OFF = 0;                          // Bc_hi, Bc_lo, Bg_hi, Bg_lo
  static constexpr int HS_OFF = 4 * IMG_BYTES;               // 2 x [128][64] fp32, swizzled (core | gate producer)
  static constexpr int O_OFF = HS_OFF + 2 * HALF_BYTES;      // 2 x [128][64] fp32, swizzled: silu(core), sigmoid(gate)
  static constexpr int GIDX_OFF = O_OFF + 2 * HALF_BYTES;    // 2 groups x 3 x 128 int
  static constexpr int EIDX_OFF = GIDX_OFF + 2 * 3 * TR * 4;  // 4 x 128 int: segment id, weight row, segment begin / end
  static constexpr int B2_OFF = EIDX_OFF + 4 * TR * 4;       // 128 floats
  static constexpr int LN_OFF = B2_OFF + 128 * 4;            // 256 floats
  static constexpr int TOTAL = LN_OFF + 256 * 4;
};

struct WsBars {
  uint64_t a_full[2], a_empty[2];
  uint64_t d_full[2], d_empty[2];
};

struct FusedArgs {
  const float* p_a;     // ATOM: pcn [N][256]        BOND: pij [Es][256]
  const float* p_b;     // ATOM: pe  [Eu][128]       BOND: px  [N][128]
  const float* p_c;     // BOND: pa [A][128], else null
  const float* wgt;     // ATOM: wag [Eu][64]        BOND: wbg_s [Es][64]
  const int32_t* idx0;  // row of p_a, first half = the SEGMENT id  (center | bond slot i)
  const int32_t* idx1;  // row of p_a, second half                  (nbr    | bond slot j)
  const int32_t* idx2;  // row of p_b                               (d2u    | atom)
  const int32_t* ptr;   // [n_seg + 1] CSR of idx0
  int32_t n_rows, n_seg;
  const float* w2t;     // [64][128]
  const float* b2;      // [128]
  const float* ln;      // [4][64] or null
  float* out;           // [n_seg][64]
  float* parts;         // [ceil(n_rows / 16)][2][64]
  float* save_pre;      // [rows][128] or null
  float* save_p;        // [rows][128] or null
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// byte offset of 16-byte chunk c (0..15) of row r in a swizzled [128][64] fp32 tile
__device__ __forceinline__ int swz(int r, int c) { return r * 256 + ((c ^ (r & 7)) << 4); }

// image element (n, kk) = src[kk * ld + col0 + n]   (64 x 64, K-major, no swizzle), all threads of the CTA
__device__ __forceinline__ void build_image_ws(uint8_t* hi, uint8_t* lo, const float* __restrict__ src, int ld, int col0,
                                               int tid) {
  for (int i = tid; i < 4096; i += WS_THREADS) {
    const int kk = i >> 6, n = i & 63;
    uint32_t h, l;
    tc::split_tf32(__ldg(src + (size_t)kk * ld + col0 + n), h, l);
    const uint32_t off = tc::kmajor_offset(n, kk, 64);
    *reinterpret_cast<uint32_t*>(hi + off) = h;
    *reinterpret_cast<uint32_t*>(lo + off) = l;
  }
}

// ---- epilogue helper: one 64-wide half of row t out of tensor memory ----------------------------------------
// p = acc + b2; y = LN(p) (or p); CORE: out = silu(y);  GATE: out = sigmoid(y); out -> row t of the swizzled tile.
// LayerNorm statistics in one sweep, shifted by the row's first element (no cancellation for |mean| >> std).
template <bool GATE>
__device__ __forceinline__ void epilogue_half(uint32_t d_addr, const float* s_b2, const float* s_gamma, const float* s_beta,
                                              bool use_ln, uint8_t* s_tile, int t, float* save_row) {
  float mean = 0.f, rstd = 1.f;
  if (use_ln) {
    float s1 = 0.f, s2 = 0.f, s1b = 0.f, s2b = 0.f, x0 = 0.f;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint32_t v[32];
      tc::tmem_ld32(d_addr + g * 32, v);
      tc::tmem_ld_wait();
      if (g == 0) x0 = __uint_as_float(v[0]) + s_b2[0];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b = lds4(s_b2 + g * 32 + q * 4);  // broadcast
        const float d0 = __uint_as_float(v[q * 4]) + b.x - x0, d1 = __uint_as_float(v[q * 4 + 1]) + b.y - x0;
        const float d2 = __uint_as_float(v[q * 4 + 2]) + b.z - x0, d3 = __uint_as_float(v[q * 4 + 3]) + b.w - x0;
        s1 += d0 + d1;
        s1b += d2 + d3;
        s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
        s2b = fmaf(d2, d2, fmaf(d3, d3, s2b));
      }
    }
    const float m1 = (s1 + s1b) * (1.f / 64.f);
    mean = x0 + m1;
    rstd = 1.f / sqrtf(fmaxf((s2 + s2b) * (1.f / 64.f) - m1 * m1, 0.f) + LN_EPS);
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    uint32_t v[32];
    tc::tmem_ld32(d_addr + g * 32, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = lds4(s_b2 + g * 32 + q * 4);
      float4 p = make_float4(__uint_as_float(v[q * 4]) + b.x, __uint_as_float(v[q * 4 + 1]) + b.y,
                             __uint_as_float(v[q * 4 + 2]) + b.z, __uint_as_float(v[q * 4 + 3]) + b.w);
      if (save_row != nullptr) stg4(save_row + g * 32 + q * 4, p);
      if (use_ln) {
        const float4 ga = lds4(s_gamma + g * 32 + q * 4), be = lds4(s_beta + g * 32 + q * 4);
        p.x = fmaf((p.x - mean) * rstd, ga.x, be.x);
        p.y = fmaf((p.y - mean) * rstd, ga.y, be.y);
        p.z = fmaf((p.z - mean) * rstd, ga.z, be.z);
        p.w = fmaf((p.w - mean) * rstd, ga.w, be.w);
      }
      float4 o;
      if (GATE) {
        o = make_float4(sigmoid_f(p.x), sigmoid_f(p.y), sigmoid_f(p.z), sigmoid_f(p.w));
      } else {
        o = make_float4(silu_f(p.x), silu_f(p.y), silu_f(p.z), silu_f(p.w));
      }
      *reinterpret_cast<float4*>(s_tile + swz(t, g * 8 + q)) = o;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(WS_THREADS, 1) gated_ws_fwd_kernel(const FusedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* s_img = smem_raw + WsSmem::IMG_OFF;
  uint8_t* s_hs = smem_raw + WsSmem::HS_OFF;
  uint8_t* s_o = smem_raw + WsSmem::O_OFF;
  int* s_gidx = reinterpret_cast<int*>(smem_raw + WsSmem::GIDX_OFF);
  int* s_eidx = reinterpret_cast<int*>(smem_raw + WsSmem::EIDX_OFF);
  float* s_b2 = reinterpret_cast<float*>(smem_raw + WsSmem::B2_OFF);
  float* s_ln = reinterpret_cast<float*>(smem_raw + WsSmem::LN_OFF);
  __shared__ __align__(8) WsBars bars;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool use_ln = a.ln != nullptr;
  const int n_tiles = (a.n_rows + TR - 1) / TR;

  // ---- one-time setup: weight images, bias / LayerNorm rows, barriers, tensor memory ------------------------
  build_image_ws(s_img, s_img + IMG_BYTES, a.w2t, 128, 0, tid);                     // core: (n=c, kk=k) = w2t[k][c]
  build_image_ws(s_img + 2 * IMG_BYTES, s_img + 3 * IMG_BYTES, a.w2t, 128, 64, tid);  // gate
  if (tid < 128) s_b2[tid] = a.b2[tid];
  if (tid < 256) s_ln[tid] = use_ln ? a.ln[tid] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars.a_full[i], 128);
      tc::mbar_init(&bars.a_empty[i], 1);
      tc::mbar_init(&bars.d_full[i], 1);
      tc::mbar_init(&bars.d_empty[i], 256);
    }
    tc::mbar_fence_init();
  }
  if (warp == 16) tc::tmem_alloc(&s_tmem, 512);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = s_tmem;
  // TMEM columns: A stage h (h = 0 core, 1 gate) at h*128 (hi) / h*128 + 64 (lo); D stage s at 256 + s*128
  // (core 0..63 | gate 64..127)

  if (warp >= 8 && warp < 16) {
    // ============================ producer groups (gather -> SiLU -> A operand) ============================
    const int half = (warp - 8) >> 2;            // 0: core columns, 1: gate columns
    const int gt = tid - 256 - half * 128;       // 0..127 inside the group; also this thread's tile row / TMEM lane
    const int tx = gt & 15, ty = gt >> 4;        // 16 lanes per row, 8 rows per pass
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    int* gi = s_gidx + half * 3 * TR;
    uint8_t* stage = s_hs + half * HALF_BYTES;
    const int col = half * 64 + tx * 4;
    const uint32_t a_hi = tmem_base + half * 128 + lane_sel, a_lo = a_hi + 64;
    int tl = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      const int base = tile * TR;
      {
        const int r = min(base + gt, a.n_rows - 1);
        gi[gt] = a.idx0[r];
        gi[TR + gt] = a.idx1[r];
        gi[2 * TR + gt] = a.idx2[r];
      }
      tc::wg_barrier(1 + half, 128);  // indices visible; every thread of the group is done with the previous half-tile
#pragma unroll 1
      for (int b = 0; b < 4; ++b) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = ty + 8 * (b * 4 + i);
          const float* s0 = a.p_a + (size_t)gi[row] * 256 + col;
          const float* s1 = a.p_a + (size_t)gi[TR + row] * 256 + 128 + col;
          const float* s2 = a.p_b + (size_t)gi[2 * TR + row] * 128 + col;
          v[i] = ldg4(s0) + ldg4(s1) + ldg4(s2);
          if (MODE == BOND) v[i] = v[i] + ldg4(a.p_c + (size_t)min(base + row, a.n_rows - 1) * 128 + col);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = ty + 8 * (b * 4 + i);
          if (a.save_pre != nullptr && base + row < a.n_rows) stg4(a.save_pre + (size_t)(base + row) * 128 + col, v[i]);
          *reinterpret_cast<float4*>(stage + swz(row, tx)) =
              make_float4(silu_f(v[i].x), silu_f(v[i].y), silu_f(v[i].z), silu_f(v[i].w));
        }
      }
      tc::wg_barrier(1 + half, 128);  // the half-tile is complete in shared memory
      tc::mbar_wait(&bars.a_empty[half], (tl & 1) ^ 1);  // the MMAs of the previous tile have read this A stage
      tc::fence_after_sync();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(stage + swz(gt, g * 4 + q));
          tc::split_tf32(v.x, hi[q * 4 + 0], lo[q * 4 + 0]);
          tc::split_tf32(v.y, hi[q * 4 + 1], lo[q * 4 + 1]);
          tc::split_tf32(v.z, hi[q * 4 + 2], lo[q * 4 + 2]);
          tc::split_tf32(v.w, hi[q * 4 + 3], lo[q * 4 + 3]);
        }
        tc::tmem_st16(a_hi + g * 16, hi);
        tc::tmem_st16(a_lo + g * 16, lo);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      mbar_arrive(&bars.a_full[half]);
    }
  } else if (warp == 16) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc = tc::idesc_tf32(128, 64);
      const uint32_t img = tc::smem_u32(s_img);
      int tl = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
        const int ds = tl & 1;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          tc::mbar_wait(&bars.a_full[half], tl & 1);
          if (half == 0) tc::mbar_wait(&bars.d_empty[ds], ((tl >> 1) & 1) ^ 1);
          tc::fence_after_sync();
          const uint32_t d_acc = tmem_base + 256 + ds * 128 + half * 64;
          const uint32_t a_hi = tmem_base + half * 128, a_lo = a_hi + 64;
          const uint32_t bhi = img + half * 2 * IMG_BYTES, blo = bhi + IMG_BYTES;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint64_t bh = tc::smem_desc_kmajor(bhi + j * 256, 128, 2048);
            const uint64_t bl = tc::smem_desc_kmajor(blo + j * 256, 128, 2048);
            tc::mma_tf32_ts(d_acc, a_hi + j * 8, bh, idesc, j > 0 ? 1u : 0u);
            tc::mma_tf32_ts(d_acc, a_lo + j * 8, bh, idesc, 1u);
            tc::mma_tf32_ts(d_acc, a_hi + j * 8, bl, idesc, 1u);
          }
          tc::mma_commit(&bars.a_empty[half]);  // A stage reusable once these MMAs have read it
          if (half == 1) tc::mma_commit(&bars.d_full[ds]);
        }
      }
    }
  } else if (warp < 8) {
    // ============================ epilogue: warps 0-3 core half, warps 4-7 gate half ============================
    const int eh = warp >> 2;        // 0: core, 1: gate
    const int t = tid & 127;         // tile row == TMEM lane
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    int* s_seg = s_eidx;             // segment id of every tile row
    int* s_wrow = s_eidx + TR;       // row of the bond weight (ATOM: d2u; BOND: slot j - slot i is the segment id)
    int* s_sa = s_eidx + 2 * TR;     // ptr[seg], ptr[seg + 1] of every tile row
    int* s_sb = s_eidx + 3 * TR;
    uint8_t* s_o1 = s_o;
    uint8_t* s_o2 = s_o + HALF_BYTES;
    int tl = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      const int base = tile * TR;
      const int ds = tl & 1;
      {
        const int r = min(base + t, a.n_rows - 1);
        if (eh == 0) {
          const int seg = a.idx0[r];
          s_seg[t] = seg;
          s_sa[t] = __ldg(a.ptr + seg);
          s_sb[t] = __ldg(a.ptr + seg + 1);
        } else {
          s_wrow[t] = MODE == BOND ? a.idx1[r] : a.idx2[r];
        }
      }
      tc::mbar_wait(&bars.d_full[ds], (tl >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t d_acc = tmem_base + 256 + ds * 128 + eh * 64 + lane_sel;
      float* save_row = (a.save_p != nullptr && base + t < a.n_rows) ? a.save_p + (size_t)(base + t) * 128 + eh * 64 : nullptr;
      if (eh == 0) {
        epilogue_half<false>(d_acc, s_b2, s_ln, s_ln + 64, use_ln, s_o1, t, save_row);
      } else {
        epilogue_half<true>(d_acc, s_b2 + 64, s_ln + 128, s_ln + 192, use_ln, s_o2, t, save_row);
      }
      tc::fence_before_sync();
      mbar_arrive(&bars.d_empty[ds]);  // the accumulator stage can be overwritten
      tc::wg_barrier(3, 256);

      // ---- segmented reduction: warp w owns the 16-row strip w, lane l the columns 2l, 2l+1 ----------------
      const int strip_lo = base + warp * STRIP;
      const int strip_hi = min(strip_lo + STRIP, a.n_rows);
      if (strip_lo < a.n_rows) {
        float2 w[STRIP];
#pragma unroll
        for (int i = 0; i < STRIP; ++i) {
          const int rr = warp * STRIP + i;
          const float2 w0 = __ldg(reinterpret_cast<const float2*>(a.wgt + (size_t)s_wrow[rr] * 64) + lane);
          if (MODE == ATOM) {
            w[i] = w0;
          } else {
            const float2 wi = __ldg(reinterpret_cast<const float2*>(a.wgt + (size_t)s_seg[rr] * 64) + lane);
            w[i] = make_float2(wi.x * w0.x, wi.y * w0.y);  // (o * w_i) * w_j
          }
        }
        float2 acc = make_float2(0.f, 0.f);
        int cur_row = warp * STRIP;  // first row of the current run
        auto emit = [&](int row0, const float2& v) {
          const int seg = s_seg[row0], sa = s_sa[row0], sb = s_sb[row0];
          float* dst;
          if (sa >= strip_lo && sb <= strip_hi) {
            dst = a.out + (size_t)seg * 64;  // the whole segment lies in this strip
          } else {
            dst = a.parts + ((size_t)(strip_lo / STRIP) * 2 + (sa < strip_lo ? 0 : 1)) * 64;
          }
          *reinterpret_cast<float2*>(dst + lane * 2) = v;
        };
#pragma unroll
        for (int i = 0; i < STRIP; ++i) {
          const int rr = warp * STRIP + i;
          if (strip_lo + i < strip_hi) {
            if (s_seg[rr] != s_seg[cur_row]) {
              emit(cur_row, acc);
              cur_row = rr;
              acc = make_float2(0.f, 0.f);
            }
            const int off = swz(rr, lane >> 1) + (lane & 1) * 8;
            const float2 c = *reinterpret_cast<const float2*>(s_o1 + off);
            const float2 g = *reinterpret_cast<const float2*>(s_o2 + off);
            acc.x = fmaf(c.x * g.x, w[i].x, acc.x);
            acc.y = fmaf(c.y * g.y, w[i].y, acc.y);
          }
        }
        emit(cur_row, acc);
      }
      tc::wg_barrier(3, 256);  // the output tiles and the index rows are free for the next tile
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 16) tc::tmem_dealloc(tmem_base, 512);
}

// out[s] for every segment that spans more than one strip (sum of its strip partials, in strip order) and for
// every empty segment (zeros); segments inside one strip were stored by the main kernel.
__global__ void seg_stitch_kernel(const int32_t* __restrict__ ptr, int n_seg, const float* __restrict__ parts,
                                  float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = idx >> 4, l = idx & 15;
  if (s >= n_seg) return;
  const int a = ptr[s], b = ptr[s + 1];
  if (a >= b) {
    stg4(out + (size_t)s * 64 + l * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  const int k0 = a / STRIP, k1 = (b - 1) / STRIP;
  if (k0 == k1) return;
  float4 acc = ldg4(parts + ((size_t)k0 * 2 + 1) * 64 + l * 4);
  for (int k = k0 + 1; k <= k1; ++k) acc = acc + ldg4(parts + ((size_t)k * 2) * 64 + l * 4);
  stg4(out + (size_t)s * 64 + l * 4, acc);
}

template <int MODE>
int launch_fused(const FusedArgs& a, cudaStream_t stream) {
  if (a.n_seg == 0) return CHG_OK;
  if (a.n_rows > 0) {
    static bool attr_set = false;
    if (!attr_set) {
      CHG_CUDA(cudaFuncSetAttribute(gated_ws_fwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, WsSmem::TOTAL));
      attr_set = true;
    }
    const int n_tiles = (a.n_rows + TR - 1) / TR;
    gated_ws_fwd_kernel<MODE><<<min(n_tiles, sm_count()), WS_THREADS, WsSmem::TOTAL, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("gated_ws_fwd_kernel: launch failed: %s", cudaGetErrorString(e));
      return CHG_ERR_CUDA;
    }
    count_launch();
  }
  seg_stitch_kernel<<<(a.n_seg * 16 + 255) / 256, 256, 0, stream>>>(a.ptr, a.n_seg, a.parts, a.out);
  CHG_LAUNCH_END();
}

}  // namespace

int atom_conv_fused_ws(const float* pcn, const float* pe, const float* wag, const int32_t* center, const int32_t* nbr,
                       const int32_t* d2u, const int32_t* ptr_c, int n_edges, int n_atoms, const float* w2t, const float* b2,
                       const float* ln, float* agg, float* save_p, float* parts, cudaStream_t stream) {
  FusedArgs a{pcn, pe, nullptr, wag, center, nbr, d2u, ptr_c, n_edges, n_atoms, w2t, b2, ln, agg, parts, nullptr, save_p};
  return launch_fused<ATOM>(a, stream);
}

int bond_conv_fused_ws(const float* pij, const float* px, const float* pa, const float* wbg, const int32_t* ang_atom,
                       const int32_t* ang_i, const int32_t* ang_j, const int32_t* ptr_i, int n_angles, int n_slots,
                       const float* w2t, const float* b2, const float* ln, float* agg, float* save_pre, float* save_p,
                       float* parts, cudaStream_t stream) {
  FusedArgs a{pij, px, pa, wbg, ang_i, ang_j, ang_atom, ptr_i, n_angles, n_slots, w2t, b2, ln, agg, parts, save_pre, save_p};
  return launch_fused<BOND>(a, stream);
}

}  // namespace gated
}  // namespace chg

using namespace chg;

// ---- C ABI: fused message + aggregation (include/chgnet_b200.h) -------------------------------------------------
extern "C" int64_t chg_gated_fused_workspace_floats(int32_t n_rows) {
  if (n_rows < 0) return -1;
  const int64_t msg = (int64_t)n_rows * 64, parts = ((int64_t)n_rows + 15) / 16 * 128;
  return (msg > parts ? msg : parts) + 64;
}

extern "C" int chg_atom_conv_fused(const float* pcn, const float* pe, const float* wag, const int32_t* center,
                                   const int32_t* nbr, const int32_t* d2u, const int32_t* ptr_c, int32_t n_edges,
                                   int32_t n_atoms, const float* w2t, const float* b2, const float* ln, float* agg,
                                   float* save_p, float* work, void* stream) {
  CHG_CHECK_ARG(n_edges >= 0 && n_atoms >= 0, "negative size");
  if (n_atoms == 0) return CHG_OK;
  CHG_CHECK_ARG(ptr_c && agg, "null pointer");
  CHG_CHECK_ARG(n_edges == 0 || (pcn && pe && wag && center && nbr && d2u && w2t && b2 && work), "null pointer");
  if (gated_impl() == 3)
    return gated::atom_conv_fused_ws(pcn, pe, wag, center, nbr, d2u, ptr_c, n_edges, n_atoms, w2t, b2, ln, agg, save_p, work,
                                     as_stream(stream));
  // A/B implementations 0..2: the unfused pair (message kernel -> segmented sum), the message in `work`
  const int rc = chg_atom_conv_fwd(pcn, pe, wag, center, nbr, d2u, n_edges, w2t, b2, ln, work, save_p, nullptr, stream);
  if (rc != CHG_OK) return rc;
  return chg_segment_sum(work, 64, nullptr, ptr_c, n_atoms, n_edges, 0, agg, 64, stream);
}

extern "C" int chg_bond_conv_fused(const float* pij, const float* px, const float* pa, const float* wbg,
                                   const int32_t* ang_atom, const int32_t* ang_i, const int32_t* ang_j,
                                   const int32_t* ptr_i, int32_t n_angles, int32_t n_slots, const float* w2t,
                                   const float* b2, const float* ln, float* agg, float* save_pre, float* save_p,
                                   float* work, void* stream) {
  CHG_CHECK_ARG(n_angles >= 0 && n_slots >= 0, "negative size");
  if (n_slots == 0) return CHG_OK;
  CHG_CHECK_ARG(ptr_i && agg, "null pointer");
  CHG_CHECK_ARG(n_angles == 0 || (pij && px && pa && wbg && ang_atom && ang_i && ang_j && w2t && b2 && work), "null pointer");
  if (gated_impl() == 3)
    return gated::bond_conv_fused_ws(pij, px, pa, wbg, ang_atom, ang_i, ang_j, ptr_i, n_angles, n_slots, w2t, b2, ln, agg,
                                     save_pre, save_p, work, as_stream(stream));
  const int rc = chg_bond_conv_fwd(pij, px, pa, wbg, ang_atom, ang_i, ang_j, n_angles, w2t, b2, ln, work, save_pre, save_p, stream);
  if (rc != CHG_OK) return rc;
  return chg_segment_sum(work, 64, nullptr, ptr_i, n_slots, n_angles, 0, agg, 64, stream);
}
