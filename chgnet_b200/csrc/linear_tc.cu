// chg_linear on the 5th-generation tensor cores (tcgen05, sm_100a), 3xTF32.
//
//   y[yr] = x[xr] @ wt (+ bias) (+ residual[yr])        x rows of k floats, wt [k][n_out]
//
// One persistent CTA per SM, two warpgroups working on alternating 128-row tiles so that one
// group's loads / epilogue overlap the other group's MMAs.  Thread t of a warpgroup owns row t
// of its tile end to end:
//   global rows (64-float chunks, coalesced) -> shared staging -> thread t reads ITS row ->
//   hi/lo TF32 split -> tcgen05.st into the warpgroup's A region of tensor memory (lane t,
//   K along columns)
//   one elected thread: 8 k-steps x 3 split terms of tcgen05.mma.kind::tf32 (A from TMEM,
//   B = the [NT x k] weight panel, resident in shared memory as hi and lo images in the
//   K-major no-swizzle canonical layout), tcgen05.commit -> mbarrier
//   tcgen05.ld of the fp32 accumulator row -> shared staging -> + bias/residual -> coalesced
//   global rows (thread-per-row 16-byte global accesses are transaction-bound: 8x the L1<->L2
//   transactions of the staged version)
// No operand ever passes through the LSU/shared-memory crossbar as an FFMA operand, which is
// what bounds the FFMA version (profiles/SUMMARY_r01.md).  TMEM: 2 x (64 hi + 64 lo + 128 D)
// = 512 columns.
#include "common.cuh"
#include "tc.cuh"

namespace chg {
namespace {

constexpr int NTHR = 256;

constexpr int IN_LD = 68;   // staging stride (floats) of a 64-float input chunk row
constexpr int OUT_LD = 36;  // staging stride of a 32-float output chunk row
constexpr int STAGE_FLOATS = 128 * IN_LD;

template <int NT>
__global__ void __launch_bounds__(NTHR, 1)
linear_tc_kernel(const float* __restrict__ x, const int32_t* __restrict__ x_rows, int m, int k,
                 const float* __restrict__ wt, const float* __restrict__ bias, const float* residual,
                 const int32_t* __restrict__ y_rows, int n_out, float* y) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* s_bhi = smem_raw;
  uint8_t* s_blo = smem_raw + (size_t)NT * k * 4;
  float* s_stage_all = reinterpret_cast<float*>(smem_raw + (size_t)2 * NT * k * 4);
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_tmem;
  __shared__ int s_yrow[2][128];

  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp = tid >> 5;
  const int col_base = blockIdx.y * NT;
  float* stage = s_stage_all + wg * STAGE_FLOATS;

  // weight panel -> hi / lo operand images (element (n, kk) = wt[kk][col_base + n]); float4 loads,
  // all of a thread's loads issued before the first dependent store
  {
    constexpr int PER = 8;  // float4 per thread per round
    const int n4 = NT / 4, total4 = n4 * k;
    for (int base4 = 0; base4 < total4; base4 += NTHR * PER) {
      float4 v[PER];
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int i4 = base4 + q * NTHR + tid;
        v[q] = i4 < total4 ? ldg4(wt + (size_t)(i4 / n4) * n_out + col_base + (i4 % n4) * 4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int i4 = base4 + q * NTHR + tid;
        if (i4 < total4) {
          const int kk = i4 / n4, n0 = (i4 % n4) * 4;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t hi, lo;
            tc::split_tf32(f4at(v[q], e), hi, lo);
            const uint32_t off = tc::kmajor_offset(n0 + e, kk, k);
            *reinterpret_cast<uint32_t*>(s_bhi + off) = hi;
            *reinterpret_cast<uint32_t*>(s_blo + off) = lo;
          }
        }
      }
    }
  }
  if (tid == 0) {
    tc::mbar_init(&s_bar[0], 1);
    tc::mbar_init(&s_bar[1], 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  const uint32_t tmem_base = s_tmem;
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t a_hi = tmem_base + wg * 256, a_lo = a_hi + 64, d_acc = a_hi + 128;
  const uint32_t idesc = tc::idesc_tf32(128, NT);
  const uint32_t bhi_addr = tc::smem_u32(s_bhi), blo_addr = tc::smem_u32(s_blo);
  const uint32_t sbo = (uint32_t)(k / 4) * 128;
  const int bar_id = 1 + wg;
  uint32_t phase = 0;

  const int n_tiles = (m + 127) / 128;
  const int c4_in = t & 15, row0_in = t >> 4;
  // software pipeline: the 16 row-chunk loads of the NEXT tile are issued right after this
  // tile's MMAs, so their latency hides behind the MMA wait and the epilogue
  float4 pre[16];
  auto issue_loads = [&](int tile_, int kc_) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int r = min(tile_ * 128 + row0_in + q * 8, m - 1);
      const int xr = x_rows != nullptr ? __ldg(x_rows + r) : r;
      pre[q] = ldg4(x + (size_t)xr * k + kc_ + c4_in * 4);
    }
  };
  int tile = blockIdx.x * 2 + wg;
  if (tile < n_tiles) issue_loads(tile, 0);
  for (; tile < n_tiles; tile += gridDim.x * 2) {
    const int base = tile * 128;
    {
      const int r = min(base + t, m - 1);
      s_yrow[wg][t] = (base + t < m) ? (y_rows != nullptr ? __ldg(y_rows + r) : r) : -1;
    }

    for (int kc = 0; kc < k; kc += 64) {
#pragma unroll
      for (int q = 0; q < 16; ++q) sts4(stage + (row0_in + q * 8) * IN_LD + c4_in * 4, pre[q]);
      tc::wg_barrier(bar_id, 128);
      // thread t: its own row -> hi/lo split -> tensor memory (lane t)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = lds4(stage + t * IN_LD + g * 16 + q * 4);
          tc::split_tf32(v.x, hi[q * 4 + 0], lo[q * 4 + 0]);
          tc::split_tf32(v.y, hi[q * 4 + 1], lo[q * 4 + 1]);
          tc::split_tf32(v.z, hi[q * 4 + 2], lo[q * 4 + 2]);
          tc::split_tf32(v.w, hi[q * 4 + 3], lo[q * 4 + 3]);
        }
        tc::tmem_st16(a_hi + lane_sel + g * 16, hi);
        tc::tmem_st16(a_lo + lane_sel + g * 16, lo);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      tc::wg_barrier(bar_id, 128);
      if (t == 0) {
        tc::fence_after_sync();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t koff = (uint32_t)((kc + 8 * j) / 4) * 128;
          const uint64_t bh = tc::smem_desc_kmajor(bhi_addr + koff, 128, sbo);
          const uint64_t bl = tc::smem_desc_kmajor(blo_addr + koff, 128, sbo);
          tc::mma_tf32_ts(d_acc, a_hi + j * 8, bh, idesc, (kc > 0 || j > 0) ? 1u : 0u);
          tc::mma_tf32_ts(d_acc, a_lo + j * 8, bh, idesc, 1u);
          tc::mma_tf32_ts(d_acc, a_hi + j * 8, bl, idesc, 1u);
        }
        tc::mma_commit(&s_bar[wg]);
      }
      {  // prefetch: next K chunk of this tile, or the first chunk of this warpgroup's next tile
        const bool more_k = kc + 64 < k;
        const int nt = more_k ? tile : tile + (int)gridDim.x * 2;
        if (nt < n_tiles) issue_loads(nt, more_k ? kc + 64 : 0);
      }
      tc::mbar_wait(&s_bar[wg], phase);
      phase ^= 1;
      tc::fence_after_sync();
    }

    // accumulator row -> staging (32 columns at a time) -> coalesced global rows
#pragma unroll 1
    for (int c = 0; c < NT; c += 32) {
      uint32_t v[16], w[16];
      tc::tmem_ld16(d_acc + lane_sel + c, v);
      tc::tmem_ld16(d_acc + lane_sel + c + 16, w);
      tc::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        sts4(stage + t * OUT_LD + q * 4, make_float4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                                                     __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3])));
        sts4(stage + t * OUT_LD + 16 + q * 4,
             make_float4(__uint_as_float(w[q * 4 + 0]), __uint_as_float(w[q * 4 + 1]), __uint_as_float(w[q * 4 + 2]),
                         __uint_as_float(w[q * 4 + 3])));
      }
      tc::wg_barrier(bar_id, 128);
      {
        const int c4 = t & 7, row0 = t >> 3;
        const int col = col_base + c + c4 * 4;
        int orow[8];
        float4 o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) orow[q] = s_yrow[wg][row0 + q * 16];
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr) b4 = ldg4(bias + col);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          o[q] = b4;
          if (residual != nullptr && orow[q] >= 0)
            o[q] = o[q] + *reinterpret_cast<const float4*>(residual + (size_t)orow[q] * n_out + col);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = o[q] + lds4(stage + (row0 + q * 16) * OUT_LD + c4 * 4);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (orow[q] >= 0) stg4(y + (size_t)orow[q] * n_out + col, o[q]);
      }
      tc::wg_barrier(bar_id, 128);
    }
    tc::fence_before_sync();  // accumulator reads ordered before the next tile's MMAs
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

template <int NT>
int launch_linear_tc(const float* x, const int32_t* x_rows, int m, int k, const float* wt, const float* bias,
                     const float* residual, const int32_t* y_rows, int n_out, float* y, cudaStream_t stream) {
  const int smem = 2 * NT * k * 4 + 2 * STAGE_FLOATS * 4;
  static int max_smem_set = 0;
  if (smem > max_smem_set) {
    CHG_CUDA(cudaFuncSetAttribute(linear_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    max_smem_set = smem;
  }
  const int n_tiles = (m + 127) / 128;
  const int col_tiles = n_out / NT;
  const int per_col = max(1, sm_count() / col_tiles);
  dim3 grid(min((n_tiles + 1) / 2, per_col), col_tiles);
  linear_tc_kernel<NT><<<grid, NTHR, smem, stream>>>(x, x_rows, m, k, wt, bias, residual, y_rows, n_out, y);
  CHG_LAUNCH_END();
}

}  // namespace

int linear_tc(const float* x, const int32_t* x_rows, int m, int k, const float* wt, const float* bias,
              const float* residual, const int32_t* y_rows, int n_out, float* y, cudaStream_t stream) {
  if (n_out % 128 == 0 && k <= 128)
    return launch_linear_tc<128>(x, x_rows, m, k, wt, bias, residual, y_rows, n_out, y, stream);
  return launch_linear_tc<64>(x, x_rows, m, k, wt, bias, residual, y_rows, n_out, y, stream);
}

}  // namespace chg
