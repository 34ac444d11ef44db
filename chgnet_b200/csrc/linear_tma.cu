// chg_linear, TMA-fed tcgen05 version (k <= 128):  y[yr] = x[xr] @ wt (+ bias) (+ residual[yr])
//
// Same tensor-core core as linear_tc.cu (3xTF32, A operand in tensor memory, weight panel
// images in shared memory, two warpgroups per CTA on alternating 128-row tiles), but no
// synchronous global traffic on the critical path:
//   * input rows arrive by cp.async.bulk (one 256-byte row chunk per thread, padded smem rows),
//     completion counted in bytes on an mbarrier (expect_tx); the copy of the NEXT work item is
//     issued as soon as the current rows have been consumed, so it overlaps MMA + epilogue;
//   * each thread drains ITS accumulator row (tcgen05.ld) into its own padded smem row and
//     ships it with a cp.async.bulk shared->global copy tracked by its own bulk group —
//     the epilogue needs no barrier at all (double-buffered 32-column chunks).
// Row gather (x_rows) / scatter (y_rows) cost nothing extra: every row is its own copy.
#include "common.cuh"
#include "tc.cuh"

namespace chg {
namespace {

constexpr int NTHR = 256;
constexpr int IN_LD = 68;    // padded staging row (floats): conflict-free 16-byte row-per-thread reads
constexpr int OUT_LD = 36;   // padded 32-column output chunk row
constexpr int IN_FLOATS = 128 * IN_LD;
constexpr int OUT_FLOATS = 128 * OUT_LD;
constexpr int WG_FLOATS = IN_FLOATS + 2 * OUT_FLOATS;

template <int NT>
__global__ void __launch_bounds__(NTHR, 1)
linear_tma_kernel(const float* __restrict__ x, const int32_t* __restrict__ x_rows, int m, int k,
                  const float* __restrict__ wt, const float* __restrict__ bias, const float* residual,
                  const int32_t* __restrict__ y_rows, int n_out, float* y) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* s_bhi = smem_raw;
  uint8_t* s_blo = smem_raw + (size_t)NT * k * 4;
  float* s_wg_all = reinterpret_cast<float*>(smem_raw + (size_t)2 * NT * k * 4);
  __shared__ __align__(8) uint64_t s_bar_mma[2], s_bar_load[2];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[NT];

  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp = tid >> 5;
  const int col_base = blockIdx.y * NT;
  float* s_in = s_wg_all + wg * WG_FLOATS;
  float* s_out = s_in + IN_FLOATS;

  for (int i = tid; i < NT * k; i += NTHR) {
    const int kk = i / NT, n = i % NT;
    uint32_t hi, lo;
    tc::split_tf32(__ldg(wt + (size_t)kk * n_out + col_base + n), hi, lo);
    const uint32_t off = tc::kmajor_offset(n, kk, k);
    *reinterpret_cast<uint32_t*>(s_bhi + off) = hi;
    *reinterpret_cast<uint32_t*>(s_blo + off) = lo;
  }
  if (tid < NT) s_bias[tid] = bias != nullptr ? bias[col_base + tid] : 0.f;
  if (tid == 0) {
    tc::mbar_init(&s_bar_mma[0], 1);
    tc::mbar_init(&s_bar_mma[1], 1);
    tc::mbar_init(&s_bar_load[0], 1);
    tc::mbar_init(&s_bar_load[1], 1);
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
  uint32_t ph_mma = 0, ph_load = 0;
  int out_buf = 0;

  const int n_tiles = (m + 127) / 128;
  const int k_chunks = k / 64;
  const int stride = gridDim.x * 2;

  // one bulk copy per thread: row t of work item (tile_, kc_) -> padded staging row t
  auto issue_load = [&](int tile_, int kc_) {
    const int base_ = tile_ * 128;
    const int nvalid = min(128, m - base_);
    if (t == 0) tc::mbar_expect_tx(&s_bar_load[wg], (uint32_t)nvalid * 256u);
    tc::wg_barrier(bar_id, 128);  // expect_tx is posted before any copy can complete
    if (t < nvalid) {
      const int r = base_ + t;
      const int xr = x_rows != nullptr ? __ldg(x_rows + r) : r;
      tc::bulk_load(s_in + t * IN_LD, x + (size_t)xr * k + kc_, 256u, &s_bar_load[wg]);
    }
  };

  int tile = blockIdx.x * 2 + wg;
  if (tile < n_tiles) issue_load(tile, 0);
  for (; tile < n_tiles; tile += stride) {
    const int base = tile * 128;
    const int row = base + t;
    const bool valid = row < m;
    for (int kc = 0; kc < k_chunks; ++kc) {
      tc::mbar_wait(&s_bar_load[wg], ph_load);
      ph_load ^= 1;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = lds4(s_in + t * IN_LD + g * 16 + q * 4);
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
      tc::wg_barrier(bar_id, 128);  // all rows consumed: staging buffer and A operand are ready
      if (t == 0) {
        tc::fence_after_sync();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t koff = (uint32_t)((kc * 64 + 8 * j) / 4) * 128;
          const uint64_t bh = tc::smem_desc_kmajor(bhi_addr + koff, 128, sbo);
          const uint64_t bl = tc::smem_desc_kmajor(blo_addr + koff, 128, sbo);
          tc::mma_tf32_ts(d_acc, a_hi + j * 8, bh, idesc, (kc > 0 || j > 0) ? 1u : 0u);
          tc::mma_tf32_ts(d_acc, a_lo + j * 8, bh, idesc, 1u);
          tc::mma_tf32_ts(d_acc, a_hi + j * 8, bl, idesc, 1u);
        }
        tc::mma_commit(&s_bar_mma[wg]);
      }
      {  // next work item's rows start flowing now (overlaps the MMAs and the epilogue)
        const bool more_k = kc + 1 < k_chunks;
        const int nt = more_k ? tile : tile + stride;
        if (nt < n_tiles) issue_load(nt, more_k ? (kc + 1) * 64 : 0);
      }
      tc::mbar_wait(&s_bar_mma[wg], ph_mma);
      ph_mma ^= 1;
      tc::fence_after_sync();
    }

    // epilogue: my accumulator row, 32 columns at a time, through my own staging row
    const int orow = valid ? (y_rows != nullptr ? __ldg(y_rows + row) : row) : 0;
#pragma unroll 1
    for (int c = 0; c < NT; c += 32) {
      float* orow_smem = s_out + out_buf * OUT_FLOATS + t * OUT_LD;
      tc::bulk_wait_read<1>();  // the copy that last read this buffer (two chunks ago) is done
      uint32_t v[16], w[16];
      tc::tmem_ld16(d_acc + lane_sel + c, v);
      tc::tmem_ld16(d_acc + lane_sel + c + 16, w);
      tc::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t* src = q < 4 ? &v[q * 4] : &w[(q - 4) * 4];
        float4 o = make_float4(__uint_as_float(src[0]), __uint_as_float(src[1]), __uint_as_float(src[2]),
                               __uint_as_float(src[3]));
        o = o + lds4(s_bias + c + q * 4);
        if (residual != nullptr && valid)
          o = o + *reinterpret_cast<const float4*>(residual + (size_t)orow * n_out + col_base + c + q * 4);
        sts4(orow_smem + q * 4, o);
      }
      tc::fence_async_smem();
      if (valid) tc::bulk_store(y + (size_t)orow * n_out + col_base + c, orow_smem, 128u);
      tc::bulk_commit();
      out_buf ^= 1;
    }
    tc::fence_before_sync();  // accumulator reads ordered before the next tile's MMAs
  }

  tc::bulk_wait_all<0>();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

template <int NT>
int launch_linear_tma(const float* x, const int32_t* x_rows, int m, int k, const float* wt, const float* bias,
                      const float* residual, const int32_t* y_rows, int n_out, float* y, cudaStream_t stream) {
  const int smem = 2 * NT * k * 4 + 2 * WG_FLOATS * 4;
  static int max_smem_set = 0;
  if (smem > max_smem_set) {
    CHG_CUDA(cudaFuncSetAttribute(linear_tma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    max_smem_set = smem;
  }
  const int n_tiles = (m + 127) / 128;
  const int col_tiles = n_out / NT;
  const int per_col = max(1, sm_count() / col_tiles);
  dim3 grid(min((n_tiles + 1) / 2, per_col), col_tiles);
  linear_tma_kernel<NT><<<grid, NTHR, smem, stream>>>(x, x_rows, m, k, wt, bias, residual, y_rows, n_out, y);
  CHG_LAUNCH_END();
}

}  // namespace

int linear_tma(const float* x, const int32_t* x_rows, int m, int k, const float* wt, const float* bias,
               const float* residual, const int32_t* y_rows, int n_out, float* y, cudaStream_t stream) {
  if (n_out % 128 == 0 && k == 64)
    return launch_linear_tma<128>(x, x_rows, m, k, wt, bias, residual, y_rows, n_out, y, stream);
  return launch_linear_tma<64>(x, x_rows, m, k, wt, bias, residual, y_rows, n_out, y, stream);
}

}  // namespace chg
