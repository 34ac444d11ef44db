// Segmented gather-reduce:  out[r] (+)= sum_{k in [ptr[r], ptr[r+1])} data[perm ? perm[k] : k]
//
// This is the atomics-free replacement of the reference's `aggregate`
// (chgnet/model/functions.py:25-37: zeros().index_add_(0, owners, data), an atomicAdd
// scatter on CUDA) for every scatter in the model: AtomConv messages -> atoms
// (layers.py:124-126; edges are center-sorted so perm == NULL and each segment is one
// contiguous HBM stream), BondConv updates -> bonds (layers.py:252-254), and the
// transposed gathers of the reverse pass (perm = the batch's neighbour / bond-j /
// atom groupings).  Deterministic: fixed summation order.
//
// A group of W/4 lanes (16 for 64-wide rows, 32 for 128-wide) owns one output row;
// each lane carries a float4 column slice; rows of the segment are streamed four at a
// time so a lane keeps four independent 16-byte loads in flight.  HBM-bound: the
// algorithmic bytes are (4*W + 4 [+4 if perm]) per input row + 4*W per output row.
#include "common.cuh"

namespace chg {
namespace {

template <int W>
__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ data, const int32_t* __restrict__ perm,
                   const int32_t* __restrict__ ptr, int n_rows, int accumulate, float* out, int out_ld) {
  constexpr int LANES = W / 4;            // lanes per output row
  constexpr int GROUPS = 256 / LANES;     // output rows per CTA pass
  const int sub = threadIdx.x % LANES;
  const int grp = threadIdx.x / LANES;
  for (int r = blockIdx.x * GROUPS + grp; r < n_rows; r += gridDim.x * GROUPS) {
    const int beg = ptr[r], end = ptr[r + 1];
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    int k = beg;
    if (perm == nullptr) {
      const float* base = data + (size_t)sub * 4;
      for (; k + 4 <= end; k += 4) {
        const float4 v0 = ldg4(base + (size_t)(k + 0) * W);
        const float4 v1 = ldg4(base + (size_t)(k + 1) * W);
        const float4 v2 = ldg4(base + (size_t)(k + 2) * W);
        const float4 v3 = ldg4(base + (size_t)(k + 3) * W);
        a0 = a0 + v0; a1 = a1 + v1; a2 = a2 + v2; a3 = a3 + v3;
      }
      for (; k < end; ++k) a0 = a0 + ldg4(base + (size_t)k * W);
    } else {
      const float* base = data + (size_t)sub * 4;
      for (; k + 4 <= end; k += 4) {
        const int i0 = perm[k], i1 = perm[k + 1], i2 = perm[k + 2], i3 = perm[k + 3];
        const float4 v0 = ldg4(base + (size_t)i0 * W);
        const float4 v1 = ldg4(base + (size_t)i1 * W);
        const float4 v2 = ldg4(base + (size_t)i2 * W);
        const float4 v3 = ldg4(base + (size_t)i3 * W);
        a0 = a0 + v0; a1 = a1 + v1; a2 = a2 + v2; a3 = a3 + v3;
      }
      for (; k < end; ++k) a0 = a0 + ldg4(base + (size_t)perm[k] * W);
    }
    // fixed combination order -> bitwise reproducible
    float4 s = (a0 + a1) + (a2 + a3);
    float* dst = out + (size_t)r * out_ld + sub * 4;
    if (accumulate) s = s + *reinterpret_cast<const float4*>(dst);
    stg4(dst, s);
  }
}

}  // namespace
}  // namespace chg

using namespace chg;

extern "C" int chg_segment_sum(const float* data, int32_t width, const int32_t* perm, const int32_t* ptr,
                               int32_t n_rows, int32_t accumulate, float* out, int32_t out_ld, void* stream) {
  CHG_CHECK_ARG(n_rows >= 0, "negative size");
  CHG_CHECK_ARG(width == 64 || width == 128, "width must be 64 or 128");
  CHG_CHECK_ARG(out_ld >= width && out_ld % 4 == 0, "out_ld must be >= width and a multiple of 4");
  if (n_rows == 0) return CHG_OK;
  CHG_CHECK_ARG(ptr && out, "null pointer");
  const int groups = width == 64 ? 16 : 8;
  const int blocks = max(1, min((n_rows + groups - 1) / groups, sm_count() * 8));
  if (width == 64)
    segment_sum_kernel<64><<<blocks, 256, 0, as_stream(stream)>>>(data, perm, ptr, n_rows, accumulate, out, out_ld);
  else
    segment_sum_kernel<128><<<blocks, 256, 0, as_stream(stream)>>>(data, perm, ptr, n_rows, accumulate, out, out_ld);
  CHG_LAUNCH_END();
}
