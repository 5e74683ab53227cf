// asq_smallm.cu — decode-sized W8A8 linear (M <= 16 token rows): an HBM-bound weight stream.
//
// With a handful of token rows the [N,K] int8 weight is read once and everything else is noise, so the job is
// to keep every SM streaming weight bytes.  The tcgen05 kernel of asq_kernels.cu hands out 256-column tiles:
// N/256 CTAs (16 for a 4096-wide projection) each pull a 1 MB slab through a 3-4 stage ring, which is latency-
// bound at ~0.75 TB/s.  Here a CTA owns 16 output columns, its warps split them as 2 column groups x 4 (8 for
// K >= 8192) K parts; a warp walks its K part in 64-byte steps: every lane loads 16 contiguous bytes of one weight row
// (8 rows x 64 bytes per warp instruction, streamed past L1) and 2 x 16 bytes of the quantised activations
// (L1 / L2 resident, [16, K] int8), and issues two mma.sync.m16n8k32.s8 — the K order inside a 64-byte step is
// permuted identically for both operands, which an integer dot product does not notice.  int32 partials of the
// four K quarters are summed through shared memory (exact), then the same fp32 dequant arithmetic as the big
// kernel runs (factor first, * acc, + bias, each rounded separately), so outputs are bit-identical to it.
//
// The activation is quantised by the stand-alone prologue kernel first (a second small launch on the stream).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "../../include/asq.h"
#include "asq_smallm.h"

int asq_glue_fail(int code, const char* fmt, ...);  // defined in asq_kernels.cu (shares the error buffer)

namespace asq_smallm {

constexpr int COLS_PER_CTA = 16;  // two 8-column mma tiles per CTA
constexpr int UNROLL = 8;         // 64-byte K steps of W in flight per lane (128 bytes per lane, 4 KB per warp)

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void mma_s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void store_out(void* y, int y_dtype, size_t idx, float v) {
  if (y_dtype == ASQ_BF16) reinterpret_cast<__nv_bfloat16*>(y)[idx] = __float2bfloat16_rn(v);
  else if (y_dtype == ASQ_F16) reinterpret_cast<__half*>(y)[idx] = __float2half_rn(v);
  else reinterpret_cast<float*>(y)[idx] = v;
}

// KS = K parts per column group (warps per CTA = 2 * KS); MB = 16-row blocks of activations (M <= 16 * MB).
template <int KS, int MB>
__global__ void __launch_bounds__(2 * KS * 32) smallm_kernel(const SmallMParams p) {
  __shared__ int partial[2 * KS][MB][4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;  // mma fragment coordinates: row / column group, position inside a k group
  const int nt = warp / KS, kq = warp % KS;
  const int col_base = blockIdx.x * COLS_PER_CTA + nt * 8;
  const int steps_total = p.K / 64;  // 64-byte K steps; part kq takes [s0, s1)
  const int s0 = static_cast<int>(static_cast<long long>(kq) * steps_total / KS);
  const int s1 = static_cast<int>(static_cast<long long>(kq + 1) * steps_total / KS);
  const int8_t* wrow = p.w + static_cast<size_t>(col_base + g) * p.K + t * 16;
  // rows g and g + 8 of every 16-row block; rows >= M read row 0 (their results are never stored)
  const int8_t* a_lo[MB];
  const int8_t* a_hi[MB];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
    const int r0 = mb * 16 + g, r1 = r0 + 8;
    a_lo[mb] = p.xq + static_cast<size_t>(r0 < p.M ? r0 : 0) * p.K + t * 16;
    a_hi[mb] = p.xq + static_cast<size_t>(r1 < p.M ? r1 : 0) * p.K + t * 16;
  }
  int c[MB][4];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[mb][i] = 0;
  int s = s0;
  for (; s + UNROLL <= s1; s += UNROLL) {
    uint4 b[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) b[u] = ldg_stream(wrow + static_cast<size_t>(s + u) * 64);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(a_lo[mb] + static_cast<size_t>(s + u) * 64));
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(a_hi[mb] + static_cast<size_t>(s + u) * 64));
        mma_s8(c[mb], lo.x, hi.x, lo.y, hi.y, b[u].x, b[u].y);
        mma_s8(c[mb], lo.z, hi.z, lo.w, hi.w, b[u].z, b[u].w);
      }
    }
  }
  for (; s < s1; ++s) {
    const uint4 b = ldg_stream(wrow + static_cast<size_t>(s) * 64);
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      const uint4 lo = __ldg(reinterpret_cast<const uint4*>(a_lo[mb] + static_cast<size_t>(s) * 64));
      const uint4 hi = __ldg(reinterpret_cast<const uint4*>(a_hi[mb] + static_cast<size_t>(s) * 64));
      mma_s8(c[mb], lo.x, hi.x, lo.y, hi.y, b.x, b.y);
      mma_s8(c[mb], lo.z, hi.z, lo.w, hi.w, b.z, b.w);
    }
  }
#pragma unroll
  for (int mb = 0; mb < MB; ++mb)
#pragma unroll
    for (int i = 0; i < 4; ++i) partial[warp][mb][i][lane] = c[mb][i];
  __syncthreads();
  if (kq != 0) return;
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int acc = c[mb][i];
      for (int q = 1; q < KS; ++q) acc += partial[warp + q][mb][i][lane];  // exact: int32 partial sums
      // accumulator fragment: c0, c1 = row g, columns 2t, 2t+1; c2, c3 = row g + 8, same columns
      const int row = mb * 16 + g + (i >> 1) * 8, col = col_base + 2 * t + (i & 1);
      if (row >= p.M) continue;
      float f = (p.col_scale != nullptr) ? __ldg(p.col_scale + col) : p.dequant_scale;
      if (p.row_scale != nullptr) f = __fmul_rn(f, __ldg(p.row_scale + row));
      float v = __fmul_rn(f, __int2float_rn(acc));
      if (p.bias != nullptr) v = __fadd_rn(v, __ldg(p.bias + col));
      store_out(p.y, p.y_dtype, static_cast<size_t>(row) * p.N + col, v);
    }
  }
}

template <int KS, int MB>
cudaError_t launch(const SmallMParams& p, cudaStream_t stream) {
  smallm_kernel<KS, MB><<<p.N / COLS_PER_CTA, 2 * KS * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace asq_smallm

bool asq_smallm_supported(int64_t M, int64_t N, int64_t K, int y_dtype) {
  // measured: with 2 or 4 row blocks per warp (M up to 64) the activation re-reads through L1 make it slower than the
  // tcgen05 kernel (28.9 vs 22.9 us at M = 64, 4096x4096), so only M <= 16 comes here
  return M >= 1 && M <= 16 && N % 16 == 0 && K % 64 == 0 && (y_dtype == ASQ_BF16 || y_dtype == ASQ_F16 || y_dtype == ASQ_F32);
}

int asq_smallm_launch(const SmallMParams& p, cudaStream_t stream) {
  using namespace asq_smallm;
  const bool long_k = p.K >= 8192;  // more K parts keep enough weight bytes in flight when N / 16 CTAs are few
  cudaError_t e;
  e = long_k ? launch<8, 1>(p, stream) : launch<4, 1>(p, stream);
  return e == cudaSuccess ? ASQ_OK : asq_glue_fail(ASQ_ERR_CUDA, "small-M kernel launch failed: %s", cudaGetErrorString(e));
}
