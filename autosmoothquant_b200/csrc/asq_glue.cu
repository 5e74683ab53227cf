// asq_glue.cu — producer-side fusions around the W8A8 linear.
//
// The reference folds 1/input_scale into the norm weight (models/llama.py:27-37, 326-339) so the
// per-tensor Linear only rounds and saturates its input (linear.py:95); its dead `LayerNormQ` /
// `dq_add_layernorm_q` (layers/nn/fused.py:2-25, csrc/kernels/fused.cu:5-24) show the intent: let the
// producer of an activation emit the int8 tensor the GEMM consumes.  These kernels do that for the
// Llama block, so three of the four GEMM launches of a layer need no phase 1 at all:
//
//   asq_add_rmsnorm_quant   x (+= delta) ; h = w * T(x * rsqrt(mean(x^2)+eps)) ; q = sat(rint(h))   -> qkv / gate|up
//   asq_silu_mul_quant      a = T(T(silu(g)) * u) ; q = sat(rint(T(a / quant_scale)))                -> down_proj
//   asq_rope_inplace        HF rotate-half RoPE applied in place to the q and k blocks of the fused qkv output
//
// All are HBM-bound elementwise / row-reduction kernels: 16-byte vector accesses, one warp per row
// (row reductions by warp shuffles), grid sized to the SM count.  Every intermediate is rounded to the
// activation dtype T exactly where eager torch would round it, so the int8 they emit equals what the
// module path (norm -> Linear.forward) would have produced up to the summation order of the variance.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "../../include/asq.h"

int asq_glue_fail(int code, const char* fmt, ...);  // defined in asq_kernels.cu (shares the error buffer)
bool asq_pdl_enabled();                             // defined in asq_kernels.cu (ASQ_PDL=0 disables)

// programmatic dependent launch (see asq_ptx.cuh): trigger the successor early, wait for the predecessor's results
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// <<<grid, block, 0, stream>>> with the programmatic-serialization attribute
template <typename... KArgs, typename... Args>
cudaError_t pdl_launch(void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(static_cast<unsigned>(block), 1, 1);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = asq_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

namespace asq_glue {

template <typename T>
struct Cvt;
template <>
struct Cvt<__nv_bfloat16> {
  __device__ static __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
  __device__ static __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
  __device__ static __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
  __device__ static __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};
template <>
struct Cvt<__half> {
  __device__ static __forceinline__ float lo(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w & 0xFFFFu))); }
  __device__ static __forceinline__ float hi(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w >> 16))); }
  __device__ static __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
  __device__ static __forceinline__ uint32_t pack(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = Cvt<T>::lo(v.x); f[1] = Cvt<T>::hi(v.x); f[2] = Cvt<T>::lo(v.y); f[3] = Cvt<T>::hi(v.y);
  f[4] = Cvt<T>::lo(v.z); f[5] = Cvt<T>::hi(v.z); f[6] = Cvt<T>::lo(v.w); f[7] = Cvt<T>::hi(v.w);
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(Cvt<T>::pack(f[0], f[1]), Cvt<T>::pack(f[2], f[3]), Cvt<T>::pack(f[4], f[5]), Cvt<T>::pack(f[6], f[7]));
}
__device__ __forceinline__ uint32_t s8x4(float v0, float v1, float v2, float v3) {
  int i0, i1, i2, i3;
  asm("cvt.rni.sat.s32.f32 %0, %1;" : "=r"(i0) : "f"(v0));
  asm("cvt.rni.sat.s32.f32 %0, %1;" : "=r"(i1) : "f"(v1));
  asm("cvt.rni.sat.s32.f32 %0, %1;" : "=r"(i2) : "f"(v2));
  asm("cvt.rni.sat.s32.f32 %0, %1;" : "=r"(i3) : "f"(v3));
  uint32_t hi, out;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(i3), "r"(i2), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(i1), "r"(i0), "r"(hi));
  return out;
}

// ------------------------------------------------------------------ add + RMSNorm (+ int8)
// One 128-thread CTA per row (grid = M: a single balanced wave, ~14 CTAs resident per SM); each thread
// keeps its NV 16-byte vectors of the row in registers between the two passes; sum of squares by warp
// shuffles + a 4-entry shared array.
constexpr int NORM_THREADS = 128;
constexpr int NORM_MAXV = 8;  // vectors per thread: H <= 128 * 8 * 8 = 8192

template <typename T, int NV>
__global__ void __launch_bounds__(NORM_THREADS) add_rmsnorm_quant_kernel(const T* x, const T* __restrict__ delta,
                                                                         const T* __restrict__ weight, T* x_out,
                                                                         T* __restrict__ h_out, int8_t* __restrict__ q_out,
                                                                         int M, int H, float eps) {
  __shared__ float warp_sums[NORM_THREADS / 32];
  pdl_sync();
  const int row = blockIdx.x;
  const int nvec = H / 8;  // vectors per row
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * H);
  const uint4* dr = delta ? reinterpret_cast<const uint4*>(delta + static_cast<size_t>(row) * H) : nullptr;
  uint4 buf[NV];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = threadIdx.x + j * NORM_THREADS;
    if (v < nvec) buf[j] = xr[v];
  }
  if (dr != nullptr) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int v = threadIdx.x + j * NORM_THREADS;
      if (v < nvec) {
        float fa[8], fd[8];
        unpack8<T>(buf[j], fa);
        unpack8<T>(__ldg(dr + v), fd);
#pragma unroll
        for (int i = 0; i < 8; ++i) fa[i] = __fadd_rn(fa[i], fd[i]);  // x + delta, rounded to T by pack8
        buf[j] = pack8<T>(fa);
        reinterpret_cast<uint4*>(x_out + static_cast<size_t>(row) * H)[v] = buf[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = threadIdx.x + j * NORM_THREADS;
    if (v < nvec) {
      float f[8];
      unpack8<T>(buf[j], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(f[i], f[i], ss);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = ss;
  __syncthreads();
  ss = warp_sums[0] + warp_sums[1] + warp_sums[2] + warp_sums[3];
  const float rstd = rsqrtf(ss / static_cast<float>(H) + eps);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = threadIdx.x + j * NORM_THREADS;
    if (v < nvec) {
      float f[8], w[8];
      unpack8<T>(buf[j], f);
      unpack8<T>(__ldg(reinterpret_cast<const uint4*>(weight) + v), w);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = Cvt<T>::rnd(__fmul_rn(w[i], Cvt<T>::rnd(__fmul_rn(f[i], rstd))));  // w * T(x*rstd)
      if (h_out != nullptr) reinterpret_cast<uint4*>(h_out + static_cast<size_t>(row) * H)[v] = pack8<T>(f);
      if (q_out != nullptr)
        reinterpret_cast<uint2*>(q_out + static_cast<size_t>(row) * H)[v] =
            make_uint2(s8x4(f[0], f[1], f[2], f[3]), s8x4(f[4], f[5], f[6], f[7]));
    }
  }
}

// ------------------------------------------------------------------ SiLU(gate) * up (+ int8)
// gate_up rows are [gate(I) | up(I)] with `row_stride` elements between rows (the fused gate|up GEMM output).
// silu uses the SFU approximations (ex2.approx / rcp.approx, relative error ~2^-21) and is then rounded to
// the 8-bit-mantissa activation dtype, so it differs from an IEEE evaluation in about 1e-4 of the elements by
// one ulp of T; with expf + IEEE division the kernel is ALU-bound at twice the runtime.
template <typename T>
__device__ __forceinline__ void silu_mul_vec(const uint4& gv, const uint4& uv, float quant_scale, float inv_quant_scale,
                                             int div_mode, int8_t* q_dst, T* a_dst) {
  float g[8], u[8], a[8];
  unpack8<T>(gv, g);
  unpack8<T>(uv, u);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float s = Cvt<T>::rnd(__fdividef(g[i], 1.0f + __expf(-g[i])));  // T(silu(g))
    a[i] = Cvt<T>::rnd(__fmul_rn(s, u[i]));
  }
  if (a_dst != nullptr) *reinterpret_cast<uint4*>(a_dst) = pack8<T>(a);
  if (q_dst != nullptr) {
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      t[i] = Cvt<T>::rnd(div_mode == ASQ_DIV_RECIPROCAL ? __fmul_rn(a[i], inv_quant_scale) : __fdiv_rn(a[i], quant_scale));
    *reinterpret_cast<uint2*>(q_dst) = make_uint2(s8x4(t[0], t[1], t[2], t[3]), s8x4(t[4], t[5], t[6], t[7]));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) silu_mul_quant_kernel(const T* __restrict__ gate_up, long long row_stride, int M, int I,
                                                             float quant_scale, float inv_quant_scale, int div_mode,
                                                             int8_t* __restrict__ q_out, T* __restrict__ a_out) {
  pdl_sync();
  // 32-bit index arithmetic (the host checks M * I / 8 < 2^31): 64-bit divisions would dominate the loop
  const uint32_t vec_per_row = static_cast<uint32_t>(I) / 8u;
  const uint32_t total = static_cast<uint32_t>(M) * vec_per_row;
  const uint32_t step = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += 2 * step) {
    // two independent vectors per iteration: four 16-byte loads in flight per thread
    const uint32_t idx2 = idx + step;
    const uint32_t row = idx / vec_per_row;
    const uint32_t v = idx - row * vec_per_row;
    const T* base = gate_up + static_cast<size_t>(row) * row_stride;
    const uint4 g0 = __ldg(reinterpret_cast<const uint4*>(base) + v);
    const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(base + I) + v);
    uint32_t row2 = 0, v2 = 0;
    uint4 g1 = make_uint4(0, 0, 0, 0), u1 = g1;
    if (idx2 < total) {
      row2 = idx2 / vec_per_row;
      v2 = idx2 - row2 * vec_per_row;
      const T* base2 = gate_up + static_cast<size_t>(row2) * row_stride;
      g1 = __ldg(reinterpret_cast<const uint4*>(base2) + v2);
      u1 = __ldg(reinterpret_cast<const uint4*>(base2 + I) + v2);
    }
    silu_mul_vec<T>(g0, u0, quant_scale, inv_quant_scale, div_mode,
                    q_out ? q_out + static_cast<size_t>(row) * I + v * 8 : nullptr,
                    a_out ? a_out + static_cast<size_t>(row) * I + v * 8 : nullptr);
    if (idx2 < total)
      silu_mul_vec<T>(g1, u1, quant_scale, inv_quant_scale, div_mode,
                      q_out ? q_out + static_cast<size_t>(row2) * I + v2 * 8 : nullptr,
                      a_out ? a_out + static_cast<size_t>(row2) * I + v2 * 8 : nullptr);
  }
}

// ------------------------------------------------------------------ RoPE in place (HF rotate-half)
// qk: rows of `row_stride` elements; the first n_heads * head_dim elements of each row are rotated.
// out[d] = T(T(x[d]*cos[d]) + T(rot[d]*sin[d])), rot[d] = d < hd/2 ? -x[d+hd/2] : x[d-hd/2]; cos/sin: [S, hd] of T.
template <typename T>
__global__ void __launch_bounds__(256) rope_kernel(T* __restrict__ qk, long long row_stride, const T* __restrict__ cos_t,
                                                   const T* __restrict__ sin_t, int M, int S, int n_heads, int head_dim) {
  pdl_sync();
  // 32-bit index arithmetic (the host checks the element count < 2^31)
  const uint32_t half_vecs = static_cast<uint32_t>(head_dim) / 16u;  // 16-byte vectors in half a head
  const uint32_t per_row = static_cast<uint32_t>(n_heads) * half_vecs;
  const uint32_t total = static_cast<uint32_t>(M) * per_row;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const uint32_t row = idx / per_row;
    const uint32_t rem = idx - row * per_row;
    const uint32_t head = rem / half_vecs;
    const uint32_t v = rem - head * half_vecs;
    const uint32_t pos = row % static_cast<uint32_t>(S);
    T* p = qk + static_cast<size_t>(row) * row_stride + static_cast<size_t>(head) * head_dim;
    uint4* lo_p = reinterpret_cast<uint4*>(p) + v;
    uint4* hi_p = reinterpret_cast<uint4*>(p + head_dim / 2) + v;
    const uint4* c_lo = reinterpret_cast<const uint4*>(cos_t + static_cast<size_t>(pos) * head_dim) + v;
    const uint4* s_lo = reinterpret_cast<const uint4*>(sin_t + static_cast<size_t>(pos) * head_dim) + v;
    const uint4* c_hi = reinterpret_cast<const uint4*>(cos_t + static_cast<size_t>(pos) * head_dim + head_dim / 2) + v;
    const uint4* s_hi = reinterpret_cast<const uint4*>(sin_t + static_cast<size_t>(pos) * head_dim + head_dim / 2) + v;
    float xl[8], xh[8], cl[8], sl[8], ch[8], sh[8], ol[8], oh[8];
    unpack8<T>(*lo_p, xl);
    unpack8<T>(*hi_p, xh);
    unpack8<T>(__ldg(c_lo), cl);
    unpack8<T>(__ldg(s_lo), sl);
    unpack8<T>(__ldg(c_hi), ch);
    unpack8<T>(__ldg(s_hi), sh);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ol[i] = __fadd_rn(Cvt<T>::rnd(__fmul_rn(xl[i], cl[i])), Cvt<T>::rnd(__fmul_rn(-xh[i], sl[i])));
      oh[i] = __fadd_rn(Cvt<T>::rnd(__fmul_rn(xh[i], ch[i])), Cvt<T>::rnd(__fmul_rn(xl[i], sh[i])));
    }
    *lo_p = pack8<T>(ol);
    *hi_p = pack8<T>(oh);
  }
}

int grid_for(long long work_items, int per_block) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(sms) * 8;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace asq_glue

extern "C" {

int asq_add_rmsnorm_quant(const void* x, const void* delta, const void* weight, void* x_out, void* h_out,
                          int8_t* q_out, int dtype, int64_t M, int64_t H, float eps, void* stream) {
  using namespace asq_glue;
  if (dtype != ASQ_BF16 && dtype != ASQ_F16) return asq_glue_fail(ASQ_ERR_INVALID, "rmsnorm: dtype must be f16 or bf16");
  if (M < 0 || H <= 0 || H % 8 != 0 || H > NORM_THREADS * 8 * NORM_MAXV)
    return asq_glue_fail(ASQ_ERR_INVALID, "rmsnorm: H=%lld must be a multiple of 8 and <= %d", (long long)H, NORM_THREADS * 8 * NORM_MAXV);
  if (M == 0) return ASQ_OK;
  if (x == nullptr || weight == nullptr || (h_out == nullptr && q_out == nullptr) || (delta != nullptr && x_out == nullptr))
    return asq_glue_fail(ASQ_ERR_INVALID, "rmsnorm: null pointer argument");
  if (!aligned16(x) || !aligned16(weight) || !aligned16(delta) || !aligned16(x_out) || !aligned16(h_out) || !aligned16(q_out))
    return asq_glue_fail(ASQ_ERR_INVALID, "rmsnorm: pointers must be 16-byte aligned");
  const int grid = static_cast<int>(M);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define ASQ_NORM_LAUNCH(TT, NVV)                                                                                        \
  launch_err = pdl_launch(add_rmsnorm_quant_kernel<TT, NVV>, grid, NORM_THREADS, st, static_cast<const TT*>(x),            \
                          static_cast<const TT*>(delta), static_cast<const TT*>(weight), static_cast<TT*>(x_out),         \
                          static_cast<TT*>(h_out), q_out, (int)M, (int)H, eps)
  const int nv = static_cast<int>((H / 8 + NORM_THREADS - 1) / NORM_THREADS);
  cudaError_t launch_err = cudaSuccess;
  if (dtype == ASQ_BF16) {
    if (nv <= 2) ASQ_NORM_LAUNCH(__nv_bfloat16, 2); else if (nv <= 4) ASQ_NORM_LAUNCH(__nv_bfloat16, 4); else ASQ_NORM_LAUNCH(__nv_bfloat16, 8);
  } else {
    if (nv <= 2) ASQ_NORM_LAUNCH(__half, 2); else if (nv <= 4) ASQ_NORM_LAUNCH(__half, 4); else ASQ_NORM_LAUNCH(__half, 8);
  }
#undef ASQ_NORM_LAUNCH
  cudaError_t e = launch_err != cudaSuccess ? launch_err : cudaGetLastError();
  return e == cudaSuccess ? ASQ_OK : asq_glue_fail(ASQ_ERR_CUDA, "rmsnorm launch failed: %s", cudaGetErrorString(e));
}

int asq_silu_mul_quant(const void* gate_up, int dtype, int64_t M, int64_t I, int64_t row_stride, float quant_scale,
                       int div_mode, int8_t* q_out, void* a_out, void* stream) {
  using namespace asq_glue;
  if (dtype != ASQ_BF16 && dtype != ASQ_F16) return asq_glue_fail(ASQ_ERR_INVALID, "silu_mul: dtype must be f16 or bf16");
  if (M < 0 || I <= 0 || I % 8 != 0 || row_stride < 2 * I || row_stride % 8 != 0)
    return asq_glue_fail(ASQ_ERR_INVALID, "silu_mul: bad shape M=%lld I=%lld stride=%lld", (long long)M, (long long)I, (long long)row_stride);
  if (M == 0) return ASQ_OK;
  if (M * (I / 8) >= (1ll << 31)) return asq_glue_fail(ASQ_ERR_UNSUPPORTED, "silu_mul: more than 2^31 vectors; split the batch");
  if (gate_up == nullptr || (q_out == nullptr && a_out == nullptr)) return asq_glue_fail(ASQ_ERR_INVALID, "silu_mul: null pointer argument");
  if (!aligned16(gate_up) || !aligned16(q_out) || !aligned16(a_out)) return asq_glue_fail(ASQ_ERR_INVALID, "silu_mul: pointers must be 16-byte aligned");
  const int grid = grid_for(M * (I / 8), 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == ASQ_BF16)
    silu_mul_quant_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(gate_up), row_stride, (int)M, (int)I,
                                                               quant_scale, 1.0f / quant_scale, div_mode, q_out,
                                                               static_cast<__nv_bfloat16*>(a_out));
  else
    silu_mul_quant_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(gate_up), row_stride, (int)M, (int)I, quant_scale,
                                                        1.0f / quant_scale, div_mode, q_out, static_cast<__half*>(a_out));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ASQ_OK : asq_glue_fail(ASQ_ERR_CUDA, "silu_mul launch failed: %s", cudaGetErrorString(e));
}

int asq_rope_inplace(void* qk, int dtype, const void* cos_table, const void* sin_table, int64_t M, int64_t S,
                     int64_t row_stride, int64_t n_heads, int64_t head_dim, void* stream) {
  using namespace asq_glue;
  if (dtype != ASQ_BF16 && dtype != ASQ_F16) return asq_glue_fail(ASQ_ERR_INVALID, "rope: dtype must be f16 or bf16");
  if (M < 0 || S <= 0 || n_heads <= 0 || head_dim <= 0 || head_dim % 16 != 0 || row_stride < n_heads * head_dim || row_stride % 8 != 0)
    return asq_glue_fail(ASQ_ERR_INVALID, "rope: bad shape");
  if (M == 0) return ASQ_OK;
  if (M * n_heads * (head_dim / 16) >= (1ll << 31)) return asq_glue_fail(ASQ_ERR_UNSUPPORTED, "rope: more than 2^31 vectors; split the batch");
  if (qk == nullptr || cos_table == nullptr || sin_table == nullptr) return asq_glue_fail(ASQ_ERR_INVALID, "rope: null pointer argument");
  if (!aligned16(qk) || !aligned16(cos_table) || !aligned16(sin_table)) return asq_glue_fail(ASQ_ERR_INVALID, "rope: pointers must be 16-byte aligned");
  const int grid = grid_for(M * n_heads * (head_dim / 16), 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == ASQ_BF16)
    rope_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<__nv_bfloat16*>(qk), row_stride, static_cast<const __nv_bfloat16*>(cos_table),
                                                     static_cast<const __nv_bfloat16*>(sin_table), (int)M, (int)S, (int)n_heads, (int)head_dim);
  else
    rope_kernel<__half><<<grid, 256, 0, st>>>(static_cast<__half*>(qk), row_stride, static_cast<const __half*>(cos_table),
                                              static_cast<const __half*>(sin_table), (int)M, (int)S, (int)n_heads, (int)head_dim);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ASQ_OK : asq_glue_fail(ASQ_ERR_CUDA, "rope launch failed: %s", cudaGetErrorString(e));
}

}  // extern "C"
