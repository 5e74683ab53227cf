// asq_nvls_probe.cu — measures what the NVSwitch data path (NVLS) sustains for the access pattern of the fused
// GEMM + all-reduce kernel: 16-byte multimem.ld_reduce (in-switch sum of all ranks' copies) and 16-byte multimem.st
// (one store lands in every rank's copy).  Not on the product path: it provides the measured ceiling the
// collective half of that kernel is reported against (profiles/r02_allreduce.md) and sized its in-flight depth.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "../../include/asq.h"

int asq_glue_fail(int code, const char* fmt, ...);

namespace {

__device__ __forceinline__ uint4 ld_reduce16(const void* a) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st16(void* a, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// mode 0: dst[i] = sum over ranks of src[i] (ld_reduce + multimem.st), 1: ld_reduce only, 2: multimem.st only.
// Every thread keeps U 16-byte requests in flight; consecutive threads touch consecutive 16-byte pieces.
template <int U>
__global__ void __launch_bounds__(1024) nvls_probe_kernel(const uint8_t* __restrict__ src_mc, uint8_t* __restrict__ dst_mc,
                                                         size_t chunks, int mode, uint32_t* sink) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  uint32_t acc = 0;
  for (size_t base = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; base < chunks; base += stride * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t c = base + u * stride;
      v[u] = make_uint4(threadIdx.x, u, 0x3f803f80u, 0x3f803f80u);
      if (mode != 2 && c < chunks) v[u] = ld_reduce16(src_mc + c * 16);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t c = base + u * stride;
      if (c < chunks) {
        if (mode != 1) mc_st16(dst_mc + c * 16, v[u]);
        else acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
      }
    }
  }
  if (mode == 1 && acc == 0x12345678u) *sink = acc;  // keeps the loads alive
}

}  // namespace

extern "C" int asq_nvls_probe(const void* src_mc, void* dst_mc, size_t bytes, int ctas, int threads, int unroll, int mode,
                              void* sink, void* stream) {
  if (src_mc == nullptr || dst_mc == nullptr || sink == nullptr || bytes % 16 != 0 || ctas < 1 || threads < 32 || threads > 1024 || mode < 0 ||
      mode > 2)
    return asq_glue_fail(ASQ_ERR_INVALID, "asq_nvls_probe: bad arguments");
  const size_t chunks = bytes / 16;
  auto* s = static_cast<const uint8_t*>(src_mc);
  auto* d = static_cast<uint8_t*>(dst_mc);
  auto* k = static_cast<uint32_t*>(sink);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (unroll) {
    case 1: nvls_probe_kernel<1><<<ctas, threads, 0, st>>>(s, d, chunks, mode, k); break;
    case 2: nvls_probe_kernel<2><<<ctas, threads, 0, st>>>(s, d, chunks, mode, k); break;
    case 4: nvls_probe_kernel<4><<<ctas, threads, 0, st>>>(s, d, chunks, mode, k); break;
    case 8: nvls_probe_kernel<8><<<ctas, threads, 0, st>>>(s, d, chunks, mode, k); break;
    case 16: nvls_probe_kernel<16><<<ctas, threads, 0, st>>>(s, d, chunks, mode, k); break;
    default: return asq_glue_fail(ASQ_ERR_INVALID, "asq_nvls_probe: unroll must be 1, 2, 4, 8 or 16");
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return asq_glue_fail(ASQ_ERR_CUDA, "asq_nvls_probe launch failed: %s", cudaGetErrorString(e));
  return ASQ_OK;
}
