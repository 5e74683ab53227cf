// asq_ceiling.cu — tcgen05 issue-rate microbenchmark: the 8-bit tensor-pipe ceiling the rooflines are held against.
//
// SURVEY 8(d) asks for the INT8 roofline denominator to be MEASURED ("in-register tcgen05 kind::i8 issue-rate
// microbench") instead of assumed to be 2 x bf16.  Every SM (or SM pair) runs the main loop of asq_linear_kernel
// with everything but the tensor pipe removed: operands are RESIDENT in shared memory (filled once, no TMA, no
// HBM / L2 traffic), one thread issues 128 x 256 x 32 (cta_group::1) or 256 x 256 x 32 (cta_group::2) MMAs
// back to back into two alternating TMEM accumulators, tcgen05.commit every 16 k-blocks with at most two
// groups in flight, no epilogue.  What it reports is therefore the rate at which the tensor pipe retires
// kind::i8 / kind::f8f6f4 MMAs of the kernel's own shape at the clock the chip sustains under that load.
//
// Measurement utility: its own small library (libasq_ceiling.so), not part of the drop-in ABI of include/asq.h.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

#include "asq_ptx.cuh"

namespace {

using namespace asq;

constexpr int kThreads = 128;
constexpr int kBlockM = 128, kTileN = 256, kBlockK = 128, kUmmaK = 32;
constexpr int kKbPerGroup = 16;                 // k-blocks (x 4 MMAs) between commits
constexpr uint32_t kABytes = kBlockM * kBlockK;  // 16 KB, as one pipeline stage of the real kernel
constexpr uint32_t kSmemBytes = 200 * 1024;      // far more than needed: forces one CTA per SM
constexpr uint32_t kSpinLimit = 1u << 27;

// bounded wait: a protocol mistake must end in an error code, not in a hung GPU
__device__ __forceinline__ bool spin_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < kSpinLimit; ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// out[4 * worker + {0,1,2,3}] = SM cycles spent issuing, ns spent issuing, MMAs issued, status (1 = ok, 2 = timeout)
template <bool FP8, int CG>
__global__ void __launch_bounds__(kThreads, 1) asq_mma_ceiling_kernel(int groups, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  constexpr uint32_t kBBytes = (kTileN / CG) * kBlockK;  // this CTA's share of the W tile
  const uint32_t sA = base, sB = base + kABytes;
  const uint32_t bar0 = base + kABytes + kBBytes;
  const uint32_t tmem_slot = bar0 + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kABytes + kBBytes + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG > 1) ? cluster_ctarank() : 0u;

  // operands: pseudo-random bytes (data-dependent switching power is part of what sets the sustained clock);
  // bit 0 cleared so that no e4m3 byte is a NaN code
  for (uint32_t i = threadIdx.x; i < (kABytes + kBBytes) / 4; i += kThreads) {
    uint32_t v = (i + 1u) * 2654435761u + blockIdx.x * 40503u;
    v ^= v >> 15; v *= 2246822519u; v ^= v >> 13;
    reinterpret_cast<uint32_t*>(base_ptr)[i] = v & 0xFEFEFEFEu;
  }
  fence_proxy_async_smem();  // generic-proxy stores -> tensor-core (async proxy) reads
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
    else         { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 1 && lane == 0 && cta_rank == 0) {
    constexpr uint32_t idesc = make_idesc(FP8, kBlockM * CG, kTileN);
    const uint64_t adesc = make_smem_desc_sw128(sA);
    const uint64_t bdesc = make_smem_desc_sw128(sB);
    bool ok = true;
    unsigned long long mmas = 0;
    const unsigned long long ns0 = globaltimer_ns();
    const long long c0 = clock64();
    for (int g = 0; g < groups && ok; ++g) {
      const uint32_t bar = bar0 + 8u * (g & 1);
      if (g >= 2) ok = spin_wait(bar, ((g >> 1) - 1) & 1u);  // group g-2 retired: at most two groups in flight
      if (!ok) break;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (g & 1) * kTileN;
      for (int kb = 0; kb < kKbPerGroup; ++kb) {
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          const uint64_t koff = static_cast<uint64_t>(k * (kUmmaK >> 4));  // +32 bytes along K inside the swizzle row
          const uint32_t accum = (kb | k) != 0;
          if (CG == 2) {
            if (FP8) mma_f8_pair(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
            else     mma_i8_pair(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
          } else {
            if (FP8) mma_f8(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
            else     mma_i8(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
          }
        }
      }
      mmas += kKbPerGroup * (kBlockK / kUmmaK);
      if (CG == 2) mma_commit_pair(bar, static_cast<uint16_t>(1u));  // arrive on the leader's barrier only
      else         mma_commit(bar);
    }
    // drain: the last two groups
    for (int g = (groups >= 2 ? groups - 2 : 0); g < groups && ok; ++g) ok = spin_wait(bar0 + 8u * (g & 1), (g >> 1) & 1u);
    const long long c1 = clock64();
    const unsigned long long ns1 = globaltimer_ns();
    unsigned long long* o = out + 4ull * (blockIdx.x / CG);
    o[0] = static_cast<unsigned long long>(c1 - c0);
    o[1] = ns1 - ns0;
    o[2] = mmas;
    o[3] = ok ? 1ull : 2ull;
  }
  tc_fence_before();
  if (CG > 1) cluster_sync_all(); else __syncthreads();  // the peer's shared memory is read until the last MMA retires
  tc_fence_after();
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_pair(tmem_base, 512);
    else         tmem_dealloc(tmem_base, 512);
  }
}

template <bool FP8, int CG>
cudaError_t launch(int ctas, int groups, unsigned long long* out, cudaStream_t stream) {
  auto kern = asq_mma_ceiling_kernel<FP8, CG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBytes));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(ctas), 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, groups, out);
}

}  // namespace

// result[0] = achieved T(FL)OP/s over the whole chip (CUDA events around `reps` launches)
// result[1] = SM cycles per MMA instruction (median worker, in-kernel clock64)
// result[2] = effective SM clock in MHz during the run (cycles / globaltimer ns, median worker)
// result[3] = milliseconds per launch        result[4] = CTAs launched        result[5] = ops per MMA instruction
// returns 0, or a negative code: -1 CUDA error (message on stderr), -2 a worker timed out, -3 bad arguments
extern "C" int asq_mma_ceiling(int fp8, int cta_group, int groups, int reps, double* result) {
  if ((cta_group != 1 && cta_group != 2) || groups < 2 || reps < 1 || result == nullptr) return -3;
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1;
  if (prop.major != 10) { fprintf(stderr, "asq_mma_ceiling: needs an sm_100 device\n"); return -1; }
  const int ctas = prop.multiProcessorCount / cta_group * cta_group;
  const int workers = ctas / cta_group;
  unsigned long long* out = nullptr;
  if (cudaMalloc(&out, sizeof(unsigned long long) * 4 * workers) != cudaSuccess) return -1;
  cudaMemset(out, 0, sizeof(unsigned long long) * 4 * workers);
  cudaStream_t stream = nullptr;
  auto go = [&]() -> cudaError_t {
    if (fp8) return cta_group == 2 ? launch<true, 2>(ctas, groups, out, stream) : launch<true, 1>(ctas, groups, out, stream);
    return cta_group == 2 ? launch<false, 2>(ctas, groups, out, stream) : launch<false, 1>(ctas, groups, out, stream);
  };
  cudaError_t e = go();  // warm-up (and clock ramp)
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  if (e == cudaSuccess) {
    cudaEventRecord(e0, stream);
    for (int r = 0; r < reps && e == cudaSuccess; ++r) e = go();
    cudaEventRecord(e1, stream);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
  }
  int rc = 0;
  if (e != cudaSuccess) {
    fprintf(stderr, "asq_mma_ceiling: %s\n", cudaGetErrorString(e));
    rc = -1;
  } else {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long* host = new unsigned long long[4 * workers];
    cudaMemcpy(host, out, sizeof(unsigned long long) * 4 * workers, cudaMemcpyDeviceToHost);
    // median worker (insertion sort by cycles: <= 148 entries)
    int* order = new int[workers];
    for (int i = 0; i < workers; ++i) order[i] = i;
    for (int i = 1; i < workers; ++i)
      for (int j = i; j > 0 && host[4 * order[j]] < host[4 * order[j - 1]]; --j) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
    const int med = order[workers / 2];
    double total_mmas = 0;
    for (int i = 0; i < workers; ++i) {
      if (host[4 * i + 3] != 1ull) rc = -2;
      total_mmas += static_cast<double>(host[4 * i + 2]);
    }
    const double ops_per_mma = 2.0 * kBlockM * cta_group * kTileN * kUmmaK;
    const double ms_per_launch = ms / reps;
    result[0] = total_mmas * ops_per_mma / (ms_per_launch * 1e-3) / 1e12;
    result[1] = host[4 * med + 2] ? static_cast<double>(host[4 * med]) / static_cast<double>(host[4 * med + 2]) : 0.0;
    result[2] = host[4 * med + 1] ? static_cast<double>(host[4 * med]) / static_cast<double>(host[4 * med + 1]) * 1e3 : 0.0;
    result[3] = ms_per_launch;
    result[4] = ctas;
    result[5] = ops_per_mma;
    delete[] host;
    delete[] order;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return rc;
}
