// asq_kernels.cu — B200 (sm_100a) SmoothQuant W8A8 / FP8 linear path: kernels + C ABI.
//
// One persistent, warp-specialised kernel per call (template: int8 | e4m3, CTAs per tile, feature set):
//
//   phase 1 (epilogue warps, all CTAs)   x[M,K] (fp32|fp16|bf16) -> 8-bit A panel in an L2-resident
//       workspace.  One warp per token row: per-token absmax with warp shuffles, true IEEE division,
//       rint, saturate.  Each finished row bumps a per-128-row-panel counter (release, gpu scope).
//   phase 2 (same launch)                TMA producer warp waits (acquire) for the panel it is about
//       to load, then streams 128x128-byte A and BNx128-byte W tiles (128B swizzle) through a
//       multi-stage mbarrier ring; one elected thread issues tcgen05.mma (kind::i8 -> int32,
//       kind::f8f6f4 -> fp32) into a double-buffered TMEM accumulator; 8 epilogue warps drain TMEM
//       with tcgen05.ld, apply the reference's fp32 dequant (+bias) arithmetic op for op, convert and
//       store through a swizzled staging tile + TMA store, overlapping the next tile's main loop.
//
// Epilogue / schedule variants of the same kernel (selected per launch, compiled per feature set — see Feat):
// per-column scale vector (fused q|k|v, gate|up), SwiGLU (+ next layer's quantisation), RoPE, residual add,
// raw int32 / alpha-beta / int8 outputs, stream-K tail, grouped (MoE experts) and batched weight selection,
// and the row-parallel variant fused with its all-reduce over NVLink peer memory.
//
// Why the A operand takes one trip through L2 instead of being converted per CTA: every N-tile CTA
// of an M panel would otherwise re-read the 16/32-bit activations and redo the division (N/BN times
// the ALU work and 2-4x the shared-memory bytes per MMA step, which is already the limiter for 8-bit
// operands).  Quantising once and letting TMA feed 8-bit tiles keeps the tensor pipe's operand
// traffic at the minimum, and the panel counters let the GEMM start as soon as its first panel
// exists, so no separate quantise / dequantise launch and no grid-wide barrier runs.
//
// Reference semantics (AniZpZ/AutoSmoothQuant): autosmoothquant/layers/nn/linear.py:83-106,
// 172-208, 278-302 (INT8), :336-369, 413-427, 551-566 (FP8), functional/quantization.py:144-211,
// csrc/int8gemm/cublasINT8MMWrapper.cc:224-354 (C = A . W^T, int32).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/asq.h"
#include "asq_ptx.cuh"
#include "asq_smallm.h"

namespace asq {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 128;  // bytes == elements for 8-bit operands; one 128B swizzle row
constexpr int UMMA_K = 32;    // K per tcgen05.mma for 8-bit operands
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;  // warp0 TMA, warp1 MMA, warps 2..9 epilogue
constexpr int SYNC_AMAX = 8190;      // per-tensor dynamic: bit pattern of the running absmax (non-negative floats order like uints)
constexpr int SYNC_AMAX_CNT = 8191;  // ... and the number of warps that contributed
constexpr int SYNC_EXIT = 0;                          // sync[0]: CTAs that finished; sync[1+p]: rows ready in panel p

enum EpiKind : int { EPI_DEQUANT = 0, EPI_RAW_I32 = 1, EPI_ALPHA_BETA = 2, EPI_SWIGLU = 3 };

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// timeline slots (per CTA): 0 start, 1 setup done, 2 phase-1 done (warp 2), 3 first panel acquired,
// 4 first full barrier passed (MMA), 5 first accumulator ready (epilogue), 6 last tile stored, 7 exit
#define ASQ_STAMP(slot) do { if (p.dbg != nullptr) p.dbg[blockIdx.x * 8 + (slot)] = globaltimer_ns(); } while (0)

struct LinearParams {
  // phase 1
  const void* x;         // nullptr: A is already 8-bit (tmA points at the caller's matrix)
  uint8_t* a_q;          // [M,K] 8-bit workspace written by phase 1
  float* row_scale;      // [M] workspace (per-token scales)
  float* row_scale_out;  // optional copy for the caller
  const float* row_scale_in;  // ASQ_ACT_ROW_SCALE_GIVEN: per-token scales supplied by the caller
  uint32_t* sync;        // phase counters, zero on entry, restored to zero on exit
  float quant_scale, inv_quant_scale, inv_qmax, qmax;
  // phase 2
  void* y;
  const float* bias;
  const float* col_scale;
  const void* bias_any;  // EPI_ALPHA_BETA bias (int8 / int32 / fp32)
  float dequant_scale, alpha, beta;
  float out_fq_scale, inv_out_fq_scale;  // != 0: fake-quantise the output through e4m3 (FP8LinearStatic, linear.py:562-564)
  // EPI_SWIGLU: w rows interleave 32 gate rows with the 32 matching up rows; the epilogue emits
  // a = T(T(silu(gate)) * up) as T (y_dtype 16-bit) or sat(rint(T(a / out_quant_scale))) as int8, [M, N/2]
  float out_quant_scale, inv_out_quant_scale;
  // ASQ_ACT_RMSNORM: phase 1 is HF RMSNorm (+ the folded 1/input_scale) followed by round-only quantisation:
  // q = sat(rint(T(norm_weight * T(x * rsqrt(mean(x^2) + eps))))), the arithmetic of asq_add_rmsnorm_quant
  const void* norm_weight;  // [K] of x's dtype
  float norm_eps;
  // y = T(residual + T(linear output)): the residual-stream add of the decoder block (HF LlamaDecoderLayer:
  // hidden = residual + o_proj(...) / + down_proj(...)) in the epilogue, so the following norm kernel reads one
  // tensor instead of two and writes no copy of the stream.  [M, N] of y's 16-bit dtype, may alias y.
  const void* residual;
  // RoPE in the epilogue (fused q|k|v projection): columns < rope_cols are rotated per 128-wide head with the
  // HF rotate-half formula, position = row % rope_S, tables [rope_S, 128] of the output dtype
  const void* rope_cos;
  const void* rope_sin;
  int rope_S, rope_cols, rope_halves_equal;
  float dequant_scale_up;  // scalar dequant scale of the up columns (gate uses dequant_scale) when col_scale == NULL
  int mid_dtype;  // T: the activation dtype gate / up / a are rounded to (ASQ_BF16 | ASQ_F16)
  int M, N, K;
  int x_dtype, y_dtype, bias_dtype, act_mode, div_mode, epi_kind, flags;
  int num_m_blocks, num_n_blocks, num_k_blocks, group, raster_m;  // num_n_blocks = TILE_N-wide tiles per row
  int tile_m_blocks;                                              // 128-row blocks per M tile (1, or 2 for CTA pairs)
  int n_units, tile_units, rounds, tail_tiles;                    // schedule (see TileWalk)
  // stream-K tail: the k-iterations of the left-over tiles are dealt evenly to all workers; partial
  // accumulators travel through `sk_partial` (one 128-row x 256-column fp32/int32 slot per CTA), handshake
  // words in `sk_flags` (one per epilogue warp, set by the contributor, cleared by the owner)
  int sk_enabled, sk_total, sk_workers;  // sk_workers <= launched workers: every participant gets >= 1 iteration
  uint32_t* sk_partial;
  uint32_t* sk_flags;
  // Grouped GEMM (MoE experts): the rows of x are sorted by group and every group's segment is padded to a
  // multiple of 256 rows; group_of_blk[row / 128] names the group (expert) of a 128-row block, -1 = padding block
  // (its tiles are skipped).  Group g multiplies by weight rows [g * N, (g + 1) * N) of the stacked [G * N, K]
  // matrix and dequantises with group_scale[g] (SwiGLU: up columns with group_scale_up[g]); the per-tensor
  // activation scale of phase 1 is group_quant_scale[g] when given.
  const int* group_of_blk;
  const float* group_scale;
  const float* group_scale_up;
  const float* group_quant_scale;
  int num_groups;
  int batch_rows;  // batched GEMM: group = row / batch_rows (no table); rows per batch, a multiple of the tile height
  // Row-parallel GEMM fused with its all-reduce over peer memory (NVLink P2P): every rank computes the partial
  // of every tile over its K shard; tile t is OWNED by rank t % world.  Non-owners push their raw int32 / fp32
  // accumulators into the owner's receive buffer and raise a flag; the owner adds them to its own accumulator
  // (exact for int32, rank order for fp32), runs the normal epilogue and TMA-stores the finished tile into the
  // y buffer of EVERY rank.  Each rank walks the tiles it does not own first, so partials arrive before their
  // owner needs them.  ar_ctl[r] / ar_recv[r] are rank r's control words and receive buffer mapped into this
  // process; words: [0] epoch of the last finished launch, [1] CTAs finished, [2+s] "rank s finished launch e",
  // [64...] one arrival flag per (source slot, owned tile, CTA of the pair, epilogue warp).
  int ar_world, ar_rank, ar_tiles, ar_cnt_max;
  int ar_partial16;    // 1: partials travel dequantised in the 16-bit output dtype (NCCL-native numerics, half the bytes)
  uint8_t* ar_y_mc;    // != NULL: NVLS multicast address of the y buffers; finished tiles leave by multimem.st (one store
                       // reaches every rank through the switch) instead of one TMA store per rank
  uint32_t* ar_ctl[8];
  uint32_t* ar_recv[8];
  // In-switch all-reduce (F_NVLS): every rank TMA-stores its 16-bit partial tiles into its OWN symmetric buffer (p.y)
  // and bumps a counter on the tile's owner (rank tile % world, ar_ctl[owner]).  A few reducer CTAs appended to the
  // grid sum the world copies of the owned tiles with multimem.ld_reduce on the buffers' multicast address (the
  // NVSwitch adds them) and multimem.st the sums into every rank's output.  All ranks walk the tiles in the SAME
  // order, so a tile's partials complete at about the same time everywhere and the reduction of one round of tiles
  // overlaps the MMAs of the next (see nvls_reducer_warp).
  const uint8_t* nvls_p_mc;  // multicast address of the partial buffers [M, N] 16-bit (NULL: not an NVLS launch)
  uint8_t* nvls_y_mc;        // multicast address of the outputs [M, N] 16-bit
  int nvls_reducers;         // CTAs at the end of the grid that only reduce (see nvls_reducer_warp)
  // "slab landed" counters, one per (tile, 64-row slab), in the same symmetric allocation: every rank bumps its LOCAL
  // copy (a remote atomic per tile would wait behind the reduction traffic on the link: measured ~25 us each), the
  // reducer reads the SUM over all ranks' copies with one multimem.ld_reduce.  Two banks alternate between launches;
  // every launch clears the bank the next one will use (nobody reads it any more: the previous launch's handshake
  // is complete) so no rank ever sees counts of an older launch.
  uint32_t* nvls_ctr;            // this rank's copy of the current bank
  const uint32_t* nvls_ctr_mc;   // multicast address of the current bank
  uint32_t* nvls_ctr_next;       // this rank's copy of the other bank (cleared by this launch)
  int nvls_ctr_words;            // words per bank in use by this launch's shape bound
  int tma_store;            // 1: outputs leave through shared-memory staging + TMA store (tmY is valid)
  unsigned long long* dbg;  // optional timeline buffer (8 slots per CTA), nullptr in production
};

// ------------------------------------------------------------------ element helpers
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int VEC = 4;  // elements per 16-byte load
  static constexpr bool IS_BF16 = false;
  __device__ static __forceinline__ void unpack(const uint4& v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  __device__ static __forceinline__ float round_to(float v) { return v; }
};
template <>
struct Elem<__half> {
  static constexpr int VEC = 8;
  static constexpr bool IS_BF16 = false;
  __device__ static __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 t = __half22float2(h);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
  __device__ static __forceinline__ float round_to(float v) { return __half2float(__float2half_rn(v)); }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int VEC = 8;
  static constexpr bool IS_BF16 = true;
  __device__ static __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  __device__ static __forceinline__ float round_to(float v) {
    return __bfloat162float(__float2bfloat16_rn(v));
  }
};

// sat_i8(rint(v)); NaN -> 0 (what torch's clamp + .to(int8) yields for the reference)
__device__ __forceinline__ uint32_t cvt_s8(float v) {
  int r;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(r) : "f"(v));
  return static_cast<uint32_t>(r) & 0xFFu;
}
// two floats -> two e4m3 bytes (round-to-nearest-even, saturate to +-448, NaN stays NaN):
// identical to clamp(+-448) followed by torch's .to(float8_e4m3fn) (quantization.py:187-190)
__device__ __forceinline__ uint32_t cvt_e4m3x2(float lo, float hi) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
  return r;
}

// four floats -> four saturated int8 in one word: sat_i8(rint(v)), NaN -> 0.  F2I (round-to-nearest-even,
// saturating to int32) then two I2IP packs that saturate to int8.
__device__ __forceinline__ uint32_t cvt_s8x4(float v0, float v1, float v2, float v3) {
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

template <bool FP8, int VEC>
__device__ __forceinline__ void pack_store(uint8_t* dst, const float (&q)[VEC]) {
  uint32_t w[VEC / 4];
#pragma unroll
  for (int i = 0; i < VEC / 4; ++i) {
    if (FP8) {
      w[i] = cvt_e4m3x2(q[4 * i], q[4 * i + 1]) | (cvt_e4m3x2(q[4 * i + 2], q[4 * i + 3]) << 16);
    } else {
      w[i] = cvt_s8x4(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
    }
  }
  if (VEC == 8) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[VEC / 4 - 1]);
  } else {
    *reinterpret_cast<uint32_t*>(dst) = w[0];
  }
}

// Quantise one token row with one warp.  Returns the row scale (per-token) or 0.
//   ROUND       q = sat(rint(x))                                    linear.py:95
//   SCALE       q = sat(rint(T(x / qs)))   division rounded to T    linear.py:290-292 / quantization.py:208-211
//   PER_TOKEN   s = f32(T(absmax) / T(qmax)); q = sat(rint(f32(x)/s))   linear.py:88-92 / quantization.py:183-189
// fp8: no rint, q = e4m3(clamp(v)).
// Loads are issued in batches of QBATCH 16-byte vectors per lane (QBATCH*512 bytes per warp in
// flight) so the phase is bandwidth- rather than latency-bound; a row that fits one batch
// (K <= 4096 for 16-bit, 2048 for fp32 inputs) is read from HBM exactly once even per-token.
constexpr int QBATCH = 16;

// rint(x / s) without a division for the int8 per-token path.  r = x * fl(1/s) is within 2 ulp of the real
// quotient, so rint(r) equals rint(fl(x / s)) unless r lies within 2^-14 of a half-integer (|r| <= ~128
// here, ulp(r) <= 2^-17); only then the IEEE division is evaluated.  Bit-exact by construction, ~5
// instructions instead of ~12 on the common path.  `inv` must be finite (caller checks s).
constexpr bool kRecipFastPath = false;
__device__ __forceinline__ float quotient_for_rint(float x, float s, float inv) {
  const float r = __fmul_rn(x, inv);
  const float d = fabsf(__fsub_rn(r, rintf(r)));
  if (d > 0.49993896484375f) return __fdiv_rn(x, s);  // 0.5 - 2^-14: too close to a tie to trust r
  return r;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <bool BF>
__device__ __forceinline__ void round_pair(float& a, float& b) {
  if (BF) {
    const uint32_t w = pack_bf16x2(a, b);
    a = __uint_as_float(w << 16);
    b = __uint_as_float(w & 0xFFFF0000u);
  } else {
    uint32_t w = pack_f16x2(a, b);
    const float2 f = __half22float2(*reinterpret_cast<__half2*>(&w));
    a = f.x;
    b = f.y;
  }
}
template <typename T>
__device__ __forceinline__ void round_pair_t(float& a, float& b) {  // (a, b) -> (f32(T(a)), f32(T(b)))
  if (sizeof(T) == 2) round_pair<Elem<T>::IS_BF16>(a, b);
}

// Element arithmetic of the prologue, selected at compile time so the per-vector code is branch-free:
//   QM_ROUND        q = sat(rint(x))
//   QM_SCALE_RECIP  q = sat(rint(T(x * (1/qs))))   torch CUDA: tensor / python scalar = multiply by the reciprocal
//   QM_SCALE_DIV    q = sat(rint(T(x / qs)))       torch CPU: true division
//   QM_ROW_DIV      q = sat(rint(f32(x) / s))      per-token / caller-supplied row scale: fp32 tensor / fp32 tensor
//   QM_TENSOR_DYN   q = T(x / s)                   per-tensor dynamic (fp8): T tensor / 0-dim T tensor
enum QMode : int { QM_ROUND = 0, QM_SCALE_RECIP = 1, QM_SCALE_DIV = 2, QM_ROW_DIV = 3, QM_TENSOR_DYN = 4 };

template <typename T, bool FP8, int QM>
__device__ __forceinline__ void quantize_vec(const uint4& v, uint8_t* dst, float scale, const LinearParams& p) {
  constexpr int VEC = Elem<T>::VEC;
  float f[VEC];
  Elem<T>::unpack(v, f);
#pragma unroll
  for (int i = 0; i < VEC; i += 2) {
    float a = f[i], b = f[i + 1];
    if (QM == QM_ROW_DIV) { a = __fdiv_rn(a, scale); b = __fdiv_rn(b, scale); }
    else if (QM == QM_SCALE_RECIP) { a = __fmul_rn(a, scale); b = __fmul_rn(b, scale); round_pair_t<T>(a, b); }  // scale = 1/qs
    else if (QM == QM_SCALE_DIV) { a = __fdiv_rn(a, scale); b = __fdiv_rn(b, scale); round_pair_t<T>(a, b); }    // scale = qs
    else if (QM == QM_TENSOR_DYN) { a = __fdiv_rn(a, scale); b = __fdiv_rn(b, scale); round_pair_t<T>(a, b); }
    f[i] = a;
    f[i + 1] = b;
  }
  pack_store<FP8, VEC>(dst, f);
}

template <typename T>
__device__ __forceinline__ float vec_absmax(const uint4& v, float amax) {
  constexpr int VEC = Elem<T>::VEC;
  float f[VEC];
  Elem<T>::unpack(v, f);
#pragma unroll
  for (int i = 0; i < VEC; ++i) amax = fmaxf(amax, fabsf(f[i]));
  return amax;
}

// 1/s for the division-free path, or 0 when s is zero / subnormal-ish / non-finite (then every element of
// the row takes the IEEE division, which also reproduces the 0/0 = NaN -> 0 of an all-zero row).
__device__ __forceinline__ float safe_inverse(float s) {
  return (s > 1e-30f && s < 1e30f) ? __frcp_rn(s) : 0.f;
}

template <typename T>
__device__ __forceinline__ float token_scale(float amax, const LinearParams& p) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  // absmax is exactly representable in T; the division by qmax happens in T (torch: fp32 math, one
  // rounding to T).  CUDA torch multiplies by the fp32 reciprocal of a scalar divisor.
  return (p.div_mode == ASQ_DIV_RECIPROCAL) ? Elem<T>::round_to(__fmul_rn(amax, p.inv_qmax))
                                            : Elem<T>::round_to(__fdiv_rn(amax, p.qmax));
}

template <typename T, bool FP8, int QM>
__device__ __forceinline__ float quantize_row(const T* __restrict__ xrow, uint8_t* __restrict__ qrow,
                                              int K, int lane, const LinearParams& p, float given_scale, bool find_absmax) {
  constexpr int VEC = Elem<T>::VEC;
  constexpr int STEP = 32 * VEC;          // elements one warp-wide vector load covers
  constexpr int CHUNK = STEP * QBATCH;    // elements per batch
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  float scale = given_scale;
  uint4 buf[QBATCH];

  if (QM == QM_ROW_DIV && find_absmax) {  // per-token: the row's own absmax first
    float amax = 0.f;
    if (K <= CHUNK) {  // whole row lives in registers: one HBM read
#pragma unroll
      for (int j = 0; j < QBATCH; ++j) {
        const int c = lane * VEC + j * STEP;
        buf[j] = (c < K) ? __ldg(reinterpret_cast<const uint4*>(xrow + c)) : zero;
      }
#pragma unroll
      for (int j = 0; j < QBATCH; ++j) amax = vec_absmax<T>(buf[j], amax);
      scale = token_scale<T>(amax, p);
#pragma unroll
      for (int j = 0; j < QBATCH; ++j) {
        const int c = lane * VEC + j * STEP;
        if (c < K) quantize_vec<T, FP8, QM>(buf[j], qrow + c, scale, p);
      }
      return scale;
    }
    for (int c0 = 0; c0 < K; c0 += CHUNK) {
#pragma unroll
      for (int j = 0; j < QBATCH; ++j) {
        const int c = c0 + lane * VEC + j * STEP;
        buf[j] = (c < K) ? __ldg(reinterpret_cast<const uint4*>(xrow + c)) : zero;
      }
#pragma unroll
      for (int j = 0; j < QBATCH; ++j) amax = vec_absmax<T>(buf[j], amax);
    }
    scale = token_scale<T>(amax, p);
  }
  for (int c0 = 0; c0 < K; c0 += CHUNK) {  // second pass of a long per-token row re-reads it from L2
#pragma unroll
    for (int j = 0; j < QBATCH; ++j) {
      const int c = c0 + lane * VEC + j * STEP;
      buf[j] = (c < K) ? __ldg(reinterpret_cast<const uint4*>(xrow + c)) : zero;
    }
#pragma unroll
    for (int j = 0; j < QBATCH; ++j) {
      const int c = c0 + lane * VEC + j * STEP;
      if (c < K) quantize_vec<T, FP8, QM>(buf[j], qrow + c, scale, p);
    }
  }
  return (QM == QM_ROW_DIV || QM == QM_TENSOR_DYN) ? scale : 0.f;  // QM_SCALE_*: `scale` carried 1/qs or qs, not a row scale
}

// RMSNorm + round-only quantisation of one row by one warp (ASQ_ACT_RMSNORM).  Bit-identical to the stand-alone
// asq_add_rmsnorm_quant kernel (asq_glue.cu): that kernel gives vector v of a row to thread v % 128 of a 128-thread
// CTA, sums squares per thread with fmaf in ascending order, reduces every warp with an xor-shuffle tree and adds
// the four warp sums left to right.  Lane l of this warp holds vectors l + 32 m, i.e. those of the virtual
// threads l + 32 w with w = m % 4: four partial sums per lane, four shuffle trees, the same final order.
template <typename T>
__device__ __forceinline__ float quantize_row_rmsnorm(const T* __restrict__ xrow, uint8_t* __restrict__ qrow, int K, int lane,
                                                      const LinearParams& p) {
  constexpr int VEC = Elem<T>::VEC;
  static_assert(VEC == 8, "RMSNorm prologue: 16-bit activations only");
  constexpr int STEP = 32 * VEC;
  constexpr int CHUNK = STEP * QBATCH;  // 4096 elements = 16 vectors per lane: (chunk start / STEP) % 4 == 0
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  const T* wv = reinterpret_cast<const T*>(p.norm_weight);
  float ssw[4] = {0.f, 0.f, 0.f, 0.f};
  uint4 buf[QBATCH];
  for (int c0 = 0; c0 < K; c0 += CHUNK) {
#pragma unroll
    for (int j = 0; j < QBATCH; ++j) {
      const int c = c0 + lane * VEC + j * STEP;
      buf[j] = (c < K) ? __ldg(reinterpret_cast<const uint4*>(xrow + c)) : zero;
    }
#pragma unroll
    for (int j = 0; j < QBATCH; ++j) {
      if (c0 + lane * VEC + j * STEP < K) {
        float f[VEC];
        Elem<T>::unpack(buf[j], f);
#pragma unroll
        for (int i = 0; i < VEC; ++i) ssw[j & 3] = fmaf(f[i], f[i], ssw[j & 3]);
      }
    }
  }
#pragma unroll
  for (int w = 0; w < 4; ++w)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssw[w] += __shfl_xor_sync(0xffffffffu, ssw[w], o);
  const float ss = ssw[0] + ssw[1] + ssw[2] + ssw[3];
  const float rstd = rsqrtf(ss / static_cast<float>(K) + p.norm_eps);
  for (int c0 = 0; c0 < K; c0 += CHUNK) {
    if (K > CHUNK) {  // a long row does not fit the registers: second read (L2)
#pragma unroll
      for (int j = 0; j < QBATCH; ++j) {
        const int c = c0 + lane * VEC + j * STEP;
        buf[j] = (c < K) ? __ldg(reinterpret_cast<const uint4*>(xrow + c)) : zero;
      }
    }
#pragma unroll
    for (int j = 0; j < QBATCH; ++j) {
      const int c = c0 + lane * VEC + j * STEP;
      if (c < K) {
        float f[VEC], g[VEC];
        Elem<T>::unpack(buf[j], f);
        Elem<T>::unpack(__ldg(reinterpret_cast<const uint4*>(wv + c)), g);
#pragma unroll
        for (int i = 0; i < VEC; ++i) f[i] = Elem<T>::round_to(__fmul_rn(g[i], Elem<T>::round_to(__fmul_rn(f[i], rstd))));
        pack_store<false, VEC>(qrow + c, f);
      }
    }
  }
  return 0.f;
}
template <>
__device__ __forceinline__ float quantize_row_rmsnorm<float>(const float*, uint8_t*, int, int, const LinearParams&) {
  return 0.f;  // rejected on the host: the norm runs in a 16-bit activation dtype
}

template <typename T, bool FP8>
__device__ __forceinline__ float quantize_row_typed(const LinearParams& p, int row, int lane, float tensor_scale) {
  const size_t off = static_cast<size_t>(row) * p.K;
  const T* xrow = reinterpret_cast<const T*>(p.x) + off;
  uint8_t* qrow = p.a_q + off;
  switch (p.act_mode) {  // warp-uniform: one branch per row, straight-line code per vector
    case ASQ_ACT_RMSNORM:
      return FP8 ? 0.f : quantize_row_rmsnorm<T>(xrow, qrow, p.K, lane, p);
    case ASQ_ACT_PER_TOKEN:
      return quantize_row<T, FP8, QM_ROW_DIV>(xrow, qrow, p.K, lane, p, 0.f, true);
    case ASQ_ACT_ROW_SCALE_GIVEN:
      return quantize_row<T, FP8, QM_ROW_DIV>(xrow, qrow, p.K, lane, p, __ldg(p.row_scale_in + row), false);
    case ASQ_ACT_PER_TENSOR_DYNAMIC:
      return quantize_row<T, FP8, QM_TENSOR_DYN>(xrow, qrow, p.K, lane, p, tensor_scale, false);
    case ASQ_ACT_SCALE: {
      float qs = p.quant_scale, inv = p.inv_quant_scale;
      if (p.group_quant_scale != nullptr) {  // grouped launch: every expert has its own input scale
        const int grp = __ldg(p.group_of_blk + row / BLOCK_M);
        qs = grp >= 0 ? __ldg(p.group_quant_scale + grp) : 1.f;
        inv = __frcp_rn(qs);  // == the host's 1.0f / qs
      }
      return (p.div_mode == ASQ_DIV_RECIPROCAL) ? quantize_row<T, FP8, QM_SCALE_RECIP>(xrow, qrow, p.K, lane, p, inv, false)
                                                : quantize_row<T, FP8, QM_SCALE_DIV>(xrow, qrow, p.K, lane, p, qs, false);
    }
    default:
      return quantize_row<T, FP8, QM_ROUND>(xrow, qrow, p.K, lane, p, 0.f, false);
  }
}

template <bool FP8>
__device__ __forceinline__ float quantize_row_any(const LinearParams& p, int row, int lane, float tensor_scale = 0.f) {
  if (p.x_dtype == ASQ_BF16) return quantize_row_typed<__nv_bfloat16, FP8>(p, row, lane, tensor_scale);
  if (p.x_dtype == ASQ_F16) return quantize_row_typed<__half, FP8>(p, row, lane, tensor_scale);
  return quantize_row_typed<float, FP8>(p, row, lane, tensor_scale);
}

// Per-tensor DYNAMIC scale (per_tensor_quantize_fp8, quantization.py:144-170): s = T(max|x|) / T(448) over the
// whole tensor.  Every participating warp folds the absmax of its rows into one global word (atomicMax on
// the bit pattern), then all warps wait for each other — a grid-wide dependency, safe because the launch
// never has more CTAs than can be co-resident.
template <typename T>
__device__ __forceinline__ float rows_absmax(const LinearParams& p, int first_row, int row_step, int lane) {
  constexpr int VEC = Elem<T>::VEC;
  float amax = 0.f;
  for (int row = first_row; row < p.M; row += row_step) {
    const T* xrow = reinterpret_cast<const T*>(p.x) + static_cast<size_t>(row) * p.K;
#pragma unroll 8
    for (int c = lane * VEC; c < p.K; c += 32 * VEC)
      amax = vec_absmax<T>(__ldg(reinterpret_cast<const uint4*>(xrow + c)), amax);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  return amax;
}

__device__ __forceinline__ float tensor_scale_phase(const LinearParams& p, int first_row, int row_step,
                                                    int participants, int lane) {
  float amax = (p.x_dtype == ASQ_BF16) ? rows_absmax<__nv_bfloat16>(p, first_row, row_step, lane)
               : (p.x_dtype == ASQ_F16) ? rows_absmax<__half>(p, first_row, row_step, lane)
                                        : rows_absmax<float>(p, first_row, row_step, lane);
  uint32_t bits = 0;
  if (lane == 0) {
    atomicMax(p.sync + SYNC_AMAX, __float_as_uint(amax));
    red_release_gpu_add(p.sync + SYNC_AMAX_CNT, 1u);
    while (ld_acquire_gpu(p.sync + SYNC_AMAX_CNT) < static_cast<uint32_t>(participants)) __nanosleep(128);
    bits = ld_acquire_gpu(p.sync + SYNC_AMAX);
  }
  bits = __shfl_sync(0xffffffffu, bits, 0);
  amax = __uint_as_float(bits);
  const bool recip = (p.div_mode == ASQ_DIV_RECIPROCAL);
  const float s = recip ? __fmul_rn(amax, p.inv_qmax) : __fdiv_rn(amax, p.qmax);  // tensor / python float, in T
  if (p.x_dtype == ASQ_BF16) return Elem<__nv_bfloat16>::round_to(s);
  if (p.x_dtype == ASQ_F16) return Elem<__half>::round_to(s);
  return s;
}

// ------------------------------------------------------------------ epilogue
__device__ __forceinline__ float load_bias_any(const void* b, int dtype, int col) {
  if (dtype == ASQ_I8) return static_cast<float>(reinterpret_cast<const int8_t*>(b)[col]);
  if (dtype == ASQ_I32) return static_cast<float>(reinterpret_cast<const int32_t*>(b)[col]);
  return reinterpret_cast<const float*>(b)[col];
}

// 32 consecutive fp32 values of a per-column vector starting at col0 (multiple of 32): vector loads when
// the whole chunk is in range and the pointer is 16-byte aligned, guarded scalar loads otherwise.
// Every index is a compile-time constant so the array stays in registers.
__device__ __forceinline__ void load_cols32(const float* __restrict__ src, int col0, int N, float (&out)[32]) {
  if (col0 + 32 <= N && (reinterpret_cast<uintptr_t>(src + col0) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = __ldg(s4 + j);
      out[4 * j] = t.x; out[4 * j + 1] = t.y; out[4 * j + 2] = t.z; out[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) out[j] = (col0 + j < N) ? __ldg(src + col0 + j) : 0.f;
  }
}

__device__ __forceinline__ float round_out(float v, int y_dtype) {
  if (y_dtype == ASQ_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  if (y_dtype == ASQ_F16) return __half2float(__float2half_rn(v));
  return v;
}

// Raw accumulators r[32] (row `row`, columns col0..col0+31) -> fp32 results v[32] following the reference's
// order of operations (linear.py:93,104 / :197-207): factor first, then * acc, then + bias, each a separate
// fp32 rounding (no FMA contraction), so the result is bit-identical to eager torch.
template <bool FP8>
__device__ __forceinline__ void epilogue_values(const uint32_t (&r)[32], float (&v)[32], int col0, float rs,
                                                const LinearParams& p, float dequant_scale) {
  const int N = p.N;
  if (p.epi_kind != EPI_ALPHA_BETA) {
    const bool per_token = (p.act_mode == ASQ_ACT_PER_TOKEN || p.act_mode == ASQ_ACT_ROW_SCALE_GIVEN ||
                            p.act_mode == ASQ_ACT_PER_TENSOR_DYNAMIC);
    const float f_scalar = per_token ? __fmul_rn(dequant_scale, rs) : dequant_scale;
    if (p.col_scale != nullptr) {
      float cs[32];
      load_cols32(p.col_scale, col0, N, cs);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float a = FP8 ? __uint_as_float(r[j]) : __int2float_rn(static_cast<int32_t>(r[j]));
        const float f = per_token ? __fmul_rn(cs[j], rs) : cs[j];
        v[j] = __fmul_rn(f, a);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float a = FP8 ? __uint_as_float(r[j]) : __int2float_rn(static_cast<int32_t>(r[j]));
        v[j] = __fmul_rn(f_scalar, a);
      }
    }
    if (p.bias != nullptr) {
      float b[32];
      load_cols32(p.bias, col0, N, b);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __fadd_rn(v[j], b[j]);
    }
    if (p.out_fq_scale != 0.f) {
      // FP8LinearStatic output fake-quantisation (linear.py:562-564): y = T(e4m3(clamp(T(T(v) / os))) * os)
      const bool recip = (p.div_mode == ASQ_DIV_RECIPROCAL);
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float t0 = round_out(v[j], p.y_dtype), t1 = round_out(v[j + 1], p.y_dtype);
        t0 = round_out(recip ? __fmul_rn(t0, p.inv_out_fq_scale) : __fdiv_rn(t0, p.out_fq_scale), p.y_dtype);
        t1 = round_out(recip ? __fmul_rn(t1, p.inv_out_fq_scale) : __fdiv_rn(t1, p.out_fq_scale), p.y_dtype);
        const uint32_t q2 = cvt_e4m3x2(t0, t1);
        uint32_t h2;
        asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"(static_cast<uint16_t>(q2)));
        const float2 d = __half22float2(*reinterpret_cast<__half2*>(&h2));
        v[j] = __fmul_rn(d.x, p.out_fq_scale);  // the final rounding to T happens in the store
        v[j + 1] = __fmul_rn(d.y, p.out_fq_scale);
      }
    }
  } else {  // EPI_ALPHA_BETA: v = alpha*acc + beta*bias  (cublasLt o8 / csrc/kernels/linear.cu epilogues)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float a = __int2float_rn(static_cast<int32_t>(r[j]));
      float t = __fmul_rn(p.alpha, a);
      if (p.bias_any != nullptr && col0 + j < N)
        t = __fadd_rn(t, __fmul_rn(p.beta, load_bias_any(p.bias_any, p.bias_dtype, col0 + j)));
      if (p.flags & ASQ_EPI_RELU) t = fmaxf(t, 0.f);
      v[j] = t;
    }
  }
}

template <bool FP8>
__device__ __forceinline__ void epilogue_values(const uint32_t (&r)[32], float (&v)[32], int col0, float rs,
                                                const LinearParams& p) {
  epilogue_values<FP8>(r, v, col0, rs, p, p.dequant_scale);
}

// 32 fp32 results -> packed output words of the requested type (w[] holds 32 * elem_size / 4 words).
__device__ __forceinline__ void pack_out16(const float (&v)[32], uint32_t (&w)[16], bool bf) {
  if (bf) {  // one warp-uniform branch, not a select per element
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = pack_f16x2(v[2 * j], v[2 * j + 1]);
  }
}

// EPI_SWIGLU: vg / vu = dequantised (fp32) gate and up values of the same 32 logical columns.  Follows the eager
// chain T(gate), T(up) -> T(silu) -> T(* up) [-> T(/ quant_scale) -> rint -> saturate] with the same
// arithmetic as asq_glue.cu's silu_mul_vec (SFU exp / reciprocal), so the fused epilogue and the stand-alone
// producer kernel emit identical bytes.  BF / OUT are compile-time so the 32 elements are straight-line code:
// every rounding to T is a packed F2FP (two values per instruction) followed by a shift / mask unpack.
//   OUT 0: w[16] = the product as T;  OUT 1: w[8] = int8, division as reciprocal multiply;  OUT 2: int8, IEEE division
template <bool BF, int OUT>
__device__ __forceinline__ void swiglu_chunk(const float (&vg)[32], const float (&vu)[32], uint32_t* w, const LinearParams& p) {
  const float qs = p.out_quant_scale, inv = p.inv_out_quant_scale;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float t[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float g0 = vg[4 * j + 2 * h], g1 = vg[4 * j + 2 * h + 1], u0 = vu[4 * j + 2 * h], u1 = vu[4 * j + 2 * h + 1];
      round_pair<BF>(g0, g1);
      round_pair<BF>(u0, u1);
      float s0 = __fdividef(g0, 1.0f + __expf(-g0)), s1 = __fdividef(g1, 1.0f + __expf(-g1));
      round_pair<BF>(s0, s1);
      float a0 = __fmul_rn(s0, u0), a1 = __fmul_rn(s1, u1);
      if (OUT == 0) {
        w[2 * j + h] = BF ? pack_bf16x2(a0, a1) : pack_f16x2(a0, a1);
      } else {
        round_pair<BF>(a0, a1);
        float t0 = OUT == 1 ? __fmul_rn(a0, inv) : __fdiv_rn(a0, qs), t1 = OUT == 1 ? __fmul_rn(a1, inv) : __fdiv_rn(a1, qs);
        round_pair<BF>(t0, t1);
        t[2 * h] = t0;
        t[2 * h + 1] = t1;
      }
    }
    if (OUT != 0) w[j] = cvt_s8x4(t[0], t[1], t[2], t[3]);
  }
}
// RoPE on one 32-column chunk pair of a 128-wide head: va = columns d .. d+31 of the first half, vb = the same
// columns of the second half (d + 64).  out[d] = T(T(x[d]*cos[d]) + T(rot[d]*sin[d])), rot = (-x2, x1), every
// product rounded to T first — the arithmetic of asq_glue.cu's rope_kernel (HF apply_rotary_pos_emb in T).
// Tables are BLOCKED [16][S][8]: entry (pos, col) lives at ((col / 8) * S + pos) * 8 + col % 8, so the 32 lanes of
// a warp (32 consecutive positions) read 512 contiguous bytes per 16-byte load instead of 32 separate lines.
// cosp / sinp point at block (d / 8) of this row's position; `blk` = S * 8 elements between blocks.
template <bool BF>
__device__ __forceinline__ void unpack_pair(uint32_t w, float& a, float& b) {
  if (BF) {
    a = __uint_as_float(w << 16);
    b = __uint_as_float(w & 0xFFFF0000u);
  } else {
    const float2 f = __half22float2(*reinterpret_cast<__half2*>(&w));
    a = f.x;
    b = f.y;
  }
}
template <bool BF>
__device__ __forceinline__ void rope_chunk(const float (&va)[32], const float (&vb)[32], uint32_t (&wa)[16], uint32_t (&wb)[16],
                                           bool rot, const uint16_t* __restrict__ cosp, const uint16_t* __restrict__ sinp,
                                           size_t blk, bool halves_equal) {
  if (!rot) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      wa[j] = BF ? pack_bf16x2(va[2 * j], va[2 * j + 1]) : pack_f16x2(va[2 * j], va[2 * j + 1]);
      wb[j] = BF ? pack_bf16x2(vb[2 * j], vb[2 * j + 1]) : pack_f16x2(vb[2 * j], vb[2 * j + 1]);
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 ca = __ldg(reinterpret_cast<const uint4*>(cosp + q * blk)), sa = __ldg(reinterpret_cast<const uint4*>(sinp + q * blk));
    // HF tables repeat their first half (emb = cat(freqs, freqs)); the caller vouches for it to halve the table reads
    const uint4 cb = halves_equal ? ca : __ldg(reinterpret_cast<const uint4*>(cosp + (8 + q) * blk));
    const uint4 sb = halves_equal ? sa : __ldg(reinterpret_cast<const uint4*>(sinp + (8 + q) * blk));
    const uint32_t caw[4] = {ca.x, ca.y, ca.z, ca.w}, saw[4] = {sa.x, sa.y, sa.z, sa.w};
    const uint32_t cbw[4] = {cb.x, cb.y, cb.z, cb.w}, sbw[4] = {sb.x, sb.y, sb.z, sb.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float x0 = va[8 * q + 2 * i], x1 = va[8 * q + 2 * i + 1], y0 = vb[8 * q + 2 * i], y1 = vb[8 * q + 2 * i + 1];
      round_pair<BF>(x0, x1);  // the projection's output in T
      round_pair<BF>(y0, y1);
      float c0, c1, s0, s1;
      unpack_pair<BF>(caw[i], c0, c1);
      unpack_pair<BF>(saw[i], s0, s1);
      float p0 = __fmul_rn(x0, c0), p1 = __fmul_rn(x1, c1), q0 = __fmul_rn(-y0, s0), q1 = __fmul_rn(-y1, s1);
      round_pair<BF>(p0, p1);
      round_pair<BF>(q0, q1);
      const float oa0 = __fadd_rn(p0, q0), oa1 = __fadd_rn(p1, q1);
      wa[4 * q + i] = BF ? pack_bf16x2(oa0, oa1) : pack_f16x2(oa0, oa1);
      unpack_pair<BF>(cbw[i], c0, c1);
      unpack_pair<BF>(sbw[i], s0, s1);
      p0 = __fmul_rn(y0, c0); p1 = __fmul_rn(y1, c1); q0 = __fmul_rn(x0, s0); q1 = __fmul_rn(x1, s1);
      round_pair<BF>(p0, p1);
      round_pair<BF>(q0, q1);
      const float ob0 = __fadd_rn(p0, q0), ob1 = __fadd_rn(p1, q1);
      wb[4 * q + i] = BF ? pack_bf16x2(ob0, ob1) : pack_f16x2(ob0, ob1);
    }
  }
}

__device__ __forceinline__ void swiglu_dispatch(const float (&vg)[32], const float (&vu)[32], uint32_t* w, const LinearParams& p) {
  const bool bf = (p.mid_dtype == ASQ_BF16);
  if (p.y_dtype != ASQ_I8) {
    if (bf) swiglu_chunk<true, 0>(vg, vu, w, p); else swiglu_chunk<false, 0>(vg, vu, w, p);
  } else if (p.div_mode == ASQ_DIV_RECIPROCAL) {
    if (bf) swiglu_chunk<true, 1>(vg, vu, w, p); else swiglu_chunk<false, 1>(vg, vu, w, p);
  } else {
    if (bf) swiglu_chunk<true, 2>(vg, vu, w, p); else swiglu_chunk<false, 2>(vg, vu, w, p);
  }
}

// Fallback store (odd N, int8 output, ...): direct global writes, predicated per element.
template <bool FP8>
__device__ __forceinline__ void store_chunk_direct(const uint32_t (&r)[32], int row, int col0, float rs,
                                                   const LinearParams& p) {
  if (row >= p.M || col0 >= p.N) return;
  const int N = p.N;
  const size_t off = static_cast<size_t>(row) * N + col0;
  if (p.epi_kind == EPI_RAW_I32) {
    int32_t* dst = reinterpret_cast<int32_t*>(p.y) + off;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < N) dst[j] = static_cast<int32_t>(r[j]);
    return;
  }
  float v[32];
  epilogue_values<FP8>(r, v, col0, rs, p);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (col0 + j < N) {
      switch (p.y_dtype) {
        case ASQ_BF16: reinterpret_cast<__nv_bfloat16*>(p.y)[off + j] = __float2bfloat16_rn(v[j]); break;
        case ASQ_F16: reinterpret_cast<__half*>(p.y)[off + j] = __float2half_rn(v[j]); break;
        case ASQ_F32: reinterpret_cast<float*>(p.y)[off + j] = v[j]; break;
        case ASQ_I32: reinterpret_cast<int32_t*>(p.y)[off + j] = __float2int_rn(v[j]); break;
        case ASQ_I8: reinterpret_cast<int8_t*>(p.y)[off + j] = static_cast<int8_t>(cvt_s8(v[j])); break;
        default: break;
      }
    }
  }
}

// Staged store: this lane's 32 results go to row `lane` of the warp's 32-row x 128-byte staging tile in
// shared memory, 16-byte chunks XOR-swizzled with (row & 7) exactly as CU_TENSOR_MAP_SWIZZLE_128B expects
// (also makes the quarter-warp stores bank-conflict free); one elected lane then issues a TMA store, which
// writes full 128-byte lines and clips rows >= M / columns >= N by itself.
// `chunk16` = index of the first 16-byte chunk inside the 128-byte row (0 or 4 for 2-byte outputs).
template <int NW>
__device__ __forceinline__ void stage_words(uint32_t stage_base, int lane, int chunk16, const uint32_t (&w)[NW]) {
#pragma unroll
  for (int c = 0; c < NW / 4; ++c) {
    const uint32_t addr = stage_base + lane * 128 + (((chunk16 + c) ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * c]), "r"(w[4 * c + 1]),
                 "r"(w[4 * c + 2]), "r"(w[4 * c + 3])
                 : "memory");
  }
}

// ------------------------------------------------------------------ tile schedule
// Tile walk.  raster_m == 0: M-tile by M-tile with N fastest, inside groups of `group` N-blocks (one
// group's W tiles stay L2-resident; the first wave only needs the first activation panels, so it can
// start while phase 1 is still quantising later rows).  raster_m == 1: groups of `group` M-tiles with M
// fastest inside a group (concurrent CTAs share W tiles).
__device__ __forceinline__ void tile_coords(int t, const LinearParams& p, int& m_blk, int& n_blk) {
  if (p.raster_m) {
    const int group_size = p.group * p.num_n_blocks;
    const int g = t / group_size;
    const int first_m = g * p.group;
    const int gm = min(p.num_m_blocks - first_m, p.group);
    const int local = t - g * group_size;
    m_blk = first_m + local % gm;
    n_blk = local / gm;
  } else {
    const int group_size = p.group * p.num_m_blocks;
    const int g = t / group_size;
    const int first_n = g * p.group;
    const int gn = min(p.num_n_blocks - first_n, p.group);
    const int local = t - g * group_size;
    m_blk = local / gn;
    n_blk = first_n + local % gn;
  }
}

// Tile geometry.  CG = CTAs per tile: 1 -> a 128-row tile per CTA; 2 -> a CTA pair (cta_group::2) owns a
// 256-row tile: each CTA stages its own 128 A rows and HALF of the W tile, the leader issues 256 x N MMAs
// that read both CTAs' shared memory, each CTA drains its own 128 TMEM lanes.  Pairing cuts the bytes
// every SM pulls from L2 per MMA by a third, which is what bounds 8-bit GEMMs at full clock.
// Tiles are up to TILE_N = 256 columns wide; the scheduler may hand out narrower tiles in units of
// UNIT_N = 64 columns (UMMA N = 64/128/192/256) to balance the last, partial round of tiles.
constexpr int TILE_N = 256;
constexpr int UNIT_N = 64;
constexpr uint32_t EPI_BUF_BYTES = 32 * 128;  // one warp's staging tile: 32 rows x 128 bytes
#ifndef ASQ_EPI_NBUF
#define ASQ_EPI_NBUF 2
#endif
constexpr uint32_t EPI_NBUF = ASQ_EPI_NBUF;   // staging buffers per warp
constexpr uint32_t EPI_BYTES = NUM_EPI_WARPS * EPI_BUF_BYTES * EPI_NBUF;
constexpr uint32_t SMEM_LIMIT = 232448;       // 227 KB per CTA on sm_100

template <int CG>
struct TileCfg {
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K;
  static constexpr uint32_t B_ROWS = TILE_N / CG;  // W rows staged per CTA for a full-width tile
  static constexpr uint32_t UNIT_ROWS = UNIT_N / CG;
  static constexpr uint32_t B_BYTES = B_ROWS * BLOCK_K;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t BUDGET = SMEM_LIMIT - EPI_BYTES - 256 - 1024;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * TILE_N;  // double-buffered accumulator
  static constexpr uint32_t EPI_OFFSET = STAGES * STAGE_BYTES;  // multiple of 1024
  static constexpr uint32_t BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
  static constexpr uint32_t SMEM_BYTES = BAR_OFFSET + 256 + 1024;  // + barriers + alignment slack
  static constexpr int TILE_M = BLOCK_M * CG;
  static_assert(STAGES >= 3, "pipeline too shallow");
};

// The sequence of tiles one worker (CTA or CTA pair) processes; every warp role walks it identically.
//   rounds x  full tiles  t = round * W + worker   (round-robin)
//   then      the R = T - rounds * W left-over tiles: one per worker, or — stream-K — their k-iterations dealt
//             evenly to the workers, every tile finished by the worker that holds its first k-block (the
//             "owner", which adds the other workers' partial accumulators before its epilogue).
struct Seg {
  int m_blk, col0, width;  // output tile
  int kb0, kb1;            // k-block range [kb0, kb1) this worker accumulates
  int role;                // SEG_COMPLETE: whole K, normal epilogue; SEG_CONTRIB: writes a partial; SEG_OWNER: adds partials
  int tile_g0;             // owner: first stream-K iteration index of its tile
  int ar_tile;             // all-reduce mode: linear tile id (owner = ar_tile % world, slot = ar_tile / world)
  int group;               // grouped GEMM: the group (expert) of this tile's rows, 0 otherwise
};
enum : int { SEG_COMPLETE = 0, SEG_CONTRIB = 1, SEG_OWNER = 2, SEG_AR_CONTRIB = 3, SEG_AR_OWNER = 4 };

// Feature bits of a kernel instantiation.  The full kernel (every epilogue, stream-K, grouping, the fused
// prologue, ...) is > 1 MB of SASS and every feature added a fixed ~0.3-0.5 us per launch; the launches of the
// Llama layer therefore use lean instantiations that compile only what they execute, everything else runs the
// full one.  A launch may use an instantiation iff its needs are a subset of the instantiation's bits.
enum Feat : int {
  F_AR = 1,        // fused all-reduce roles (+ the peers' output maps as launch parameters)
  F_SK = 2,        // stream-K tail
  F_GROUP = 4,     // grouped / batched weight selection, skipped padding tiles
  F_SWIGLU = 8,    // SwiGLU epilogue
  F_ROPE = 16,     // RoPE epilogue
  F_DEQ16 = 32,    // the standard staged 16-bit dequant epilogue
  F_OUT_ANY = 64,  // 4-byte staged outputs, raw int32, alpha/beta, direct (unstaged) stores
  F_PHASE1 = 128,  // fused activation-quantisation prologue
  F_RESID = 256,   // residual add in the 16-bit epilogue: y = T(residual + T(dequantised result))
  F_NVLS = 512,    // in-switch all-reduce of the 16-bit partial tiles (multimem.ld_reduce / multimem.st)
  F_FULL = F_SK | F_GROUP | F_SWIGLU | F_ROPE | F_DEQ16 | F_OUT_ANY | F_PHASE1 | F_RESID,
};

constexpr int extra_kind(int feat) { return (feat & F_AR) ? 1 : ((feat & F_RESID) ? 2 : 0); }

template <int FEAT>
struct TileWalk {
  const LinearParams& p;
  int worker, W, round, g, g_end;
  __device__ static long long sk_begin(const LinearParams& p, int w) {
    return static_cast<long long>(w) * p.sk_total / p.sk_workers;
  }
  __device__ bool sk_on() const { return (FEAT & F_SK) != 0 && p.sk_enabled != 0; }
  __device__ TileWalk(const LinearParams& p_, int worker_, int W_) : p(p_), worker(worker_), W(W_), round(0), g(0), g_end(0) {
    if (sk_on() && worker_ < p_.sk_workers) {
      g = static_cast<int>(sk_begin(p_, worker_));
      g_end = static_cast<int>(sk_begin(p_, worker_ + 1));
    } else if (!sk_on() && worker_ < p_.tail_tiles) {
      g = worker_;  // plain tail: one left-over tile per worker
      g_end = worker_ + 1;
    }
  }
  __device__ void set_tile(Seg& sg, int t) const {
    int n_blk;
    if ((FEAT & F_AR) != 0 && p.ar_world > 1) {
      // walk order of this rank: the tiles owned by rank+1, rank+2, ... first, its own tiles last
      int i = t, owner = p.ar_rank;
      for (int s = 1; s <= p.ar_world; ++s) {
        owner = (p.ar_rank + s) % p.ar_world;
        const int cnt = (p.ar_tiles - owner + p.ar_world - 1) / p.ar_world;
        if (i < cnt) break;
        i -= cnt;
      }
      t = owner + p.ar_world * i;
      sg.ar_tile = t;
      sg.role = (owner == p.ar_rank) ? SEG_AR_OWNER : SEG_AR_CONTRIB;
    }
    if ((FEAT & F_NVLS) != 0) sg.ar_tile = t;  // same walk order on every rank; owner = t % world
    tile_coords(t, p, sg.m_blk, n_blk);
    sg.group = 0;
    if ((FEAT & F_GROUP) != 0)
      sg.group = (p.group_of_blk != nullptr) ? __ldg(p.group_of_blk + sg.m_blk * p.tile_m_blocks)
                 : (p.batch_rows > 0 ? (sg.m_blk * p.tile_m_blocks * BLOCK_M) / p.batch_rows : 0);
    const int U = p.tile_units;  // tile width in 64-column units: 4, or fewer for decode-sized problems
    sg.col0 = n_blk * U * UNIT_N;
    sg.width = min(U, p.n_units - n_blk * U) * UNIT_N;
  }
  __device__ bool next(Seg& sg) {
    sg.kb0 = 0;
    sg.kb1 = p.num_k_blocks;
    sg.role = SEG_COMPLETE;
    sg.tile_g0 = 0;
    while (round < p.rounds) {
      set_tile(sg, round * W + worker);
      ++round;
      if (sg.group >= 0) return true;  // group -1: padding rows of a grouped launch, nothing to compute
    }
    if (g >= g_end) return false;
    if (!sk_on()) {
      while (g < g_end) {
        set_tile(sg, p.rounds * W + g);
        ++g;
        if (sg.group >= 0) return true;
      }
      return false;
    }
    // stream-K: this worker's share [g, g_end) of the tail's k-iterations, cut at tile boundaries
    const int nkb = p.num_k_blocks;
    const int j = g / nkb;
    sg.kb0 = g - j * nkb;
    sg.kb1 = min(nkb, sg.kb0 + (g_end - g));
    sg.tile_g0 = j * nkb;
    g += sg.kb1 - sg.kb0;
    set_tile(sg, p.rounds * W + j);
    if (sg.kb0 == 0) sg.role = (sg.kb1 == nkb) ? SEG_COMPLETE : SEG_OWNER;
    else sg.role = SEG_CONTRIB;
    return true;
  }
};

// Stream-K scratch addressing.  Slot = one CTA's 128 x 256 accumulator of a contributor; inside a slot every
// epilogue warp owns 16 KB laid out [chunk(4)][16-byte quad(8)][lane(32)] so that both the contributor's
// stores and the owner's loads are fully coalesced 512-byte warp accesses.
constexpr uint32_t SK_WARP_WORDS = 4 * 8 * 32 * 4;                  // 4096 words = 16 KB per warp
constexpr uint32_t SK_SLOT_WORDS = NUM_EPI_WARPS * SK_WARP_WORDS;   // 128 KB per CTA
__device__ __forceinline__ uint32_t* sk_warp_base(const LinearParams& p, int contributor, int cg, uint32_t cta_rank, int ew) {
  return p.sk_partial + (static_cast<size_t>(contributor) * cg + cta_rank) * SK_SLOT_WORDS + ew * SK_WARP_WORDS;
}
__device__ __forceinline__ uint32_t* sk_flag(const LinearParams& p, int contributor, int cg, uint32_t cta_rank, int ew) {
  return p.sk_flags + (static_cast<size_t>(contributor) * cg + cta_rank) * NUM_EPI_WARPS + ew;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// All-reduce scratch addressing on rank `dst` (its receive buffer / flag words as mapped into this process).
// `slot` = (source - dst - 1 + world) % world numbers the world-1 possible sources of a destination.
constexpr int AR_FLAG_BASE = 64;
__device__ __forceinline__ size_t ar_index(const LinearParams& p, int slot, int owned_idx, int cg, uint32_t cta_rank, int ew) {
  return ((static_cast<size_t>(slot) * p.ar_cnt_max + owned_idx) * cg + cta_rank) * NUM_EPI_WARPS + ew;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void red_release_sys_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- in-switch reduction (NVLS).  One 16-byte multimem.ld_reduce returns the SUM over all ranks' copies of eight
// 16-bit values (the switch adds them with fp32 accumulation); multimem.st writes 16 bytes into every rank's copy.
template <bool BF>
__device__ __forceinline__ uint4 multimem_ld_reduce_16(const void* mc_addr) {
  uint4 v;
  if (BF) asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc_addr) : "memory");
  else    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc_addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t multimem_ld_reduce_u32(const uint32_t* mc_addr) {
  uint32_t v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u32 %0, [%1];" : "=r"(v) : "l"(mc_addr) : "memory");
  return v;
}
template <bool BF>
__device__ __forceinline__ void multimem_st_16(void* mc_addr, const uint4& v) {
  if (BF) asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else    asm volatile("multimem.st.relaxed.sys.global.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Reducer CTAs of an NVLS launch (the last p.nvls_reducers CTAs of the grid; they compute no tiles).  The unit of work
// is a 64-row slab of a tile this rank owns (t = rank + world * i): it has its own counter, bumped by the four
// epilogue warps per rank that write it, and is reduced by ONE reducer warp — no CTA-wide synchronisation, so while
// one warp waits for a slab to land the others keep streaming.  Slabs are dealt round-robin to all reducer warps in
// walk order, i.e. in the order in which the GEMM CTAs of all ranks finish them.  A warp sums the world copies of
// its slab in the switch (multimem.ld_reduce, NVLS_INFLIGHT 16-byte requests in flight per lane, a warp-wide
// request covers 512 contiguous bytes of a row) and broadcasts the sums (multimem.st).  The GEMM CTAs never wait for
// a reducer: the in-switch reduction of one round of tiles overlaps the MMAs of the next.
constexpr int NVLS_INFLIGHT = 16;
constexpr int NVLS_SLAB_ROWS = 64;
template <bool BF, int CG>
__device__ __forceinline__ void nvls_reducer_warp(const LinearParams& p, int rwarp, int num_rwarps, int lane) {
  constexpr int SLABS = CG * BLOCK_M / NVLS_SLAB_ROWS;  // per tile
  // timeline slots of a reducer CTA (its warp 0): 2 first slab landed, 3 first slab reduced, 4 ns spent waiting for
  // counters, 5 ns spent reducing, 6 last slab reduced
  const bool stamp = p.dbg != nullptr && (threadIdx.x == 0);
  unsigned long long t_wait = 0, t_red = 0, t_a = 0, t_b = 0;
  bool first = true;
  const int owned = (p.ar_tiles - p.ar_rank + p.ar_world - 1) / p.ar_world;
  const uint32_t target = static_cast<uint32_t>(p.ar_world) * 4u;  // 2 quadrants x 2 column halves per rank
  const int N = p.N;
  for (int u = rwarp; u < owned * SLABS; u += num_rwarps) {
    const int i = u / SLABS, slab = u - i * SLABS;
    // the slab has landed everywhere when the SUM of all ranks' local counters reaches world x 4: one in-switch
    // reduction per poll instead of one remote load per rank
    const size_t ci = static_cast<size_t>(p.ar_rank + p.ar_world * i) * SLABS + slab;
    if (stamp) t_a = globaltimer_ns();
    if (lane == 0) {
      // cheap local poll first (every rank runs the same schedule: when this rank's part has landed the others' are
      // about to), then one switch round trip per poll
      while (ld_acquire_gpu(p.nvls_ctr + ci) != 4u) __nanosleep(128);
      while (multimem_ld_reduce_u32(p.nvls_ctr_mc + ci) != target) __nanosleep(256);
      // The data requests below are issued only after this poll has returned (they are control-dependent on it) and
      // are served by the switch from the ranks' L2s, never from a local cache, so a GPU-scope fence orders them; a
      // system-scope fence here waits behind every multimem request this SM has in flight (measured ~20 us per slab).
      __threadfence();
    }
    __syncwarp();
    if (stamp) { t_b = globaltimer_ns(); t_wait += t_b - t_a; if (first) ASQ_STAMP(2); }
    int m_blk, n_blk;
    tile_coords(p.ar_rank + p.ar_world * i, p, m_blk, n_blk);
    const int row0 = m_blk * BLOCK_M * CG + slab * NVLS_SLAB_ROWS;
    const int col0 = n_blk * p.tile_units * UNIT_N;
    const int rows = min(NVLS_SLAB_ROWS, p.M - row0);
    const int cpr = min(p.tile_units * UNIT_N, N - col0) / 8;  // 16-byte pieces per slab row (N % 8 == 0)
    const int pieces = rows * cpr;                            // <= 0: the slab lies beyond M
    for (int base = lane; base < pieces; base += 32 * NVLS_INFLIGHT) {
      uint4 v[NVLS_INFLIGHT];
      size_t off[NVLS_INFLIGHT];
#pragma unroll
      for (int k = 0; k < NVLS_INFLIGHT; ++k) {
        const int c = base + k * 32;
        const int r = c / cpr;
        off[k] = (static_cast<size_t>(row0 + r) * N + col0 + (c - r * cpr) * 8) * 2;
        if (c < pieces) v[k] = multimem_ld_reduce_16<BF>(p.nvls_p_mc + off[k]);
      }
#pragma unroll
      for (int k = 0; k < NVLS_INFLIGHT; ++k)
        if (base + k * 32 < pieces) multimem_st_16<BF>(p.nvls_y_mc + off[k], v[k]);
    }
    if (stamp) { t_red += globaltimer_ns() - t_b; if (first) ASQ_STAMP(3); first = false; }
  }
  if (stamp) { p.dbg[blockIdx.x * 8 + 4] = t_wait; p.dbg[blockIdx.x * 8 + 5] = t_red; ASQ_STAMP(6); }
}

// Output maps of the other ranks' y buffers (all-reduce mode), in rank order skipping self.  Only the AR
// instantiations of the kernel carry them: 896 bytes of launch parameters cost ~2 us per launch (measured).
// KIND 0: nothing; 1: the peers' y maps (all-reduce); 2: one map over the residual tensor (F_RESID without F_AR).
template <int KIND>
struct ExtraMapsT {
  char unused;
};
template <>
struct ExtraMapsT<1> {
  CUtensorMap m[7];
};
template <>
struct ExtraMapsT<2> {
  CUtensorMap m[1];
};

// ------------------------------------------------------------------ the kernel
// MC = CTA pairs per cluster (1 or 2).  With MC == 2 a cluster of four CTAs owns a 256-row x 512-column
// super tile: pair p computes columns [p*256, p*256+256), both pairs need the same activation rows, so every
// CTA fetches only HALF of its 128 A rows and multicasts them to its twin in the other pair.  Per k-block a
// CTA then pulls 8 KB (A) + 16 KB (W) from L2 instead of 32 KB; the main loop is L2-feed bound.
template <bool FP8, int CG, int MC, int FEAT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
asq_linear_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmBu, const __grid_constant__ CUtensorMap tmY,
                  const __grid_constant__ ExtraMapsT<extra_kind(FEAT)> tmPeers, const LinearParams p) {
  constexpr bool AR = (FEAT & F_AR) != 0;
  constexpr bool kSK = (FEAT & F_SK) != 0, kGROUP = (FEAT & F_GROUP) != 0, kSWIGLU = (FEAT & F_SWIGLU) != 0;
  constexpr bool kROPE = (FEAT & F_ROPE) != 0, kDEQ16 = (FEAT & F_DEQ16) != 0, kOUT_ANY = (FEAT & F_OUT_ANY) != 0;
  constexpr bool kLoop = (FEAT & (F_DEQ16 | F_OUT_ANY | F_SK | F_AR)) != 0;  // the generic per-group epilogue loop
  constexpr bool kRESID = (FEAT & F_RESID) != 0;
  constexpr bool kNVLS = (FEAT & F_NVLS) != 0;
  using Cfg = TileCfg<CG>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  const uint32_t bar_base = base + Cfg::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + Cfg::BAR_OFFSET + 8u * (2 * STAGES + 4));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  static_assert(MC == 1 || CG == 2, "multicast clusters are built from CTA pairs");
  const uint32_t cluster_rank = (CG * MC > 1) ? cluster_ctarank() : 0u;
  const uint32_t cta_rank = cluster_rank & (CG - 1);              // 0 = leader of the pair
  const int pair_idx = static_cast<int>(cluster_rank) / CG;       // which pair of the cluster (MC == 2)
  // CTAs, CTA pairs or 4-CTA clusters that compute tiles (an NVLS launch appends reducer CTAs, which compute none)
  const int num_workers = (static_cast<int>(gridDim.x) - ((FEAT & F_NVLS) != 0 ? p.nvls_reducers : 0)) / (CG * MC);
  const int worker = blockIdx.x / (CG * MC);
  const bool fused = (FEAT & F_PHASE1) != 0 && (p.x != nullptr);
  if (threadIdx.x == 0) ASQ_STAMP(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmBu);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);   // the leader's arrive.expect_tx; TMA bytes of both CTAs land here
      mbar_init(empty_bar(s), MC);  // one tcgen05.commit per pair that reads this slot (multicast to its CTAs)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), NUM_EPI_WARPS * CG);  // epilogue warps of every CTA of the pair
    }
    if ((FEAT & F_RESID) != 0 && (FEAT & F_AR) == 0)
      for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(bar_base + 8u * (2 * STAGES + 5 + w), 1);  // residual tile landed
    mbar_fence_init();
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
    else         { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG * MC > 1) cluster_sync_all(); else __syncthreads();  // peer barriers must exist before remote arrives
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // barriers, TMEM and descriptors are set up: let the next kernel of the stream start ITS set-up on SMs we
  // free, and wait for our predecessor's results before touching global memory
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) ASQ_STAMP(1);
  if constexpr (kNVLS) {  // clear the counter bank of the NEXT launch (see LinearParams::nvls_ctr)
    for (int i = blockIdx.x * NUM_THREADS + threadIdx.x; i < p.nvls_ctr_words; i += gridDim.x * NUM_THREADS) p.nvls_ctr_next[i] = 0u;
  }

  if (kNVLS && worker >= num_workers) {
    // ===================== NVLS reducer CTA: in-switch sums of the tiles this rank owns =====================
    const int rwarp = (static_cast<int>(blockIdx.x) - num_workers * CG) * (NUM_THREADS / 32) + warp;
    const int num_rwarps = p.nvls_reducers * (NUM_THREADS / 32);
    if (p.y_dtype == ASQ_BF16) nvls_reducer_warp<true, CG>(p, rwarp, num_rwarps, lane);
    else                       nvls_reducer_warp<false, CG>(p, rwarp, num_rwarps, lane);
  } else if (warp == 0) {
    // ===================== TMA producer (one thread per CTA) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      TileWalk<FEAT> walk(p, worker, num_workers);
      Seg sg;
      while (walk.next(sg)) {
        const int m_blk = sg.m_blk;
        // MC == 2: the walk hands out 512-wide super tiles, this pair takes its 256-column half (a half that
        // lies beyond N still runs, on zero-filled W, because its A multicasts feed the other pair)
        const int col0 = sg.col0 + pair_idx * TILE_N;
        int width = (MC == 2) ? min(TILE_N / UNIT_N, p.n_units - col0 / UNIT_N) * UNIT_N : sg.width;
        if (width <= 0) width = UNIT_N;
        const int row0 = m_blk * Cfg::TILE_M + static_cast<int>(cta_rank) * BLOCK_M;  // this CTA's A rows
        const int b_rows = width / CG;                                                  // this CTA's W rows
        const int w_row = sg.group * p.N + col0 + static_cast<int>(cta_rank) * b_rows;  // group > 0: stacked expert weights
        const uint32_t stage_tx = CG * (Cfg::A_BYTES + static_cast<uint32_t>(b_rows) * BLOCK_K);
        bool panel_ready = !fused || row0 >= p.M;
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), stage_tx);
          const uint32_t sA = base + stage * Cfg::A_BYTES;
          const uint32_t sB = base + STAGES * Cfg::A_BYTES + stage * Cfg::B_BYTES;
          // W does not depend on phase 1: issue it before (possibly) waiting for the A panel
          if (width == TILE_N) {
            if (CG == 2) tma_load_2d_pair(sB, &tmB, full_bar(stage), kb * BLOCK_K, w_row);
            else         tma_load_2d(sB, &tmB, full_bar(stage), kb * BLOCK_K, w_row);
          } else {
            for (int u = 0; u * UNIT_N < width; ++u) {  // narrow tile: one 64-column unit per load
              const uint32_t dst = sB + u * (Cfg::UNIT_ROWS * BLOCK_K);
              const int r = w_row + u * static_cast<int>(Cfg::UNIT_ROWS);
              if (CG == 2) tma_load_2d_pair(dst, &tmBu, full_bar(stage), kb * BLOCK_K, r);
              else         tma_load_2d(dst, &tmBu, full_bar(stage), kb * BLOCK_K, r);
            }
          }
          if (!panel_ready) {
            const uint32_t need = static_cast<uint32_t>(min(BLOCK_M, p.M - row0));
            const uint32_t* flag = p.sync + 1 + row0 / BLOCK_M;
            while (ld_acquire_gpu(flag) < need) __nanosleep(64);
            fence_proxy_async_all();  // phase-1 generic-proxy stores -> TMA (async proxy) reads
            panel_ready = true;
            if (first) ASQ_STAMP(3);
          }
          if (MC == 2) {
            // fetch rows [pair_idx*64, +64) of this CTA's 128 and deliver them to both pairs' same-ranked CTAs
            const uint16_t mask = static_cast<uint16_t>((1u << cta_rank) | (1u << (cta_rank + CG)));
            tma_load_2d_pair_mcast(sA + pair_idx * (Cfg::A_BYTES / 2), &tmA, full_bar(stage), kb * BLOCK_K,
                                   row0 + pair_idx * (BLOCK_M / 2), mask);
          } else if (CG == 2) {
            tma_load_2d_pair(sA, &tmA, full_bar(stage), kb * BLOCK_K, row0);
          } else {
            tma_load_2d(sA, &tmA, full_bar(stage), kb * BLOCK_K, row0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc_base = make_idesc(FP8, BLOCK_M * CG, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      TileWalk<FEAT> walk(p, worker, num_workers);
      Seg sg;
      for (; walk.next(sg); ++it) {
        int width = sg.width;
        if (MC == 2) {
          width = min(TILE_N / UNIT_N, p.n_units - (sg.col0 + pair_idx * TILE_N) / UNIT_N) * UNIT_N;
          if (width <= 0) width = UNIT_N;
        }
        const uint32_t idesc = idesc_base | (static_cast<uint32_t>(width >> 3) << 17);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogues (both CTAs) drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TILE_N;
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (it == 0 && kb == sg.kb0) ASQ_STAMP(4);
          const uint64_t adesc = make_smem_desc_sw128(base + stage * Cfg::A_BYTES);
          const uint64_t bdesc = make_smem_desc_sw128(base + STAGES * Cfg::A_BYTES + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            const uint64_t koff = static_cast<uint64_t>(k * (UMMA_K >> 4));
            const uint32_t accum = (kb != sg.kb0) || (k != 0);
            if (CG == 2) {
              if (FP8) mma_f8_pair(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
              else     mma_i8_pair(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
            } else {
              if (FP8) mma_f8(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
              else     mma_i8(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
            }
          }
          // smem slot reusable once these MMAs retire: in both CTAs of the pair, and (MC == 2) in the other
          // pair too, whose multicast loads write into this pair's slot
          if (CG == 2) mma_commit_pair(empty_bar(stage), static_cast<uint16_t>((1u << (CG * MC)) - 1u));
          else         mma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue warps of both CTAs of this pair
        if (CG == 2) mma_commit_pair(tfull_bar(acc), static_cast<uint16_t>(3u << (pair_idx * CG)));
        else         mma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ===================== phase 1: activation quantisation =====================
    const int ew = warp - 2;  // 0..7
    if (fused) {
      const int warps_total = NUM_EPI_WARPS * gridDim.x;
      const bool tensor_dyn = (p.act_mode == ASQ_ACT_PER_TENSOR_DYNAMIC);
      const float ts = tensor_dyn ? tensor_scale_phase(p, ew * gridDim.x + blockIdx.x, warps_total, warps_total, lane) : 0.f;
      for (int row = ew * gridDim.x + blockIdx.x; row < p.M; row += warps_total) {
        if (kGROUP && p.group_of_blk != nullptr && __ldg(p.group_of_blk + row / BLOCK_M) < 0) continue;  // padding block of a grouped launch
        const float s = quantize_row_any<FP8>(p, row, lane, ts);
        if (lane == 0 && (p.act_mode == ASQ_ACT_PER_TOKEN || p.act_mode == ASQ_ACT_ROW_SCALE_GIVEN || tensor_dyn)) {
          p.row_scale[row] = s;
          if (p.row_scale_out != nullptr) p.row_scale_out[row] = s;
        }
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) red_release_gpu_add(p.sync + 1 + row / BLOCK_M, 1u);  // release: orders the row's stores
      }
      if (ew == 0 && lane == 0) ASQ_STAMP(2);
    }
    // ===================== phase 2: epilogue =====================
    // A tile is width/64 column groups; warps with half == 0 take the even groups, half == 1 the odd ones.
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int half = ew >> 2;
    const uint32_t tempty_leader0 = (CG == 2) ? mapa_shared(tempty_bar(0), pair_idx * CG) : tempty_bar(0);
    const uint32_t tempty_leader1 = (CG == 2) ? mapa_shared(tempty_bar(1), pair_idx * CG) : tempty_bar(1);
    const uint32_t stage_base = base + Cfg::EPI_OFFSET + ew * (EPI_BUF_BYTES * EPI_NBUF);
    const bool out16 = (p.y_dtype == ASQ_BF16 || p.y_dtype == ASQ_F16);
    const bool staged = p.tma_store != 0;
    const int elem = out16 ? 2 : (p.y_dtype == ASQ_I8 ? 1 : 4);
    const bool per_token_epi = (p.act_mode == ASQ_ACT_PER_TOKEN || p.act_mode == ASQ_ACT_ROW_SCALE_GIVEN ||
                                p.act_mode == ASQ_ACT_PER_TENSOR_DYNAMIC) &&
                               (p.epi_kind == EPI_DEQUANT || p.epi_kind == EPI_SWIGLU);
    uint32_t gcount = 0;  // staging tiles issued by this warp (buffer = gcount & 1)
    uint32_t resid_phase = 0;  // parity of this warp's residual-tile barrier
    // all-reduce mode: this launch's epoch = 1 + the epoch of the last launch that finished on this rank
    const uint32_t ar_epoch = (AR && p.ar_world > 1) ? __ldcg(p.ar_ctl[p.ar_rank]) + 1u : 0u;
    int it = 0;
    TileWalk<FEAT> walk(p, worker, num_workers);
    Seg sg;
    for (; walk.next(sg); ++it) {
      const int m_blk = sg.m_blk;
      const int tile_col0 = sg.col0 + pair_idx * TILE_N;
      int width = sg.width;
      if (MC == 2) width = max(0, min(TILE_N / UNIT_N, p.n_units - tile_col0 / UNIT_N)) * UNIT_N;  // 0: nothing to store
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int row0 = m_blk * Cfg::TILE_M + static_cast<int>(cta_rank) * BLOCK_M + quad * 32;
      const int row = row0 + lane;
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * TILE_N;
      const int ngroups = width / UNIT_N;
      float rs = 0.f;
      const float tile_scale = (kGROUP && p.group_scale != nullptr) ? __ldg(p.group_scale + sg.group) : p.dequant_scale;
      const float tile_scale_up = (kGROUP && p.group_scale_up != nullptr) ? __ldg(p.group_scale_up + sg.group) : p.dequant_scale_up;
      // stream-K owner: the workers after this one hold the rest of this tile's K range
      int n_contrib = 0;
      if (kSK && sg.role == SEG_OWNER) {
        const long long tile_end = static_cast<long long>(sg.tile_g0) + p.num_k_blocks;
        for (int c = worker + 1; c < p.sk_workers && TileWalk<FEAT>::sk_begin(p, c) < tile_end; ++c) ++n_contrib;
        if (lane == 0) {
          for (int c = 0; c < n_contrib; ++c) {
            const uint32_t* f = sk_flag(p, worker + 1 + c, CG, cta_rank, ew);
            while (ld_acquire_gpu(f) == 0u) __nanosleep(64);
          }
        }
        __syncwarp();
      }
      if (AR && sg.role == SEG_AR_OWNER) {  // the other ranks' partials of this tile must have landed in our buffer
        if (lane == 0) {
          for (int sl = 0; sl < p.ar_world - 1; ++sl) {
            const uint32_t* f = p.ar_ctl[p.ar_rank] + AR_FLAG_BASE + ar_index(p, sl, sg.ar_tile / p.ar_world, CG, cta_rank, ew);
            while (ld_acquire_sys(f) != ar_epoch) __nanosleep(100);
          }
        }
        __syncwarp();
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (it == 0 && ew == 0 && lane == 0) ASQ_STAMP(5);
      if (per_token_epi && row < p.M) rs = __ldcg(p.row_scale + row);
      if (kSWIGLU && p.epi_kind == EPI_SWIGLU) {
        // Interleaved gate|up tile: group g = 32 gate columns + the 32 matching up columns -> 32 outputs.  This
        // warp takes groups 2*half and 2*half+1, i.e. 64 adjacent output columns of its 32 rows: one staging
        // tile (int8: 32 rows x 64 B, dense; 16-bit: 32 rows x 128 B, swizzled) and one TMA store per tile.
        const int g_first = 2 * half;
        if (g_first < ngroups) {
          const uint32_t buf = stage_base + (gcount % EPI_NBUF) * EPI_BUF_BYTES;
          if (gcount >= EPI_NBUF) {
            if (lane == 0) tma_store_wait_read<EPI_NBUF - 1>();
            __syncwarp();
          }
#pragma unroll 1
          for (int gi = 0; gi < 2 && g_first + gi < ngroups; ++gi) {
            const int g = g_first + gi;
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(taddr0 + g * UNIT_N, r0);
            tmem_ld_32x32(taddr0 + g * UNIT_N + 32, r1);
            tmem_ld_wait();
            float vg[32], vu[32];
            epilogue_values<FP8>(r0, vg, tile_col0 + g * UNIT_N, rs, p, tile_scale);
            epilogue_values<FP8>(r1, vu, tile_col0 + g * UNIT_N + 32, rs, p, tile_scale_up);
            uint32_t w[16];
            swiglu_dispatch(vg, vu, w, p);
            if (out16) {
              stage_words<16>(buf, lane, gi * 4, w);
            } else {
              const uint32_t addr = buf + lane * 64 + gi * 32;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 16), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < p.M) {
            tma_store_2d(&tmY, buf, ((tile_col0 >> 1) + g_first * 32) * elem, row0);
            tma_store_commit();
          }
          ++gcount;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
          else         mbar_arrive(tempty_bar(acc));
        }
        continue;
      }
      if (kROPE && p.rope_cos != nullptr) {
        // Fused q|k|v projection with RoPE: this warp takes groups 2*half and 2*half+1 = one whole 128-wide
        // head of its 32 rows, so both operands of the rotation (columns d and d + 64) are in its registers.
        // Both staging buffers are used per tile (one per 64-column group), then two TMA stores.
        const int g0 = 2 * half;
        if (g0 < ngroups) {
          if (gcount > 0) {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
          const uint32_t buf_a = stage_base, buf_b = stage_base + EPI_BUF_BYTES;
          const int colh = tile_col0 + g0 * UNIT_N;
          const bool rot = colh < p.rope_cols;
          const bool bf = (p.y_dtype == ASQ_BF16);
          const size_t blk = static_cast<size_t>(p.rope_S) * 8;  // elements between 8-column table blocks
          const size_t tab = static_cast<size_t>(row % p.rope_S) * 8;
          const uint16_t* cos_row = reinterpret_cast<const uint16_t*>(p.rope_cos) + tab;
          const uint16_t* sin_row = reinterpret_cast<const uint16_t*>(p.rope_sin) + tab;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t ra[32], rb[32];
            tmem_ld_32x32(taddr0 + g0 * UNIT_N + c * 32, ra);
            tmem_ld_32x32(taddr0 + g0 * UNIT_N + UNIT_N + c * 32, rb);
            tmem_ld_wait();
            float va[32], vb[32];
            epilogue_values<FP8>(ra, va, colh + c * 32, rs, p, tile_scale);
            epilogue_values<FP8>(rb, vb, colh + UNIT_N + c * 32, rs, p, tile_scale);
            uint32_t wa[16], wb[16];
            if (bf) rope_chunk<true>(va, vb, wa, wb, rot, cos_row + c * 4 * blk, sin_row + c * 4 * blk, blk, p.rope_halves_equal != 0);
            else    rope_chunk<false>(va, vb, wa, wb, rot, cos_row + c * 4 * blk, sin_row + c * 4 * blk, blk, p.rope_halves_equal != 0);
            stage_words<16>(buf_a, lane, c * 4, wa);
            stage_words<16>(buf_b, lane, c * 4, wb);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < p.M) {
            tma_store_2d(&tmY, buf_a, colh * elem, row0);
            if (colh + UNIT_N < p.N) tma_store_2d(&tmY, buf_b, (colh + UNIT_N) * elem, row0);
            tma_store_commit();
          }
          gcount += 2;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
          else         mbar_arrive(tempty_bar(acc));
        }
        continue;
      }
      int chunk_pair = 0;  // index of the (r0, r1) pair inside this warp's stream-K region
#pragma unroll 1
      for (int g = half; kLoop && g < ngroups; g += 2, ++chunk_pair) {
        const uint32_t taddr = taddr0 + g * UNIT_N;
        const int col0 = tile_col0 + g * UNIT_N;
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(taddr, r0);
        tmem_ld_32x32(taddr + 32, r1);
        if (kSK && sg.role == SEG_CONTRIB) {
          // raw partial accumulators -> scratch, coalesced: uint4 index (chunk*8 + quad16)*32 + lane
          tmem_ld_wait();
          uint4* dst = reinterpret_cast<uint4*>(sk_warp_base(p, worker, CG, cta_rank, ew)) + chunk_pair * 2 * 8 * 32 + lane;
#pragma unroll
          for (int q = 0; q < 8; ++q) __stcg(dst + q * 32, make_uint4(r0[4 * q], r0[4 * q + 1], r0[4 * q + 2], r0[4 * q + 3]));
#pragma unroll
          for (int q = 0; q < 8; ++q) __stcg(dst + (8 + q) * 32, make_uint4(r1[4 * q], r1[4 * q + 1], r1[4 * q + 2], r1[4 * q + 3]));
          continue;
        }
        if (AR && sg.role == SEG_AR_CONTRIB && p.ar_partial16) {
          // dequantised partial (+ bias on the rank that holds it), rounded to the output dtype: 2 bytes per element
          tmem_ld_wait();
          const int owner = sg.ar_tile % p.ar_world;
          const int slot = (p.ar_rank - owner - 1 + p.ar_world) % p.ar_world;
          uint4* dst = reinterpret_cast<uint4*>(p.ar_recv[owner] + ar_index(p, slot, sg.ar_tile / p.ar_world, CG, cta_rank, ew) * SK_WARP_WORDS) +
                       chunk_pair * 8 * 32 + lane;
          float v[32];
          uint32_t w[16];
          epilogue_values<FP8>(r0, v, col0, rs, p, tile_scale);
          pack_out16(v, w, p.y_dtype == ASQ_BF16);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q * 32] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
          epilogue_values<FP8>(r1, v, col0 + 32, rs, p, tile_scale);
          pack_out16(v, w, p.y_dtype == ASQ_BF16);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[(4 + q) * 32] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
          continue;
        }
        if (AR && sg.role == SEG_AR_CONTRIB) {
          // raw partial accumulators -> the owner's receive buffer over NVLink, coalesced 512-byte warp stores
          tmem_ld_wait();
          const int owner = sg.ar_tile % p.ar_world;
          const int slot = (p.ar_rank - owner - 1 + p.ar_world) % p.ar_world;
          uint4* dst = reinterpret_cast<uint4*>(p.ar_recv[owner] + ar_index(p, slot, sg.ar_tile / p.ar_world, CG, cta_rank, ew) * SK_WARP_WORDS) +
                       chunk_pair * 2 * 8 * 32 + lane;
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q * 32] = make_uint4(r0[4 * q], r0[4 * q + 1], r0[4 * q + 2], r0[4 * q + 3]);
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[(8 + q) * 32] = make_uint4(r1[4 * q], r1[4 * q + 1], r1[4 * q + 2], r1[4 * q + 3]);
          continue;
        }
        if (AR && sg.role == SEG_AR_OWNER && !p.ar_partial16) {
          tmem_ld_wait();
          for (int sl = 0; sl < p.ar_world - 1; ++sl) {  // fixed source order: deterministic fp32 sums
            const uint4* src = reinterpret_cast<const uint4*>(p.ar_recv[p.ar_rank] + ar_index(p, sl, sg.ar_tile / p.ar_world, CG, cta_rank, ew) * SK_WARP_WORDS) +
                               chunk_pair * 2 * 8 * 32 + lane;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 a = __ldcg(src + q * 32);
              const uint4 b = __ldcg(src + (8 + q) * 32);
              if (FP8) {
                r0[4 * q] = __float_as_uint(__uint_as_float(r0[4 * q]) + __uint_as_float(a.x));
                r0[4 * q + 1] = __float_as_uint(__uint_as_float(r0[4 * q + 1]) + __uint_as_float(a.y));
                r0[4 * q + 2] = __float_as_uint(__uint_as_float(r0[4 * q + 2]) + __uint_as_float(a.z));
                r0[4 * q + 3] = __float_as_uint(__uint_as_float(r0[4 * q + 3]) + __uint_as_float(a.w));
                r1[4 * q] = __float_as_uint(__uint_as_float(r1[4 * q]) + __uint_as_float(b.x));
                r1[4 * q + 1] = __float_as_uint(__uint_as_float(r1[4 * q + 1]) + __uint_as_float(b.y));
                r1[4 * q + 2] = __float_as_uint(__uint_as_float(r1[4 * q + 2]) + __uint_as_float(b.z));
                r1[4 * q + 3] = __float_as_uint(__uint_as_float(r1[4 * q + 3]) + __uint_as_float(b.w));
              } else {  // int32 partial sums: exact, so the result equals the unsharded GEMM bit for bit
                r0[4 * q] += a.x; r0[4 * q + 1] += a.y; r0[4 * q + 2] += a.z; r0[4 * q + 3] += a.w;
                r1[4 * q] += b.x; r1[4 * q + 1] += b.y; r1[4 * q + 2] += b.z; r1[4 * q + 3] += b.w;
              }
            }
          }
        }
        if (kSK && sg.role == SEG_OWNER) {
          tmem_ld_wait();
          for (int c = 0; c < n_contrib; ++c) {
            const uint4* src = reinterpret_cast<const uint4*>(sk_warp_base(p, worker + 1 + c, CG, cta_rank, ew)) + chunk_pair * 2 * 8 * 32 + lane;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 a = __ldcg(src + q * 32);
              const uint4 b = __ldcg(src + (8 + q) * 32);
              if (FP8) {
                r0[4 * q] = __float_as_uint(__uint_as_float(r0[4 * q]) + __uint_as_float(a.x));
                r0[4 * q + 1] = __float_as_uint(__uint_as_float(r0[4 * q + 1]) + __uint_as_float(a.y));
                r0[4 * q + 2] = __float_as_uint(__uint_as_float(r0[4 * q + 2]) + __uint_as_float(a.z));
                r0[4 * q + 3] = __float_as_uint(__uint_as_float(r0[4 * q + 3]) + __uint_as_float(a.w));
                r1[4 * q] = __float_as_uint(__uint_as_float(r1[4 * q]) + __uint_as_float(b.x));
                r1[4 * q + 1] = __float_as_uint(__uint_as_float(r1[4 * q + 1]) + __uint_as_float(b.y));
                r1[4 * q + 2] = __float_as_uint(__uint_as_float(r1[4 * q + 2]) + __uint_as_float(b.z));
                r1[4 * q + 3] = __float_as_uint(__uint_as_float(r1[4 * q + 3]) + __uint_as_float(b.w));
              } else {  // int32 partial sums: exact and order-independent
                r0[4 * q] += a.x; r0[4 * q + 1] += a.y; r0[4 * q + 2] += a.z; r0[4 * q + 3] += a.w;
                r1[4 * q] += b.x; r1[4 * q + 1] += b.y; r1[4 * q + 2] += b.z; r1[4 * q + 3] += b.w;
              }
            }
          }
        }
        if (kDEQ16 && staged && out16) {
          // 64 output columns (two TMEM chunks) fill one 128-byte wide staging tile
          const uint32_t buf = stage_base + (gcount % EPI_NBUF) * EPI_BUF_BYTES;
          if (gcount >= EPI_NBUF) {  // the store that last used this buffer must have read it
            if (lane == 0) tma_store_wait_read<EPI_NBUF - 1>();
            __syncwarp();
          }
          if constexpr (kRESID && !AR) {
            if (p.residual != nullptr) {
              const uint32_t rbar = bar_base + 8u * (2 * STAGES + 5 + ew);
              if (lane == 0) {
                mbar_arrive_expect_tx(rbar, EPI_BUF_BYTES);
                tma_load_2d(buf, &tmPeers.m[0], rbar, col0 * elem, row0);  // 32 rows x 64 columns of the residual
              }
              mbar_wait(rbar, resid_phase);
              resid_phase ^= 1u;
            }
          }
          tmem_ld_wait();
          float v[32];
          uint32_t w[16];
          const bool ar16_owner = (AR && sg.role == SEG_AR_OWNER && p.ar_partial16);
          const bool bf_out = (p.y_dtype == ASQ_BF16);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            epilogue_values<FP8>(h ? r1 : r0, v, col0 + h * 32, rs, p, tile_scale);
            if (ar16_owner) {
              // own partial rounded to the output dtype like everybody's, then the peers' 16-bit partials added in
              // fp32 in fixed source order; one final rounding in pack_out16 (world 2: exactly NCCL's bf16 sum)
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                if (bf_out) round_pair<true>(v[j], v[j + 1]); else round_pair<false>(v[j], v[j + 1]);
              }
              for (int sl = 0; sl < p.ar_world - 1; ++sl) {
                const uint4* src = reinterpret_cast<const uint4*>(p.ar_recv[p.ar_rank] + ar_index(p, sl, sg.ar_tile / p.ar_world, CG, cta_rank, ew) * SK_WARP_WORDS) +
                                   chunk_pair * 8 * 32 + h * 4 * 32 + lane;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 a = __ldcg(src + q * 32);
                  const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float lo, hi;
                    if (bf_out) unpack_pair<true>(aw[i], lo, hi); else unpack_pair<false>(aw[i], lo, hi);
                    v[8 * q + 2 * i] = __fadd_rn(v[8 * q + 2 * i], lo);
                    v[8 * q + 2 * i + 1] = __fadd_rn(v[8 * q + 2 * i + 1], hi);
                  }
                }
              }
            }
            if constexpr (kRESID && !AR) {
              if (p.residual != nullptr) {
                // residual add in the activation dtype: T(res + T(v)), the eager `residual + linear(x)`.  The tile was
                // TMA-loaded into this staging buffer (same 128B-swizzled layout as the output), so every lane
                // reads its row's 16-byte pieces from shared memory; out-of-range elements are zero-filled.
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint32_t aw[4];
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(aw[0]), "=r"(aw[1]), "=r"(aw[2]), "=r"(aw[3])
                               : "r"(buf + lane * 128 + (((h * 4 + q) ^ (lane & 7)) << 4)));
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float lo, hi, x0 = v[8 * q + 2 * i], x1 = v[8 * q + 2 * i + 1];
                    if (bf_out) { unpack_pair<true>(aw[i], lo, hi); round_pair<true>(x0, x1); }
                    else        { unpack_pair<false>(aw[i], lo, hi); round_pair<false>(x0, x1); }
                    v[8 * q + 2 * i] = __fadd_rn(lo, x0);
                    v[8 * q + 2 * i + 1] = __fadd_rn(hi, x1);
                  }
                }
              }
            }
            pack_out16(v, w, bf_out);
            stage_words<16>(buf, lane, h * 4, w);
          }
          if (AR && p.ar_y_mc != nullptr && sg.role == SEG_AR_OWNER) {
            // NVLS: read the staged tile back row-major (8 lanes x 16 bytes = one 128-byte row) and multicast it
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + (lane >> 3), c = lane & 7;
              uint32_t x0, x1, x2, x3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                           : "r"(buf + r * 128 + ((c ^ (r & 7)) << 4)));
              const int grow = row0 + r, gcol = col0 + c * 8;
              if (grow < p.M && gcol < p.N) {
                uint8_t* dst = p.ar_y_mc + (static_cast<size_t>(grow) * p.N + gcol) * 2;
                asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(x0)),
                             "f"(__uint_as_float(x1)), "f"(__uint_as_float(x2)), "f"(__uint_as_float(x3)) : "memory");
              }
            }
            __syncwarp();  // the buffer may be rewritten by the next group: every lane has read its part
          } else {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && row0 < p.M && col0 < p.N) {
              tma_store_2d(&tmY, buf, col0 * elem, row0);
              if constexpr (AR) {
                for (int pr = 0; pr < p.ar_world - 1; ++pr) tma_store_2d(&tmPeers.m[pr], buf, col0 * elem, row0);  // all-gather by stores
              }
              tma_store_commit();
            }
          }
          ++gcount;
        } else if (kOUT_ANY && staged) {
          // 4-byte outputs: each TMEM chunk (32 columns) is one 128-byte wide staging tile
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t buf = stage_base + (gcount % EPI_NBUF) * EPI_BUF_BYTES;
            if (gcount >= EPI_NBUF) {
              if (lane == 0) tma_store_wait_read<EPI_NBUF - 1>();
              __syncwarp();
            }
            const int c0 = col0 + h * 32;
            if (p.epi_kind == EPI_RAW_I32) {
              stage_words<32>(buf, lane, 0, h ? r1 : r0);
            } else {
              float v[32];
              uint32_t w[32];
              epilogue_values<FP8>(h ? r1 : r0, v, c0, rs, p, tile_scale);
#pragma unroll
              for (int j = 0; j < 32; ++j)
                w[j] = (p.y_dtype == ASQ_I32) ? static_cast<uint32_t>(__float2int_rn(v[j])) : __float_as_uint(v[j]);
              stage_words<32>(buf, lane, 0, w);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && row0 < p.M && c0 < p.N) {
              tma_store_2d(&tmY, buf, c0 * elem, row0);
              tma_store_commit();
            }
            ++gcount;
          }
        } else if (kOUT_ANY) {
          tmem_ld_wait();
          store_chunk_direct<FP8>(r0, row, col0, rs, p);
          store_chunk_direct<FP8>(r1, row, col0 + 32, rs, p);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
        else         mbar_arrive(tempty_bar(acc));
        if (kSK && sg.role == SEG_CONTRIB) st_release_gpu(sk_flag(p, worker, CG, cta_rank, ew), 1u);  // partial published
        if (AR && sg.role == SEG_AR_CONTRIB) {  // __syncwarp above ordered the other lanes' stores before this fence
          const int owner = sg.ar_tile % p.ar_world;
          const int slot = (p.ar_rank - owner - 1 + p.ar_world) % p.ar_world;
          fence_acq_rel_sys();
          st_release_sys(p.ar_ctl[owner] + AR_FLAG_BASE + ar_index(p, slot, sg.ar_tile / p.ar_world, CG, cta_rank, ew), ar_epoch);
        }
        if (kSK && sg.role == SEG_OWNER)
          for (int c = 0; c < n_contrib; ++c) *sk_flag(p, worker + 1 + c, CG, cta_rank, ew) = 0u;  // consumed: re-arm
      }
      if constexpr (kNVLS) {
        // This warp's part of the tile's partial is on its way into this rank's buffer: once it has LANDED (bulk
        // stores complete, not merely read from shared memory) bump the tile's counter on its owner.  The accumulator
        // was handed back above, so the next tile's MMAs are already running.
        if (lane == 0) {
          const unsigned long long ts0 = (p.dbg != nullptr && ew == 0) ? globaltimer_ns() : 0ull;
          tma_store_wait_all();
          fence_proxy_async_all();  // async-proxy (TMA) writes before the generic-proxy release below
          // Release at GPU scope: the partial tile and the counter live in THIS GPU's memory, whose L2 is the point of
          // coherence the switch reads through.  A system-scope fence here waits behind all multimem traffic the
          // reducer CTAs have in flight: measured 22 us per tile (profiles/r02_allreduce.md), 30x the tile's math.
          const int slab = static_cast<int>(cta_rank) * (BLOCK_M / NVLS_SLAB_ROWS) + (quad >> 1);
          red_release_gpu_add(p.nvls_ctr + static_cast<size_t>(sg.ar_tile) * (CG * BLOCK_M / NVLS_SLAB_ROWS) + slab, 1u);
          if (p.dbg != nullptr && ew == 0) p.dbg[blockIdx.x * 8 + 2] += globaltimer_ns() - ts0;  // ns spent signalling
        }
        __syncwarp();
      }
    }
    if (lane == 0) tma_store_wait_all();  // staged tiles fully written before the CTA retires
    if (ew == 0 && lane == 0) ASQ_STAMP(6);
  }

  tc_fence_before();
  if (CG * MC > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (threadIdx.x == 0) ASQ_STAMP(7);
  if ((AR || kNVLS) && p.ar_world > 1 && threadIdx.x == 0) {
    // This CTA's tiles are stored everywhere (each epilogue warp waited for its TMA stores before the sync above).
    // The last CTA of the rank tells every peer "rank r finished launch e", waits for the same word from every
    // peer (their stores into OUR y and their reads of OUR partials are then complete) and advances the epoch.
    uint32_t* ctl = p.ar_ctl[p.ar_rank];
    const uint32_t epoch = __ldcg(ctl) + 1u;
    // stores into other ranks' memory (peer TMA stores / multimem.st) must be performed system-wide before this CTA
    // counts as finished; the GEMM CTAs of an NVLS launch wrote local memory only
    if (!kNVLS || worker >= num_workers) __threadfence_system();
    else __threadfence();
    if (kNVLS && worker >= num_workers) ASQ_STAMP(1);  // reducer CTA: its multimem stores are performed system-wide
    const uint32_t done = atomicAdd(ctl + 1, 1u);
    if (done == gridDim.x - 1) {
      // timeline of the handshake (last CTA), slots of a virtual CTA after the grid: 0 entered, 1 peers told, 2 peers heard, 3 done
      unsigned long long* hs = p.dbg != nullptr ? p.dbg + static_cast<size_t>(gridDim.x) * 8 : nullptr;
      if (hs != nullptr) hs[0] = globaltimer_ns();
      ctl[1] = 0u;
      __threadfence_system();
      for (int r = 0; r < p.ar_world; ++r)
        if (r != p.ar_rank) st_release_sys(p.ar_ctl[r] + 2 + p.ar_rank, epoch);
      if (hs != nullptr) hs[1] = globaltimer_ns();
      for (int r = 0; r < p.ar_world; ++r)
        if (r != p.ar_rank)
          while (ld_acquire_sys(ctl + 2 + r) != epoch) __nanosleep(200);
      if (hs != nullptr) hs[2] = globaltimer_ns();
      ctl[0] = epoch;
      __threadfence_system();
      if (hs != nullptr) hs[3] = globaltimer_ns();
    }
  }
  if (fused && threadIdx.x == 0) {
    // last CTA out restores the phase counters so the workspace is reusable by the next launch
    __threadfence();
    const uint32_t done = atomicAdd(p.sync + SYNC_EXIT, 1u);
    if (done == gridDim.x - 1) {
      const int panels = (p.M + BLOCK_M - 1) / BLOCK_M;
      for (int i = 0; i < panels; ++i) p.sync[1 + i] = 0u;
      p.sync[SYNC_AMAX] = 0u;
      p.sync[SYNC_AMAX_CNT] = 0u;
      p.sync[SYNC_EXIT] = 0u;
      __threadfence();
    }
  }
}

// Stand-alone prologue (debug / parity tap): same row routine, no GEMM.
template <bool FP8>
__global__ void __launch_bounds__(256) asq_quantize_kernel(const LinearParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < p.M; row += warps_total) {
    const float s = quantize_row_any<FP8>(p, row, lane);
    if (lane == 0 && p.act_mode == ASQ_ACT_PER_TOKEN && p.row_scale != nullptr) p.row_scale[row] = s;
  }
}

}  // namespace asq

int asq_glue_fail(int code, const char* fmt, ...);  // formats into the thread-local error buffer (host TU)
bool asq_pdl_enabled();                             // ASQ_PDL=0 turns programmatic dependent launch off (host TU)

#if !defined(ASQ_TU) || ASQ_TU == 0
// ====================================================================== host side
namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace

bool asq_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("ASQ_PDL"); on = (e != nullptr && e[0] == '1') ? 1 : 0; }
  return on == 1;
}

// error reporting shared with asq_glue.cu
int asq_glue_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

struct DeviceState {
  bool probed = false;
  bool supported = false;
  int sm_count = 0;
  bool attrs_set = false;
};
DeviceState g_dev[64];
std::mutex g_mu;
EncodeTiledFn g_encode = nullptr;

int get_device(int* dev_out, DeviceState** st_out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ASQ_ERR_CUDA, "no CUDA device available: %s", cudaGetErrorString(e));
  }
  if (dev < 0 || dev >= 64) return fail(ASQ_ERR_CUDA, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_mu);
  DeviceState& st = g_dev[dev];
  if (!st.probed) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return fail(ASQ_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    st.supported = (prop.major == 10 && prop.minor == 0);
    st.sm_count = prop.multiProcessorCount;
    st.probed = true;
  }
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
      return fail(ASQ_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *dev_out = dev;
  *st_out = &st;
  return ASQ_OK;
}

// 2-D map over a row-major [rows, K] byte matrix, box = [box_rows, 128 bytes], 128B swizzle.
int make_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t K, int box_rows) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(K)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(asq::BLOCK_K), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim, gstride, box,
                        estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ASQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return ASQ_OK;
}

// Dense (unswizzled) map over a row-major [rows, row_bytes] byte matrix, box = [box_rows, box_bytes].
int make_tmap_plain(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t row_bytes, int box_bytes, int box_rows) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(row_bytes), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(row_bytes)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_bytes), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim, gstride, box,
                        estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ASQ_ERR_CUDA, "cuTensorMapEncodeTiled (plain) failed (%d)", static_cast<int>(r));
  return ASQ_OK;
}

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  uint8_t* a_q;
  float* row_scale;
  uint32_t* sync;
  uint32_t* sk_flags;
  uint32_t* sk_partial;
};
// Layout: [phase counters: kSyncBytes] [stream-K handshake words: kSkFlagBytes] [stream-K partial accumulators:
// kSkPartialBytes] [row scales: M fp32] [8-bit copy of x: M*K].  Everything the kernels expect to find zeroed
// sits at fixed offsets, so a cached, zero-initialised buffer stays valid when M and K change.
constexpr size_t kSyncBytes = 32768;  // 8192 counters -> up to 8189 panels (about 1M rows)
constexpr int kSkMaxCtas = 160;       // >= SM count of any sm_100 part
constexpr size_t kSkFlagBytes = 8192; // kSkMaxCtas * 8 warps * 4 bytes, rounded up
constexpr size_t kSkPartialBytes = static_cast<size_t>(kSkMaxCtas) * asq::SK_SLOT_WORDS * 4;  // 20 MB
constexpr size_t kFixedBytes = kSyncBytes + kSkFlagBytes + kSkPartialBytes;
size_t ws_layout(int64_t M, int64_t K, void* base, Workspace* w) {
  const size_t rs = round_up(static_cast<size_t>(M) * 4, 1024);
  const size_t aq = round_up(static_cast<size_t>(M) * K, 1024);
  if (w != nullptr) {
    uint8_t* b = static_cast<uint8_t*>(base);
    w->sync = reinterpret_cast<uint32_t*>(b);
    w->sk_flags = reinterpret_cast<uint32_t*>(b + kSyncBytes);
    w->sk_partial = reinterpret_cast<uint32_t*>(b + kSyncBytes + kSkFlagBytes);
    w->row_scale = reinterpret_cast<float*>(b + kFixedBytes);
    w->a_q = b + kFixedBytes + rs;
  }
  return kFixedBytes + rs + aq;
}

}  // namespace
#endif  // ASQ_TU host part

// ---------------------------------------------------------------------- kernel launchers
// The six instantiations of the kernel dominate the build time, so the build compiles this file once per
// instantiation in parallel (-DASQ_TU=1..22: only the kernel + its launcher) plus once for the host side
// (-DASQ_TU=0: everything else, launchers declared `extern template`).  Without ASQ_TU it is one ordinary TU.
namespace asq_launch {
constexpr int kMaxDevices = 64;

template <bool FP8, int CG, int MC, int FEAT>
int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBu, const CUtensorMap& tmY,
               const asq::ExtraMapsT<asq::extra_kind(FEAT)>& tmPeers, const asq::LinearParams& p, int workers,
               cudaStream_t stream) {
  using Cfg = asq::TileCfg<CG>;
  const int extra_ctas = (FEAT & asq::F_NVLS) != 0 ? p.nvls_reducers : 0;  // reducer CTAs appended to the grid
  auto kern = asq::asq_linear_kernel<FP8, CG, MC, FEAT>;
  cudaError_t e;
  {
    static bool attr_set[kMaxDevices] = {};  // per template instantiation, per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev]) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
      if (e != cudaSuccess) return asq_glue_fail(ASQ_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      attr_set[dev] = true;
    }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(workers * CG * MC + extra_ctas), 1, 1);
  cfg.blockDim = dim3(asq::NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG * MC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = asq_pdl_enabled() ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmBu, tmY, tmPeers, p);
  if (e != cudaSuccess) return asq_glue_fail(ASQ_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return ASQ_OK;
}

// How many 4-CTA clusters of the multicast variant can be co-resident (GPC packing strands a few SMs, e.g.
// 33 clusters = 132 of 148 SMs); the persistent grid must not exceed it.  Cached per device; 0 = unusable.
template <bool FP8>
int max_multicast_clusters(int dev) {
  static int cached[kMaxDevices];
  static bool known[kMaxDevices] = {};
  if (known[dev]) return cached[dev];
  using Cfg = asq::TileCfg<2>;
  auto kern = asq::asq_linear_kernel<FP8, 2, 2, asq::F_FULL>;
  int n = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) == cudaSuccess) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(4 * 64, 1, 1);
    cfg.blockDim = dim3(asq::NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  } else {
    cudaGetLastError();
  }
  cached[dev] = n;
  known[dev] = true;
  return n;
}

#if defined(ASQ_TU)
#if ASQ_TU == 0
#define ASQ_LAUNCH_DECL extern template
#else
#define ASQ_LAUNCH_DECL template
#endif
#define ASQ_LAUNCH_INST(FP8, CG, MC, FEAT)                                                                      \
  ASQ_LAUNCH_DECL int launch_cfg<FP8, CG, MC, FEAT>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,  \
                                                    const CUtensorMap&,                                          \
                                                    const asq::ExtraMapsT<asq::extra_kind(FEAT)>&,                \
                                                    const asq::LinearParams&, int, cudaStream_t);
// lean instantiations for the launches of a Llama layer (int8, CTA pairs): see asq::Feat
#define ASQ_LEAN_PLAIN (asq::F_DEQ16)
#define ASQ_LEAN_PHASE1 (asq::F_DEQ16 | asq::F_PHASE1)
#define ASQ_LEAN_SWIGLU (asq::F_SWIGLU)
#define ASQ_LEAN_ROPE (asq::F_ROPE)
#define ASQ_LEAN_AR (asq::F_AR | asq::F_DEQ16)
#define ASQ_LEAN_PLAIN_SK (asq::F_DEQ16 | asq::F_SK)
#define ASQ_LEAN_PHASE1_SK (asq::F_DEQ16 | asq::F_PHASE1 | asq::F_SK)
#define ASQ_LEAN_PLAIN_RESID (asq::F_DEQ16 | asq::F_RESID)
#define ASQ_LEAN_PHASE1_RESID (asq::F_DEQ16 | asq::F_PHASE1 | asq::F_RESID)
#define ASQ_LEAN_PHASE1_ROPE (asq::F_ROPE | asq::F_PHASE1)
#define ASQ_LEAN_PHASE1_SWIGLU (asq::F_SWIGLU | asq::F_PHASE1)
#define ASQ_LEAN_NVLS (asq::F_NVLS | asq::F_DEQ16)
#if ASQ_TU == 0 || ASQ_TU == 1
ASQ_LAUNCH_INST(false, 1, 1, asq::F_FULL)
#endif
#if ASQ_TU == 0 || ASQ_TU == 2
ASQ_LAUNCH_INST(false, 2, 1, asq::F_FULL)
#endif
#if ASQ_TU == 0 || ASQ_TU == 3
ASQ_LAUNCH_INST(true, 1, 1, asq::F_FULL)
#endif
#if ASQ_TU == 0 || ASQ_TU == 4
ASQ_LAUNCH_INST(true, 2, 1, asq::F_FULL)
#endif
#if defined(ASQ_ENABLE_MC) && (ASQ_TU == 0 || ASQ_TU == 5)
ASQ_LAUNCH_INST(false, 2, 2, asq::F_FULL)
ASQ_LAUNCH_DECL int max_multicast_clusters<false>(int);
#endif
#if defined(ASQ_ENABLE_MC) && (ASQ_TU == 0 || ASQ_TU == 6)
ASQ_LAUNCH_INST(true, 2, 2, asq::F_FULL)
ASQ_LAUNCH_DECL int max_multicast_clusters<true>(int);
#endif
#if ASQ_TU == 0 || ASQ_TU == 7
ASQ_LAUNCH_INST(false, 1, 1, asq::F_FULL | asq::F_AR)
#endif
#if ASQ_TU == 0 || ASQ_TU == 8
ASQ_LAUNCH_INST(false, 2, 1, asq::F_FULL | asq::F_AR)
#endif
#if ASQ_TU == 0 || ASQ_TU == 9
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_PLAIN)
#endif
#if ASQ_TU == 0 || ASQ_TU == 10
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_PHASE1)
#endif
#if ASQ_TU == 0 || ASQ_TU == 11
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_SWIGLU)
#endif
#if ASQ_TU == 0 || ASQ_TU == 12
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_ROPE)
#endif
#if ASQ_TU == 0 || ASQ_TU == 13
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_AR)
#endif
#if ASQ_TU == 0 || ASQ_TU == 14
ASQ_LAUNCH_INST(true, 2, 1, ASQ_LEAN_PHASE1)
#endif
#if ASQ_TU == 0 || ASQ_TU == 15
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_PLAIN_SK)
#endif
#if ASQ_TU == 0 || ASQ_TU == 16
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_PHASE1_SK)
#endif
#if ASQ_TU == 0 || ASQ_TU == 17
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_PLAIN_RESID)
#endif
#if ASQ_TU == 0 || ASQ_TU == 18
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_PHASE1_RESID)
#endif
#if ASQ_TU == 0 || ASQ_TU == 21
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_PHASE1_ROPE)
#endif
#if ASQ_TU == 0 || ASQ_TU == 22
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_PHASE1_SWIGLU)
#endif
#if ASQ_TU == 0 || ASQ_TU == 23
ASQ_LAUNCH_INST(false, 2, 1, ASQ_LEAN_NVLS)
ASQ_LAUNCH_INST(false, 1, 1, ASQ_LEAN_NVLS)
#endif
#if ASQ_TU == 0 || ASQ_TU == 24
ASQ_LAUNCH_INST(true, 2, 1, ASQ_LEAN_NVLS)
ASQ_LAUNCH_INST(true, 1, 1, ASQ_LEAN_NVLS)
#endif
#if ASQ_TU == 0 || ASQ_TU == 19
ASQ_LAUNCH_INST(false, 1, 1, ASQ_LEAN_PLAIN)
ASQ_LAUNCH_INST(false, 1, 1, ASQ_LEAN_PLAIN_SK)
#endif
#if ASQ_TU == 0 || ASQ_TU == 20
ASQ_LAUNCH_INST(false, 1, 1, ASQ_LEAN_PHASE1)
ASQ_LAUNCH_INST(false, 1, 1, ASQ_LEAN_PHASE1_SK)
#endif
#endif  // ASQ_TU
}  // namespace asq_launch

#if !defined(ASQ_TU) || ASQ_TU == 0
namespace {
using asq_launch::launch_cfg;
using asq_launch::max_multicast_clusters;

// CTA pairs (256-row tiles) whenever there is more than one 128-row panel.  ASQ_FORCE_CG=1|2 overrides
// the pairing (profiling / tests).
int pick_cta_group(int64_t M) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("ASQ_FORCE_CG");
    forced = (e != nullptr && (e[0] == '1' || e[0] == '2')) ? (e[0] - '0') : 0;
  }
  if (forced) return forced;
  return M > asq::BLOCK_M ? 2 : 1;
}

// Optional workspace of the GEMM-only entry points: only the stream-K region is used.
int attach_streamk(asq::LinearParams& p, void* workspace, size_t workspace_bytes) {
  if (workspace == nullptr) return ASQ_OK;  // no workspace: no stream-K split, everything else works
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(ASQ_ERR_INVALID, "workspace must be 1024-byte aligned");
  if (workspace_bytes < kFixedBytes) return fail(ASQ_ERR_WORKSPACE, "workspace needs %zu bytes, got %zu", kFixedBytes, workspace_bytes);
  Workspace ws;
  ws_layout(0, 16, workspace, &ws);
  p.sk_flags = ws.sk_flags;
  p.sk_partial = ws.sk_partial;
  return ASQ_OK;
}

// Shared launcher: `a8` is the 8-bit A matrix TMA reads (caller's matrix, or the workspace copy).
int launch_linear(bool fp8, const void* a8, const void* w, asq::LinearParams& p, cudaStream_t stream,
                  void* const* peer_y = nullptr) {
  int dev;
  DeviceState* st;
  int rc = get_device(&dev, &st);
  if (rc != ASQ_OK) return rc;
  if (!st->supported)
    return fail(ASQ_ERR_CUDA, "device %d is not compute capability 10.0 (sm_100a kernels only)", dev);
  {  // ASQ_DEBUG_TIMELINE=<device pointer, hex>: per-CTA phase timestamps (profiling builds of the caller)
    const char* e = getenv("ASQ_DEBUG_TIMELINE");
    p.dbg = (e != nullptr) ? reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 16)) : nullptr;
  }
  const int cg = pick_cta_group(p.M);
  const int tile_m = asq::BLOCK_M * cg;
  int mc_clusters = 0;
  // 4-CTA multicast clusters (two pairs sharing the activation rows): ASQ_MC=2 enables, =1 disables
  int mc = 1;
  {
    static int mc_env = -1;
    if (mc_env < 0) { const char* e = getenv("ASQ_MC"); mc_env = (e != nullptr && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0; }
    const int want = mc_env ? mc_env : 1;
    if (want == 2 && cg == 2 && p.N > asq::TILE_N && p.epi_kind != asq::EPI_SWIGLU && p.rope_cos == nullptr && p.ar_world <= 1 && p.group_of_blk == nullptr && p.batch_rows == 0) {
#if defined(ASQ_ENABLE_MC) || !defined(ASQ_TU)
      const int clusters = fp8 ? max_multicast_clusters<true>(dev) : max_multicast_clusters<false>(dev);
#else
      const int clusters = 0;  // the 4-CTA multicast experiment (measured neutral) is compiled only with -DASQ_ENABLE_MC
#endif
      const long long super_tiles = static_cast<long long>((p.M + tile_m - 1) / tile_m) * ((p.N + 2 * asq::TILE_N - 1) / (2 * asq::TILE_N));
      if (clusters > 0 && super_tiles >= clusters) mc = 2;
      if (mc == 2) mc_clusters = clusters;
    }
  }
  if (p.nvls_p_mc != nullptr) {
    // reducer CTAs (ASQ_NVLS_REDUCERS, default 32): enough 16-byte requests in flight to keep the NVLink path busy
    // (profiles/r02_allreduce.md: ~1 MB per GPU), taken from the SMs the GEMM would otherwise use
    static int red_env = -1;
    if (red_env < 0) { const char* e = getenv("ASQ_NVLS_REDUCERS"); red_env = (e != nullptr && atoi(e) > 0) ? atoi(e) : 0; }
    int r = red_env ? red_env : 32;
    if (r > st->sm_count / 2) r = st->sm_count / 2;
    p.nvls_reducers = (r + cg - 1) / cg * cg;
  }
  const int max_workers = (mc == 2) ? mc_clusters : (st->sm_count - p.nvls_reducers) / cg;
  p.num_m_blocks = (p.M + tile_m - 1) / tile_m;
  p.tile_m_blocks = cg;
  p.n_units = (p.N + asq::UNIT_N - 1) / asq::UNIT_N;
  p.num_k_blocks = (p.K + asq::BLOCK_K - 1) / asq::BLOCK_K;
  // Tile width in 64-column units.  Narrower tiles for decode-sized problems were measured slower on B200
  // (profiles/r01_notes.md: with 8-16 KB W boxes the 3-5 stage ring keeps too few bytes in flight), so the
  // width stays 256; ASQ_TILE_UNITS=1|2 narrows it for experiments.
  p.tile_units = asq::TILE_N / asq::UNIT_N;
  {
    static int tu_env = -1;
    if (tu_env < 0) { const char* e = getenv("ASQ_TILE_UNITS"); tu_env = (e != nullptr && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0; }
    if (mc == 2) {
      p.tile_units = 2 * asq::TILE_N / asq::UNIT_N;  // the walk hands out 512-wide super tiles, one half per pair
    } else if (tu_env && p.epi_kind != asq::EPI_SWIGLU && p.rope_cos == nullptr && p.ar_world <= 1 && p.group_of_blk == nullptr && p.batch_rows == 0) {
      p.tile_units = tu_env;
    } else {
      // Wave quantisation: pick 256- or 192-column tiles, whichever needs less (rounds x width); 192-wide tiles
      // move ~17% more operand bytes per MAC, charged as a 4% penalty.  E.g. M=2048, N=12288 on 74 CTA pairs:
      // 384 tiles -> 6 rounds x 256 vs 512 tiles -> 7 rounds x 192 (= 5.25 x 256).
      auto cost = [&](int units, double penalty) {
        const long long t = static_cast<long long>(p.num_m_blocks) * ((p.n_units + units - 1) / units);
        return static_cast<double>((t + max_workers - 1) / max_workers) * units * penalty;
      };
      // Measured: a round of 192-wide tiles takes as long as a round of 256-wide ones (the k-block time is set
      // by operand staging, not by MMA width), so the model below loses in practice; ASQ_TILE_192=1 enables it.
      static int use192 = -1;
      if (use192 < 0) { const char* e = getenv("ASQ_TILE_192"); use192 = (e != nullptr && e[0] == '1'); }
      if (use192 && cost(3, 1.04) < cost(4, 1.0)) p.tile_units = 3;
    }
  }
  p.num_n_blocks = (p.n_units + p.tile_units - 1) / p.tile_units;
  {
    // ASQ_RASTER=n<G> | m<G> overrides the walk (experiments)
    static int raster_env = -1, group_env = 0;
    if (raster_env < 0) {
      const char* e = getenv("ASQ_RASTER");
      raster_env = 0;
      if (e != nullptr && e[0] == 'm') { raster_env = 1; group_env = atoi(e + 1); }
      else if (e != nullptr && e[0] == 'n') { raster_env = 2; group_env = atoi(e + 1); }
    }
    // default: keep one group's W tiles (group * width * K bytes) within ~48 MB of the 126 MB L2
    const long long per_block = static_cast<long long>(p.tile_units) * asq::UNIT_N * p.K;
    long long gn = (48ll << 20) / per_block;
    p.raster_m = 0;
    p.group = static_cast<int>(gn < 1 ? 1 : (gn > p.num_n_blocks ? p.num_n_blocks : gn));
    if (raster_env == 1) { p.raster_m = 1; p.group = group_env > 0 ? group_env : p.num_m_blocks; }
    if (raster_env == 2 && group_env > 0) p.group = group_env;
  }
  // `rounds` full tiles per worker, then the left-over tiles: one each, or split along K (stream-K)
  const long long tiles = static_cast<long long>(p.num_m_blocks) * p.num_n_blocks;
  p.rounds = static_cast<int>(tiles / max_workers);
  p.tail_tiles = static_cast<int>(tiles - static_cast<long long>(p.rounds) * max_workers);
  int workers = p.rounds > 0 ? max_workers : p.tail_tiles;

  // Stream-K tail: instead of one more (mostly idle) round, the left-over tiles' k-iterations are dealt evenly
  // to the workers; also spreads problems with fewer tiles than SMs (decode-sized M) over the whole chip.
  p.sk_enabled = 0;
  {
    int sk_env_effective = 1;
    // Measured (profiles/r01_notes.md): with >= 1 full round the fix-up (contributor store -> owner load ->
    // epilogue, ~4 us on the critical path) costs more than the saved fraction of a round at K <= 11008, and
    // with many contributors per tile the owner's serial additions dominate; by default the split is used
    // only when there are fewer tiles than workers AND K is long (>= 64 k-blocks), at most 4-way.
    // ASQ_STREAMK=0 disables it, ASQ_STREAMK=2 forces it for every tail.
    static int sk_env = -1;
    if (sk_env < 0) { const char* e = getenv("ASQ_STREAMK"); sk_env = (e == nullptr) ? 1 : (e[0] - '0'); }
    if (sk_env == 1 && (p.rounds > 0 || p.num_k_blocks < 64)) sk_env_effective = 0;  // decode-sized M with a long K only
    constexpr int kMinIters = 4;  // k-blocks per stream-K segment, bounds the fix-up overhead
    const long long total_it = static_cast<long long>(p.tail_tiles) * p.num_k_blocks;
    if (mc == 1 && p.residual == nullptr && sk_env_effective && sk_env && p.tail_tiles > 0 && p.tail_tiles < max_workers && total_it / kMinIters >= 2 && total_it < (1ll << 30)) {
      if (p.sk_partial != nullptr && p.sk_flags != nullptr && max_workers * cg <= kSkMaxCtas) {  // caller gave a workspace
        long long sw = total_it / kMinIters;
        if (sk_env == 1 && sw > 4ll * p.tail_tiles) sw = 4ll * p.tail_tiles;  // the owner adds its contributors serially
        p.sk_workers = static_cast<int>(sw < max_workers ? sw : max_workers);
        if (p.sk_workers > p.tail_tiles) {  // otherwise whole tiles per worker are just as good
          p.sk_enabled = 1;
          p.sk_total = static_cast<int>(total_it);
          workers = p.rounds > 0 ? max_workers : p.sk_workers;
        }
      }
    }
  }

  CUtensorMap tmA, tmB, tmBu;
  rc = make_tmap(&tmA, a8, p.M, p.K, mc == 2 ? asq::BLOCK_M / 2 : asq::BLOCK_M);
  if (rc != ASQ_OK) return rc;
  const int64_t w_rows = static_cast<int64_t>(p.N) * (p.num_groups > 0 ? p.num_groups : 1);  // grouped: stacked expert weights
  rc = make_tmap(&tmB, w, w_rows, p.K, asq::TILE_N / cg);
  if (rc != ASQ_OK) return rc;
  rc = make_tmap(&tmBu, w, w_rows, p.K, asq::UNIT_N / cg);
  if (rc != ASQ_OK) return rc;
  // Output map: y viewed as bytes [M, N*elem]; 32-row x 128-byte boxes (one epilogue warp's staging tile).
  CUtensorMap tmY;
  memset(&tmY, 0, sizeof(tmY));
  {
    const int elem = (p.y_dtype == ASQ_BF16 || p.y_dtype == ASQ_F16) ? 2 : (p.y_dtype == ASQ_I8 ? 1 : 4);
    const bool swiglu = (p.epi_kind == asq::EPI_SWIGLU);  // output is [M, N/2]
    const long long row_bytes = static_cast<long long>(swiglu ? p.N / 2 : p.N) * elem;
    static int no_tma_store = -1;
    if (no_tma_store < 0) { const char* e = getenv("ASQ_NO_TMA_STORE"); no_tma_store = (e != nullptr && e[0] == '1'); }
    p.tma_store = (!no_tma_store && elem != 1 && row_bytes % 16 == 0) ? 1 : 0;
    if (swiglu) {
      p.tma_store = 1;  // the SwiGLU epilogue only exists as a staged store
      if (row_bytes % 16 != 0) return fail(ASQ_ERR_INVALID, "swiglu: output row pitch (%lld bytes) must be a multiple of 16", row_bytes);
      rc = (elem == 1) ? make_tmap_plain(&tmY, p.y, p.M, row_bytes, 64, 32) : make_tmap(&tmY, p.y, p.M, row_bytes, 32);
      if (rc != ASQ_OK) return rc;
    } else if (p.tma_store) {
      rc = make_tmap(&tmY, p.y, p.M, row_bytes, 32);
      if (rc != ASQ_OK) return rc;
    } else {
      tmY = tmA;  // unused, but must be a valid descriptor
    }
  }
  if (p.residual != nullptr) {
    const bool out16r = (p.y_dtype == ASQ_BF16 || p.y_dtype == ASQ_F16);
    if (!out16r || !p.tma_store || p.epi_kind != asq::EPI_DEQUANT || p.rope_cos != nullptr || p.ar_world > 1 || p.N % 8 != 0 ||
        (reinterpret_cast<uintptr_t>(p.residual) & 15))
      return fail(ASQ_ERR_INVALID, "residual add needs a 16-bit dequantised output with N %% 8 == 0 and a 16-byte aligned residual");
  }
  if (p.nvls_p_mc != nullptr) {
    // in-switch all-reduce: the ordinary 16-bit staged epilogue writes the partial tiles, the owners reduce them
    const bool out16n = (p.y_dtype == ASQ_BF16 || p.y_dtype == ASQ_F16);
    if (!p.tma_store || !out16n || p.epi_kind != asq::EPI_DEQUANT || p.x != nullptr || p.sk_enabled || mc != 1)
      return fail(ASQ_ERR_INVALID, "nvls all-reduce needs 8-bit activations and a 16-byte aligned 16-bit output row pitch");
    p.ar_tiles = p.num_m_blocks * p.num_n_blocks;
    const asq::ExtraMapsT<0> none_nvls{};
    if (cg == 2) return fp8 ? launch_cfg<true, 2, 1, ASQ_LEAN_NVLS>(tmA, tmB, tmBu, tmY, none_nvls, p, workers, stream)
                            : launch_cfg<false, 2, 1, ASQ_LEAN_NVLS>(tmA, tmB, tmBu, tmY, none_nvls, p, workers, stream);
    return fp8 ? launch_cfg<true, 1, 1, ASQ_LEAN_NVLS>(tmA, tmB, tmBu, tmY, none_nvls, p, workers, stream)
               : launch_cfg<false, 1, 1, ASQ_LEAN_NVLS>(tmA, tmB, tmBu, tmY, none_nvls, p, workers, stream);
  }
  if (p.ar_world > 1) {
    if (!p.tma_store || peer_y == nullptr) return fail(ASQ_ERR_INVALID, "all-reduce mode needs a 16-byte aligned 16-bit output row pitch");
    if (fp8 || mc != 1) return fail(ASQ_ERR_UNSUPPORTED, "all-reduce mode exists for the int8 path only");
    asq::ExtraMapsT<1> tmPeers;
    memset(&tmPeers, 0, sizeof(tmPeers));
    const long long row_bytes = static_cast<long long>(p.N) * 2;
    for (int i = 0; i < p.ar_world - 1; ++i) {
      rc = make_tmap(&tmPeers.m[i], peer_y[i], p.M, row_bytes, 32);
      if (rc != ASQ_OK) return rc;
    }
    p.ar_tiles = p.num_m_blocks * p.num_n_blocks;
    p.ar_cnt_max = (p.ar_tiles + p.ar_world - 1) / p.ar_world;
    const char* lean = getenv("ASQ_LEAN");
    if (cg == 2 && p.x == nullptr && !p.sk_enabled && p.epi_kind == asq::EPI_DEQUANT && p.out_fq_scale == 0.f && p.dbg == nullptr &&
        !(lean != nullptr && lean[0] == '0'))
      return launch_cfg<false, 2, 1, ASQ_LEAN_AR>(tmA, tmB, tmBu, tmY, tmPeers, p, workers, stream);
    return cg == 2 ? launch_cfg<false, 2, 1, asq::F_FULL | asq::F_AR>(tmA, tmB, tmBu, tmY, tmPeers, p, workers, stream)
                   : launch_cfg<false, 1, 1, asq::F_FULL | asq::F_AR>(tmA, tmB, tmBu, tmY, tmPeers, p, workers, stream);
  }
  const asq::ExtraMapsT<0> none{};
  asq::ExtraMapsT<2> resid;  // instantiations with F_RESID (the full kernel included) carry one map over the residual
  memset(&resid, 0, sizeof(resid));
  if (p.residual != nullptr) {
    rc = make_tmap(&resid.m[0], p.residual, p.M, static_cast<long long>(p.N) * 2, 32);
    if (rc != ASQ_OK) return rc;
  } else {
    resid.m[0] = tmA;  // unused, but must be a valid descriptor
  }
  constexpr int FULL = asq::F_FULL;
#if defined(ASQ_ENABLE_MC) || !defined(ASQ_TU)
  if (mc == 2) return fp8 ? launch_cfg<true, 2, 2, FULL>(tmA, tmB, tmBu, tmY, resid, p, workers, stream)
                          : launch_cfg<false, 2, 2, FULL>(tmA, tmB, tmBu, tmY, resid, p, workers, stream);
#endif
  if (fp8) {
    const char* lean = getenv("ASQ_LEAN");
    const bool out16 = (p.y_dtype == ASQ_BF16 || p.y_dtype == ASQ_F16);
    // the FP8 per-token / static module forward (BASELINE config 5): fused prologue + 16-bit dequant epilogue
    if (cg == 2 && p.x != nullptr && !p.sk_enabled && p.group_of_blk == nullptr && p.batch_rows == 0 && p.rope_cos == nullptr &&
        p.epi_kind == asq::EPI_DEQUANT && p.tma_store && out16 && p.out_fq_scale == 0.f && p.dbg == nullptr &&
        p.act_mode != ASQ_ACT_PER_TENSOR_DYNAMIC && !(lean != nullptr && lean[0] == '0'))
      return launch_cfg<true, 2, 1, ASQ_LEAN_PHASE1>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
    return cg == 2 ? launch_cfg<true, 2, 1, FULL>(tmA, tmB, tmBu, tmY, resid, p, workers, stream)
                   : launch_cfg<true, 1, 1, FULL>(tmA, tmB, tmBu, tmY, resid, p, workers, stream);
  }
  // What this launch needs; a lean instantiation is used when it covers exactly that (ASQ_LEAN=0 disables).
  static int lean_env = -1;
  if (lean_env < 0) { const char* e = getenv("ASQ_LEAN"); lean_env = (e != nullptr && e[0] == '0') ? 0 : 1; }
  const bool out16 = (p.y_dtype == ASQ_BF16 || p.y_dtype == ASQ_F16);
  int need = 0;
  if (p.x != nullptr) need |= asq::F_PHASE1;
  if (p.sk_enabled) need |= asq::F_SK;
  if (p.group_of_blk != nullptr || p.batch_rows > 0) need |= asq::F_GROUP;
  if (p.epi_kind == asq::EPI_SWIGLU) need |= asq::F_SWIGLU;
  else if (p.rope_cos != nullptr) need |= asq::F_ROPE;
  else if (p.epi_kind == asq::EPI_DEQUANT && p.tma_store && out16 && p.out_fq_scale == 0.f) need |= asq::F_DEQ16;
  else need |= asq::F_OUT_ANY;
  if (p.residual != nullptr) need |= asq::F_RESID;
  if (p.dbg != nullptr) need |= asq::F_OUT_ANY;  // timeline runs: always the full kernel
  if (cg == 2) {
    if (lean_env) {
      if (need == (ASQ_LEAN_PLAIN)) return launch_cfg<false, 2, 1, ASQ_LEAN_PLAIN>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
      if (need == (ASQ_LEAN_PHASE1)) return launch_cfg<false, 2, 1, ASQ_LEAN_PHASE1>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
      if (need == (ASQ_LEAN_SWIGLU)) return launch_cfg<false, 2, 1, ASQ_LEAN_SWIGLU>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
      if (need == (ASQ_LEAN_ROPE)) return launch_cfg<false, 2, 1, ASQ_LEAN_ROPE>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
      if (need == (ASQ_LEAN_PHASE1_ROPE)) return launch_cfg<false, 2, 1, ASQ_LEAN_PHASE1_ROPE>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
      if (need == (ASQ_LEAN_PHASE1_SWIGLU)) return launch_cfg<false, 2, 1, ASQ_LEAN_PHASE1_SWIGLU>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
      if (need == (ASQ_LEAN_PLAIN_RESID)) return launch_cfg<false, 2, 1, ASQ_LEAN_PLAIN_RESID>(tmA, tmB, tmBu, tmY, resid, p, workers, stream);
      if (need == (ASQ_LEAN_PHASE1_RESID)) return launch_cfg<false, 2, 1, ASQ_LEAN_PHASE1_RESID>(tmA, tmB, tmBu, tmY, resid, p, workers, stream);
      if (need == (ASQ_LEAN_PLAIN_SK)) return launch_cfg<false, 2, 1, ASQ_LEAN_PLAIN_SK>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
      if (need == (ASQ_LEAN_PHASE1_SK)) return launch_cfg<false, 2, 1, ASQ_LEAN_PHASE1_SK>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
    }
    return launch_cfg<false, 2, 1, FULL>(tmA, tmB, tmBu, tmY, resid, p, workers, stream);
  }
  // decode-sized launches (M <= 128, one CTA per tile): the same lean sets for the module forward and the int8-in entry
  if (lean_env) {
    if (need == (ASQ_LEAN_PLAIN)) return launch_cfg<false, 1, 1, ASQ_LEAN_PLAIN>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
    if (need == (ASQ_LEAN_PHASE1)) return launch_cfg<false, 1, 1, ASQ_LEAN_PHASE1>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
    if (need == (ASQ_LEAN_PLAIN_SK)) return launch_cfg<false, 1, 1, ASQ_LEAN_PLAIN_SK>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
    if (need == (ASQ_LEAN_PHASE1_SK)) return launch_cfg<false, 1, 1, ASQ_LEAN_PHASE1_SK>(tmA, tmB, tmBu, tmY, none, p, workers, stream);
  }
  return launch_cfg<false, 1, 1, FULL>(tmA, tmB, tmBu, tmY, resid, p, workers, stream);
}

int check_common(const void* a, const void* w, const void* y, int64_t M, int64_t N, int64_t K) {
  if (M < 0 || N <= 0 || K <= 0) return fail(ASQ_ERR_INVALID, "bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  if (K % 16 != 0) return fail(ASQ_ERR_INVALID, "K=%lld must be a multiple of 16 (TMA row pitch)", (long long)K);
  if (M > 0x7fffff00LL || N > 0x7fffff00LL || K > 0x7fffff00LL) return fail(ASQ_ERR_INVALID, "dimension too large");
  if (M == 0) return ASQ_OK;
  if (a == nullptr || w == nullptr || y == nullptr) return fail(ASQ_ERR_INVALID, "null pointer argument");
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) ||
      (reinterpret_cast<uintptr_t>(y) & 15))
    return fail(ASQ_ERR_INVALID, "x, w and y must be 16-byte aligned");
  return ASQ_OK;
}

bool is_float_dtype(int d) { return d == ASQ_F32 || d == ASQ_F16 || d == ASQ_BF16; }

// Decode-sized launches (M <= 16) take the weight-streaming kernel of asq_smallm.cu (ASQ_SMALLM=0 disables).
bool smallm_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("ASQ_SMALLM"); on = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

int fused_linear(bool fp8, const void* x, int x_dtype, const void* w, const float* bias, void* y, int y_dtype,
                 int64_t M, int64_t N, int64_t K, int act_mode, float quant_scale, float dequant_scale,
                 const float* col_scale, float* row_scale_out, int div_mode, void* workspace,
                 size_t workspace_bytes, void* stream, float out_fq_scale = 0.f, const void* residual = nullptr) {
  int rc = check_common(x, w, y, M, N, K);
  if (rc != ASQ_OK) return rc;
  if (!is_float_dtype(x_dtype) || !is_float_dtype(y_dtype)) return fail(ASQ_ERR_INVALID, "x/y dtype must be f32, f16 or bf16");
  if (div_mode != ASQ_DIV_RECIPROCAL && div_mode != ASQ_DIV_EXACT) return fail(ASQ_ERR_INVALID, "bad div_mode %d", div_mode);
  if (act_mode == ASQ_ACT_PER_TENSOR_DYNAMIC && !fp8)
    return fail(ASQ_ERR_UNSUPPORTED, "per-tensor dynamic activation scale exists for the fp8 path only (quantization.py:144-170)");
  if (act_mode != ASQ_ACT_ROUND && act_mode != ASQ_ACT_SCALE && act_mode != ASQ_ACT_PER_TOKEN &&
      act_mode != ASQ_ACT_ROW_SCALE_GIVEN && act_mode != ASQ_ACT_PER_TENSOR_DYNAMIC)
    return fail(ASQ_ERR_INVALID, "bad act_mode %d", act_mode);
  if (act_mode == ASQ_ACT_ROW_SCALE_GIVEN && row_scale_out == nullptr && M > 0)
    return fail(ASQ_ERR_INVALID, "ASQ_ACT_ROW_SCALE_GIVEN needs the [M] row scales in row_scale");
  if (fp8 && act_mode == ASQ_ACT_ROUND) return fail(ASQ_ERR_INVALID, "ASQ_ACT_ROUND is int8-only");
  if (M == 0) return ASQ_OK;
  if ((M + asq::BLOCK_M - 1) / asq::BLOCK_M + 1 > static_cast<int64_t>(kSyncBytes / 4) - 2)
    return fail(ASQ_ERR_UNSUPPORTED, "M=%lld exceeds the %zu row panels one launch tracks; split the batch", (long long)M, kSyncBytes / 4 - 1);
  if (workspace == nullptr || workspace_bytes < asq_workspace_bytes(M, K))
    return fail(ASQ_ERR_WORKSPACE, "workspace needs %zu bytes, got %zu", asq_workspace_bytes(M, K), workspace_bytes);
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(ASQ_ERR_INVALID, "workspace must be 1024-byte aligned");

  Workspace ws;
  ws_layout(M, K, workspace, &ws);
  if (!fp8 && residual == nullptr && out_fq_scale == 0.f && act_mode != ASQ_ACT_PER_TENSOR_DYNAMIC && smallm_enabled() &&
      asq_smallm_supported(M, N, K, y_dtype)) {
    // decode-sized M: stand-alone prologue into the workspace, then the weight-streaming kernel (two small launches)
    const bool per_token = (act_mode == ASQ_ACT_PER_TOKEN || act_mode == ASQ_ACT_ROW_SCALE_GIVEN);
    float* rs = per_token ? ws.row_scale : nullptr;
    if (act_mode == ASQ_ACT_ROW_SCALE_GIVEN) rs = row_scale_out;
    rc = asq_quantize_act(x, x_dtype, ws.a_q, rs, M, K, act_mode, quant_scale, div_mode, 0, stream);
    if (rc != ASQ_OK) return rc;
    if (act_mode == ASQ_ACT_PER_TOKEN && row_scale_out != nullptr) {
      cudaError_t e = cudaMemcpyAsync(row_scale_out, ws.row_scale, static_cast<size_t>(M) * 4, cudaMemcpyDeviceToDevice,
                                      static_cast<cudaStream_t>(stream));
      if (e != cudaSuccess) return fail(ASQ_ERR_CUDA, "row scale copy failed: %s", cudaGetErrorString(e));
    }
    SmallMParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.xq = reinterpret_cast<const int8_t*>(ws.a_q); sp.w = static_cast<const int8_t*>(w); sp.row_scale = rs;
    sp.col_scale = col_scale; sp.bias = bias; sp.y = y; sp.dequant_scale = dequant_scale;
    sp.M = static_cast<int>(M); sp.N = static_cast<int>(N); sp.K = static_cast<int>(K); sp.y_dtype = y_dtype;
    return asq_smallm_launch(sp, static_cast<cudaStream_t>(stream));
  }
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.a_q = ws.a_q; p.row_scale = ws.row_scale; p.sync = ws.sync;
  p.sk_flags = ws.sk_flags; p.sk_partial = ws.sk_partial;
  if (act_mode == ASQ_ACT_ROW_SCALE_GIVEN) p.row_scale_in = row_scale_out;  // in: caller-supplied scales
  else p.row_scale_out = row_scale_out;
  p.quant_scale = quant_scale; p.inv_quant_scale = 1.0f / quant_scale;
  p.qmax = fp8 ? 448.0f : 127.0f; p.inv_qmax = 1.0f / p.qmax;
  p.y = y; p.bias = bias; p.col_scale = col_scale; p.dequant_scale = dequant_scale;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.x_dtype = x_dtype; p.y_dtype = y_dtype; p.act_mode = act_mode; p.div_mode = div_mode;
  p.epi_kind = asq::EPI_DEQUANT;
  p.residual = residual;
  p.out_fq_scale = out_fq_scale;
  p.inv_out_fq_scale = out_fq_scale != 0.f ? 1.0f / out_fq_scale : 0.f;
  return launch_linear(fp8, ws.a_q, w, p, static_cast<cudaStream_t>(stream));
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

int asq_version(void) { return ASQ_VERSION; }
const char* asq_last_error(void) { return g_err; }

int asq_device_supported(void) {
  int dev;
  DeviceState* st;
  if (get_device(&dev, &st) != ASQ_OK) return 0;
  return st->supported ? 1 : 0;
}

size_t asq_workspace_bytes(int64_t M, int64_t K) {
  if (M <= 0 || K <= 0) return kFixedBytes;  // GEMM-only entry points: counters + stream-K region
  return ws_layout(M, K, nullptr, nullptr);
}

int asq_w8a8_linear(const void* x, int x_dtype, const int8_t* w, const float* bias, void* y, int y_dtype,
                    int64_t M, int64_t N, int64_t K, int act_mode, float quant_scale, float dequant_scale,
                    const float* col_scale, float* row_scale_out, int div_mode, void* workspace,
                    size_t workspace_bytes, void* stream) {
  return fused_linear(false, x, x_dtype, w, bias, y, y_dtype, M, N, K, act_mode, quant_scale, dequant_scale,
                      col_scale, row_scale_out, div_mode, workspace, workspace_bytes, stream);
}

int asq_fp8_linear(const void* x, int x_dtype, const uint8_t* w_e4m3, const float* bias, void* y, int y_dtype,
                   int64_t M, int64_t N, int64_t K, int act_mode, float in_scale, float w_scale, float out_scale,
                   float* row_scale_out, int div_mode, void* workspace, size_t workspace_bytes, void* stream) {
  // dynamic scales live in the row-scale vector: y = acc * (w_scale * s[m]); static: y = acc * (w_scale * in_scale)
  const float ds = (act_mode == ASQ_ACT_SCALE) ? w_scale * in_scale : w_scale;
  return fused_linear(true, x, x_dtype, w_e4m3, bias, y, y_dtype, M, N, K, act_mode, in_scale, ds, nullptr,
                      row_scale_out, div_mode, workspace, workspace_bytes, stream, out_scale);
}

int asq_fp8_linear_cs(const void* x, int x_dtype, const uint8_t* w_e4m3, const float* bias, void* y, int y_dtype,
                      int64_t M, int64_t N, int64_t K, int act_mode, const float* w_col_scale, float* row_scale_out,
                      int div_mode, void* workspace, size_t workspace_bytes, void* stream) {
  if (w_col_scale == nullptr) return fail(ASQ_ERR_INVALID, "asq_fp8_linear_cs needs the [N] column scales");
  if (act_mode != ASQ_ACT_PER_TOKEN && act_mode != ASQ_ACT_PER_TENSOR_DYNAMIC && act_mode != ASQ_ACT_ROW_SCALE_GIVEN)
    return fail(ASQ_ERR_UNSUPPORTED, "asq_fp8_linear_cs: dynamic activation scales only (act_mode %d)", act_mode);
  // the epilogue multiplies by col_scale[n] * s[m]: the scalar dequant scale is unused when a column vector is given
  return fused_linear(true, x, x_dtype, w_e4m3, bias, y, y_dtype, M, N, K, act_mode, 1.0f, 1.0f, w_col_scale,
                      row_scale_out, div_mode, workspace, workspace_bytes, stream, 0.f);
}

// ---- RMSNorm as the prologue of the q|k|v and gate|up launches (the norm kernel disappears from the layer)
namespace {
int prep_rmsnorm_prologue(asq::LinearParams& p, const void* x, int x_dtype, const void* norm_weight, float eps, int64_t M,
                          int64_t K, void* workspace, size_t workspace_bytes) {
  if (x_dtype != ASQ_BF16 && x_dtype != ASQ_F16) return fail(ASQ_ERR_INVALID, "rmsnorm prologue: x dtype must be f16 or bf16");
  if (norm_weight == nullptr || (reinterpret_cast<uintptr_t>(norm_weight) & 15) || K % 8 != 0)
    return fail(ASQ_ERR_INVALID, "rmsnorm prologue: norm weight must be a 16-byte aligned [K] vector, K %% 8 == 0");
  if ((M + asq::BLOCK_M - 1) / asq::BLOCK_M + 1 > static_cast<int64_t>(kSyncBytes / 4) - 2)
    return fail(ASQ_ERR_UNSUPPORTED, "M=%lld exceeds the row panels one launch tracks; split the batch", (long long)M);
  if (workspace == nullptr || workspace_bytes < asq_workspace_bytes(M, K))
    return fail(ASQ_ERR_WORKSPACE, "workspace needs %zu bytes, got %zu", asq_workspace_bytes(M, K), workspace_bytes);
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(ASQ_ERR_INVALID, "workspace must be 1024-byte aligned");
  Workspace ws;
  ws_layout(M, K, workspace, &ws);
  p.x = x; p.x_dtype = x_dtype; p.a_q = ws.a_q; p.row_scale = ws.row_scale; p.sync = ws.sync;
  p.act_mode = ASQ_ACT_RMSNORM; p.norm_weight = norm_weight; p.norm_eps = eps;
  p.quant_scale = 1.f; p.inv_quant_scale = 1.f; p.qmax = 127.0f; p.inv_qmax = 1.0f / 127.0f;
  return ASQ_OK;
}
}  // namespace

int asq_w8a8_rmsnorm_linear_rope(const void* x, int x_dtype, const void* norm_weight, float eps, const int8_t* w,
                                 const float* bias, void* y, int64_t M, int64_t N, int64_t K, float dequant_scale,
                                 const float* col_scale, const void* cos_table, const void* sin_table, int64_t S,
                                 int64_t rope_cols, int64_t head_dim, int halves_equal, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  int rc = check_common(x, w, y, M, N, K);
  if (rc != ASQ_OK || M == 0) return rc;
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  rc = prep_rmsnorm_prologue(p, x, x_dtype, norm_weight, eps, M, K, workspace, workspace_bytes);
  if (rc != ASQ_OK) return rc;
  if (cos_table != nullptr) {
    if (head_dim != 128) return fail(ASQ_ERR_UNSUPPORTED, "rope epilogue: head_dim %lld (only 128)", (long long)head_dim);
    if (N % 128 != 0 || rope_cols % 128 != 0 || rope_cols < 0 || rope_cols > N || S <= 0 || S > 0x7fffffffLL || sin_table == nullptr ||
        (reinterpret_cast<uintptr_t>(cos_table) & 15) || (reinterpret_cast<uintptr_t>(sin_table) & 15))
      return fail(ASQ_ERR_INVALID, "rope epilogue: bad table / column arguments");
    p.rope_cos = cos_table; p.rope_sin = sin_table; p.rope_S = static_cast<int>(S); p.rope_cols = static_cast<int>(rope_cols);
    p.rope_halves_equal = halves_equal ? 1 : 0;
  }
  p.y = y; p.bias = bias; p.col_scale = col_scale; p.dequant_scale = dequant_scale;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.y_dtype = x_dtype; p.epi_kind = asq::EPI_DEQUANT;
  return launch_linear(false, p.a_q, w, p, static_cast<cudaStream_t>(stream));
}

int asq_w8a8_rmsnorm_gateup_swiglu(const void* x, int x_dtype, const void* norm_weight, float eps, const int8_t* w_il,
                                   const float* bias_il, void* out, int out_dtype, int64_t M, int64_t N, int64_t K,
                                   float gate_dequant_scale, float up_dequant_scale, const float* col_scale_il,
                                   float out_quant_scale, int div_mode, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  int rc = check_common(x, w_il, out, M, N, K);
  if (rc != ASQ_OK) return rc;
  if (N % 64 != 0) return fail(ASQ_ERR_INVALID, "swiglu: N=%lld (2 x intermediate) must be a multiple of 64", (long long)N);
  if (out_dtype != ASQ_I8 && out_dtype != x_dtype) return fail(ASQ_ERR_INVALID, "swiglu: out_dtype must be i8 or equal x's dtype");
  if (div_mode != ASQ_DIV_RECIPROCAL && div_mode != ASQ_DIV_EXACT) return fail(ASQ_ERR_INVALID, "bad div_mode %d", div_mode);
  if (out_dtype == ASQ_I8 && !(out_quant_scale > 0.f)) return fail(ASQ_ERR_INVALID, "swiglu: out_quant_scale must be positive");
  if (M == 0) return ASQ_OK;
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  rc = prep_rmsnorm_prologue(p, x, x_dtype, norm_weight, eps, M, K, workspace, workspace_bytes);
  if (rc != ASQ_OK) return rc;
  p.y = out; p.bias = bias_il; p.col_scale = col_scale_il;
  p.dequant_scale = gate_dequant_scale; p.dequant_scale_up = up_dequant_scale;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.y_dtype = out_dtype; p.mid_dtype = x_dtype; p.epi_kind = asq::EPI_SWIGLU; p.div_mode = div_mode;
  p.out_quant_scale = out_quant_scale; p.inv_out_quant_scale = out_dtype == ASQ_I8 ? 1.0f / out_quant_scale : 0.f;
  return launch_linear(false, p.a_q, w_il, p, static_cast<cudaStream_t>(stream));
}

int asq_w8a8_linear_q8_rope(const int8_t* xq, const float* row_scale, const int8_t* w, const float* bias, void* y,
                            int y_dtype, int64_t M, int64_t N, int64_t K, float dequant_scale, const float* col_scale,
                            const void* cos_table, const void* sin_table, int64_t S, int64_t rope_cols, int64_t head_dim,
                            int halves_equal, void* stream) {
  int rc = check_common(xq, w, y, M, N, K);
  if (rc != ASQ_OK) return rc;
  if (y_dtype != ASQ_BF16 && y_dtype != ASQ_F16) return fail(ASQ_ERR_INVALID, "rope epilogue: y dtype must be f16 or bf16");
  if (head_dim != 128) return fail(ASQ_ERR_UNSUPPORTED, "rope epilogue: head_dim %lld (only 128; use asq_rope_inplace)", (long long)head_dim);
  if (N % 128 != 0 || rope_cols % 128 != 0 || rope_cols < 0 || rope_cols > N)
    return fail(ASQ_ERR_INVALID, "rope epilogue: N=%lld and rope_cols=%lld must be multiples of 128, rope_cols <= N", (long long)N, (long long)rope_cols);
  if (S <= 0 || S > 0x7fffffffLL) return fail(ASQ_ERR_INVALID, "rope epilogue: bad table length S=%lld", (long long)S);
  if (M == 0) return ASQ_OK;
  if (cos_table == nullptr || sin_table == nullptr || (reinterpret_cast<uintptr_t>(cos_table) & 15) || (reinterpret_cast<uintptr_t>(sin_table) & 15))
    return fail(ASQ_ERR_INVALID, "rope epilogue: cos / sin tables must be non-null and 16-byte aligned");
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  p.y = y; p.bias = bias; p.col_scale = col_scale; p.dequant_scale = dequant_scale;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.y_dtype = y_dtype; p.epi_kind = asq::EPI_DEQUANT;
  p.act_mode = row_scale != nullptr ? ASQ_ACT_ROW_SCALE_GIVEN : ASQ_ACT_ROUND;
  p.row_scale = const_cast<float*>(row_scale);
  p.rope_cos = cos_table; p.rope_sin = sin_table; p.rope_S = static_cast<int>(S); p.rope_cols = static_cast<int>(rope_cols);
  p.rope_halves_equal = halves_equal ? 1 : 0;
  return launch_linear(false, xq, w, p, static_cast<cudaStream_t>(stream));
}

int asq_w8a8_linear_res(const void* x, int x_dtype, const int8_t* w, const float* bias, const void* residual, void* y,
                        int y_dtype, int64_t M, int64_t N, int64_t K, int act_mode, float quant_scale, float dequant_scale,
                        const float* col_scale, float* row_scale_out, int div_mode, void* workspace,
                        size_t workspace_bytes, void* stream) {
  return fused_linear(false, x, x_dtype, w, bias, y, y_dtype, M, N, K, act_mode, quant_scale, dequant_scale,
                      col_scale, row_scale_out, div_mode, workspace, workspace_bytes, stream, 0.f, residual);
}

static int linear_q8_impl(const int8_t* xq, const float* row_scale, const int8_t* w, const float* bias, const void* residual,
                          void* y, int y_dtype, int64_t M, int64_t N, int64_t K, float dequant_scale, const float* col_scale,
                          void* workspace, size_t workspace_bytes, void* stream);

int asq_w8a8_linear_q8_res(const int8_t* xq, const float* row_scale, const int8_t* w, const float* bias, const void* residual,
                           void* y, int y_dtype, int64_t M, int64_t N, int64_t K, float dequant_scale, const float* col_scale,
                           void* stream) {
  return linear_q8_impl(xq, row_scale, w, bias, residual, y, y_dtype, M, N, K, dequant_scale, col_scale, nullptr, 0, stream);
}

int asq_w8a8_linear_q8(const int8_t* xq, const float* row_scale, const int8_t* w, const float* bias, void* y,
                       int y_dtype, int64_t M, int64_t N, int64_t K, float dequant_scale, const float* col_scale,
                       void* workspace, size_t workspace_bytes, void* stream) {
  return linear_q8_impl(xq, row_scale, w, bias, nullptr, y, y_dtype, M, N, K, dequant_scale, col_scale, workspace, workspace_bytes, stream);
}

static int linear_q8_impl(const int8_t* xq, const float* row_scale, const int8_t* w, const float* bias, const void* residual,
                          void* y, int y_dtype, int64_t M, int64_t N, int64_t K, float dequant_scale, const float* col_scale,
                          void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(xq, w, y, M, N, K);
  if (rc != ASQ_OK || M == 0) return rc;
  if (!is_float_dtype(y_dtype)) return fail(ASQ_ERR_INVALID, "y dtype must be f32, f16 or bf16");
  if (residual == nullptr && smallm_enabled() && asq_smallm_supported(M, N, K, y_dtype)) {
    SmallMParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.xq = xq; sp.w = w; sp.row_scale = row_scale; sp.col_scale = col_scale; sp.bias = bias; sp.y = y;
    sp.dequant_scale = dequant_scale;
    sp.M = static_cast<int>(M); sp.N = static_cast<int>(N); sp.K = static_cast<int>(K); sp.y_dtype = y_dtype;
    return asq_smallm_launch(sp, static_cast<cudaStream_t>(stream));
  }
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  rc = attach_streamk(p, workspace, workspace_bytes);
  if (rc != ASQ_OK) return rc;
  p.y = y; p.bias = bias; p.col_scale = col_scale; p.dequant_scale = dequant_scale;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.y_dtype = y_dtype; p.epi_kind = asq::EPI_DEQUANT; p.residual = residual;
  // phase 1 is skipped (p.x == nullptr); the epilogue reads caller-supplied per-token scales if given
  p.act_mode = row_scale != nullptr ? ASQ_ACT_ROW_SCALE_GIVEN : ASQ_ACT_ROUND;
  p.row_scale = const_cast<float*>(row_scale);
  return launch_linear(false, xq, w, p, static_cast<cudaStream_t>(stream));
}

int asq_w8a8_gateup_swiglu_q8(const int8_t* xq, const float* row_scale, const int8_t* w_il, const float* bias_il,
                              void* out, int out_dtype, int mid_dtype, int64_t M, int64_t N, int64_t K,
                              float gate_dequant_scale, float up_dequant_scale, const float* col_scale_il,
                              float out_quant_scale, int div_mode, void* stream) {
  int rc = check_common(xq, w_il, out, M, N, K);
  if (rc != ASQ_OK) return rc;
  if (N % 64 != 0) return fail(ASQ_ERR_INVALID, "swiglu: N=%lld (2 x intermediate) must be a multiple of 64", (long long)N);
  if (mid_dtype != ASQ_BF16 && mid_dtype != ASQ_F16) return fail(ASQ_ERR_INVALID, "swiglu: mid_dtype must be f16 or bf16");
  if (out_dtype != ASQ_I8 && out_dtype != mid_dtype) return fail(ASQ_ERR_INVALID, "swiglu: out_dtype must be i8 or equal mid_dtype");
  if (div_mode != ASQ_DIV_RECIPROCAL && div_mode != ASQ_DIV_EXACT) return fail(ASQ_ERR_INVALID, "bad div_mode %d", div_mode);
  if (out_dtype == ASQ_I8 && !(out_quant_scale > 0.f)) return fail(ASQ_ERR_INVALID, "swiglu: out_quant_scale must be positive");
  if (M == 0) return ASQ_OK;
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  p.y = out; p.bias = bias_il; p.col_scale = col_scale_il;
  p.dequant_scale = gate_dequant_scale; p.dequant_scale_up = up_dequant_scale;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.y_dtype = out_dtype; p.mid_dtype = mid_dtype; p.epi_kind = asq::EPI_SWIGLU; p.div_mode = div_mode;
  p.out_quant_scale = out_quant_scale; p.inv_out_quant_scale = out_dtype == ASQ_I8 ? 1.0f / out_quant_scale : 0.f;
  p.act_mode = row_scale != nullptr ? ASQ_ACT_ROW_SCALE_GIVEN : ASQ_ACT_ROUND;
  p.row_scale = const_cast<float*>(row_scale);
  return launch_linear(false, xq, w_il, p, static_cast<cudaStream_t>(stream));
}

// ---- batched INT8 GEMM (csrc/kernels/bmm.cu): c[b] = epilogue(alpha * a[b] . w[b]^T)
int asq_i8bmm(const int8_t* a, const int8_t* w, void* c, int c_dtype, int64_t batch, int64_t M, int64_t N, int64_t K,
              float alpha, void* stream) {
  if (batch < 0) return fail(ASQ_ERR_INVALID, "bmm: negative batch");
  if (c_dtype != ASQ_I8 && c_dtype != ASQ_I32 && c_dtype != ASQ_F32) return fail(ASQ_ERR_INVALID, "bmm: output dtype must be i8, i32 or f32");
  int rc = check_common(a, w, c, M, N, K);
  if (rc != ASQ_OK || M == 0 || batch == 0) return rc;
  if (batch * M > 0x7fffff00LL || batch * N > 0x7fffff00LL) return fail(ASQ_ERR_INVALID, "bmm: batch * rows too large");
  const size_t elem = c_dtype == ASQ_I8 ? 1 : 4;
  auto fill = [&](asq::LinearParams& p, void* y, int64_t rows) {
    memset(&p, 0, sizeof(p));
    p.y = y; p.M = static_cast<int>(rows); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
    p.y_dtype = c_dtype; p.act_mode = ASQ_ACT_ROUND;
    if (c_dtype == ASQ_I32) { p.epi_kind = asq::EPI_RAW_I32; }           // bmm_s8t_s8n_s32t (alpha = 1)
    else { p.epi_kind = asq::EPI_ALPHA_BETA; p.alpha = alpha; p.beta = 0.f; }  // _f32t: alpha*acc; _s8t: sat(rint(alpha*acc))
  };
  const int tile_m = asq::BLOCK_M * pick_cta_group(batch * M);
  if (batch > 1 && M % tile_m == 0 && batch <= 65535) {
    // every batch is a whole number of M tiles: ONE launch over the stacked [batch*M, K] x [batch*N, K] operands,
    // the tile's batch index selects the weight rows (the grouped-GEMM path with group = row / M)
    asq::LinearParams p;
    fill(p, c, batch * M);
    p.batch_rows = static_cast<int>(M);
    p.num_groups = static_cast<int>(batch);
    return launch_linear(false, a, w, p, static_cast<cudaStream_t>(stream));
  }
  for (int64_t b = 0; b < batch; ++b) {  // ragged M: one launch per batch entry
    asq::LinearParams p;
    fill(p, static_cast<uint8_t*>(c) + static_cast<size_t>(b) * M * N * elem, M);
    if ((reinterpret_cast<uintptr_t>(p.y) & 15) || ((b * M * K) & 15) || ((b * N * K) & 15))
      return fail(ASQ_ERR_INVALID, "bmm: per-batch matrices must stay 16-byte aligned (M*K, N*K, M*N*elem multiples of 16)");
    rc = launch_linear(false, a + b * M * K, w + b * N * K, p, static_cast<cudaStream_t>(stream));
    if (rc != ASQ_OK) return rc;
  }
  return ASQ_OK;
}

// ---- grouped (MoE expert) linear: one launch for all experts
int asq_w8a8_grouped_linear(const void* x, int x_dtype, const int8_t* w_stacked, void* y, int y_dtype,
                            int64_t M_pad, int64_t N, int64_t K, int num_groups, const int32_t* group_of_blk,
                            const float* group_dequant_scale, const float* group_dequant_scale_up,
                            const float* group_quant_scale, int act_mode, float* row_scale_out, int swiglu,
                            int div_mode, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(x, w_stacked, y, M_pad, N, K);
  if (rc != ASQ_OK) return rc;
  if (!is_float_dtype(x_dtype)) return fail(ASQ_ERR_INVALID, "grouped: x dtype must be f32, f16 or bf16");
  if (num_groups < 1 || num_groups > 4096 || group_of_blk == nullptr || group_dequant_scale == nullptr)
    return fail(ASQ_ERR_INVALID, "grouped: need 1..4096 groups, the block->group table and the per-group dequant scales");
  if (M_pad % asq::BLOCK_M != 0) return fail(ASQ_ERR_INVALID, "grouped: M_pad=%lld must be a multiple of 128 (segments padded to 256 rows)", (long long)M_pad);
  if (act_mode != ASQ_ACT_ROUND && act_mode != ASQ_ACT_SCALE && act_mode != ASQ_ACT_PER_TOKEN && act_mode != ASQ_ACT_ROW_SCALE_GIVEN)
    return fail(ASQ_ERR_INVALID, "grouped: act_mode must be ROUND, SCALE, PER_TOKEN or ROW_SCALE_GIVEN");
  if (act_mode == ASQ_ACT_ROW_SCALE_GIVEN && row_scale_out == nullptr && M_pad > 0)
    return fail(ASQ_ERR_INVALID, "grouped: ASQ_ACT_ROW_SCALE_GIVEN needs the [M_pad] row scales in row_scale");
  if (act_mode == ASQ_ACT_SCALE && group_quant_scale == nullptr) return fail(ASQ_ERR_INVALID, "grouped: ASQ_ACT_SCALE needs group_quant_scale");
  if (div_mode != ASQ_DIV_RECIPROCAL && div_mode != ASQ_DIV_EXACT) return fail(ASQ_ERR_INVALID, "bad div_mode %d", div_mode);
  if (swiglu) {
    if (N % 64 != 0) return fail(ASQ_ERR_INVALID, "grouped swiglu: N=%lld (2 x intermediate) must be a multiple of 64", (long long)N);
    if (y_dtype != ASQ_BF16 && y_dtype != ASQ_F16) return fail(ASQ_ERR_INVALID, "grouped swiglu: y dtype must be f16 or bf16");
    if (group_dequant_scale_up == nullptr) return fail(ASQ_ERR_INVALID, "grouped swiglu: need the up projections' scales");
  } else {
    if (!is_float_dtype(y_dtype)) return fail(ASQ_ERR_INVALID, "grouped: y dtype must be f32, f16 or bf16");
    if ((N * (y_dtype == ASQ_F32 ? 4 : 2)) % 16 != 0) return fail(ASQ_ERR_INVALID, "grouped: output rows must be 16-byte multiples");
  }
  if (M_pad == 0) return ASQ_OK;
  if (workspace == nullptr || workspace_bytes < asq_workspace_bytes(M_pad, K))
    return fail(ASQ_ERR_WORKSPACE, "workspace needs %zu bytes, got %zu", asq_workspace_bytes(M_pad, K), workspace_bytes);
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(ASQ_ERR_INVALID, "workspace must be 1024-byte aligned");
  Workspace ws;
  ws_layout(M_pad, K, workspace, &ws);
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.a_q = ws.a_q; p.row_scale = ws.row_scale; p.sync = ws.sync;  // no stream-K in grouped launches
  if (act_mode == ASQ_ACT_ROW_SCALE_GIVEN) p.row_scale_in = row_scale_out;  // in: caller-supplied scales (tensor-parallel w2)
  else p.row_scale_out = row_scale_out;
  p.quant_scale = 1.f; p.inv_quant_scale = 1.f; p.qmax = 127.0f; p.inv_qmax = 1.0f / 127.0f;
  p.y = y; p.dequant_scale = 1.f; p.dequant_scale_up = 1.f;
  p.M = static_cast<int>(M_pad); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.x_dtype = x_dtype; p.y_dtype = y_dtype; p.mid_dtype = y_dtype; p.act_mode = act_mode; p.div_mode = div_mode;
  p.epi_kind = swiglu ? asq::EPI_SWIGLU : asq::EPI_DEQUANT;
  p.group_of_blk = group_of_blk; p.group_scale = group_dequant_scale; p.group_scale_up = group_dequant_scale_up;
  p.group_quant_scale = group_quant_scale; p.num_groups = num_groups;
  return launch_linear(false, ws.a_q, w_stacked, p, static_cast<cudaStream_t>(stream));
}

// ---- row-parallel GEMM fused with its all-reduce over peer memory
int asq_ar_buffer_bytes(int64_t M, int64_t N, int world, size_t* recv_bytes, size_t* ctl_bytes) {
  if (M <= 0 || N <= 0 || world < 2 || world > 8 || recv_bytes == nullptr || ctl_bytes == nullptr)
    return fail(ASQ_ERR_INVALID, "asq_ar_buffer_bytes: bad arguments (world must be 2..8)");
  // per tile one 128-row x 256-column 4-byte slot per CTA; pairs own 256-row tiles, so the total is the same
  const int cg = pick_cta_group(M);
  const int64_t tiles = ((M + asq::BLOCK_M * cg - 1) / (asq::BLOCK_M * cg)) * ((N + asq::TILE_N - 1) / asq::TILE_N);
  const int64_t cnt_max = (tiles + world - 1) / world;
  const int64_t slots = static_cast<int64_t>(world - 1) * cnt_max * cg;  // CTA slots
  *recv_bytes = static_cast<size_t>(slots) * asq::SK_SLOT_WORDS * 4;
  *ctl_bytes = round_up((asq::AR_FLAG_BASE + static_cast<size_t>(slots) * asq::NUM_EPI_WARPS) * 4, 1024);
  return ASQ_OK;
}

int asq_w8a8_linear_q8_allreduce(const int8_t* xq, const float* row_scale, const int8_t* w, const float* bias,
                                 void* const* y_all, int y_dtype, int64_t M, int64_t N, int64_t K,
                                 float dequant_scale, const float* col_scale, void* const* recv_all,
                                 void* const* ctl_all, int rank, int world, int partial16, void* y_multicast,
                                 void* stream) {
  if (world < 2 || world > 8 || rank < 0 || rank >= world || y_all == nullptr || recv_all == nullptr || ctl_all == nullptr)
    return fail(ASQ_ERR_INVALID, "allreduce: bad rank %d / world %d or null pointer tables", rank, world);
  for (int r = 0; r < world; ++r)
    if (y_all[r] == nullptr || recv_all[r] == nullptr || ctl_all[r] == nullptr || (reinterpret_cast<uintptr_t>(y_all[r]) & 15) ||
        (reinterpret_cast<uintptr_t>(recv_all[r]) & 15))
      return fail(ASQ_ERR_INVALID, "allreduce: rank %d buffers must be non-null and 16-byte aligned", r);
  int rc = check_common(xq, w, y_all[rank], M, N, K);
  if (rc != ASQ_OK) return rc;
  if (y_dtype != ASQ_BF16 && y_dtype != ASQ_F16) return fail(ASQ_ERR_INVALID, "allreduce: y dtype must be f16 or bf16");
  if (N % 8 != 0) return fail(ASQ_ERR_INVALID, "allreduce: N=%lld must be a multiple of 8 (16-byte output rows)", (long long)N);
  if (M == 0) return fail(ASQ_ERR_INVALID, "allreduce: M must be positive (every rank has to launch)");
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  p.y = y_all[rank]; p.bias = bias; p.col_scale = col_scale; p.dequant_scale = dequant_scale;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.y_dtype = y_dtype; p.epi_kind = asq::EPI_DEQUANT;
  p.act_mode = row_scale != nullptr ? ASQ_ACT_ROW_SCALE_GIVEN : ASQ_ACT_ROUND;
  p.row_scale = const_cast<float*>(row_scale);
  p.ar_world = world; p.ar_rank = rank;
  p.ar_partial16 = partial16 ? 1 : 0;
  p.ar_y_mc = static_cast<uint8_t*>(y_multicast);
  if (y_multicast != nullptr && (reinterpret_cast<uintptr_t>(y_multicast) & 15))
    return fail(ASQ_ERR_INVALID, "allreduce: the multicast address must be 16-byte aligned");
  void* peers[7];
  int n = 0;
  for (int r = 0; r < world; ++r) {
    p.ar_ctl[r] = static_cast<uint32_t*>(ctl_all[r]);
    p.ar_recv[r] = static_cast<uint32_t*>(recv_all[r]);
    if (r != rank) peers[n++] = y_all[r];
  }
  return launch_linear(false, xq, w, p, static_cast<cudaStream_t>(stream), peers);
}

int asq_q8_linear_allreduce_nvls(const void* xq, int fp8, const float* row_scale, const void* w, const float* bias,
                                 void* partial_local, const void* partial_mc, void* y_mc, int y_dtype, int64_t M,
                                 int64_t N, int64_t K, float dequant_scale, const float* col_scale,
                                 void* const* ctl_all, void* counters_local, const void* counters_mc,
                                 size_t counter_bank_bytes, int launch_parity, int rank, int world, void* stream) {
  if (world < 2 || world > 8 || rank < 0 || rank >= world || ctl_all == nullptr)
    return fail(ASQ_ERR_INVALID, "nvls allreduce: bad rank %d / world %d or null control table", rank, world);
  if (counters_local == nullptr || counters_mc == nullptr || counter_bank_bytes % 16 != 0 || (launch_parity != 0 && launch_parity != 1))
    return fail(ASQ_ERR_INVALID, "nvls allreduce: bad counter banks");
  {
    const int64_t slabs = ((M + asq::NVLS_SLAB_ROWS - 1) / asq::NVLS_SLAB_ROWS + 3) * ((N + asq::TILE_N - 1) / asq::TILE_N);
    if (static_cast<size_t>(slabs) * 4 > counter_bank_bytes)
      return fail(ASQ_ERR_WORKSPACE, "nvls allreduce: [%lld, %lld] needs %lld counter bytes per bank, got %zu", (long long)M, (long long)N,
                  (long long)slabs * 4, counter_bank_bytes);
  }
  for (int r = 0; r < world; ++r)
    if (ctl_all[r] == nullptr) return fail(ASQ_ERR_INVALID, "nvls allreduce: rank %d control buffer is null", r);
  if (partial_local == nullptr || partial_mc == nullptr || y_mc == nullptr || (reinterpret_cast<uintptr_t>(partial_mc) & 15) ||
      (reinterpret_cast<uintptr_t>(y_mc) & 15))
    return fail(ASQ_ERR_INVALID, "nvls allreduce: the partial / output buffers must be non-null and 16-byte aligned");
  int rc = check_common(xq, w, partial_local, M, N, K);
  if (rc != ASQ_OK) return rc;
  if (y_dtype != ASQ_BF16 && y_dtype != ASQ_F16) return fail(ASQ_ERR_INVALID, "nvls allreduce: y dtype must be f16 or bf16");
  if (N % 8 != 0) return fail(ASQ_ERR_INVALID, "nvls allreduce: N=%lld must be a multiple of 8 (16-byte output rows)", (long long)N);
  if (M == 0) return fail(ASQ_ERR_INVALID, "nvls allreduce: M must be positive (every rank has to launch)");
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  p.y = partial_local; p.bias = bias; p.col_scale = col_scale; p.dequant_scale = dequant_scale;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.y_dtype = y_dtype; p.epi_kind = asq::EPI_DEQUANT;
  p.act_mode = row_scale != nullptr ? ASQ_ACT_ROW_SCALE_GIVEN : ASQ_ACT_ROUND;
  p.row_scale = const_cast<float*>(row_scale);
  p.ar_world = world; p.ar_rank = rank;
  p.nvls_p_mc = static_cast<const uint8_t*>(partial_mc);
  p.nvls_y_mc = static_cast<uint8_t*>(y_mc);
  p.nvls_ctr = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(counters_local) + launch_parity * counter_bank_bytes);
  p.nvls_ctr_next = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(counters_local) + (1 - launch_parity) * counter_bank_bytes);
  p.nvls_ctr_mc = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(counters_mc) + launch_parity * counter_bank_bytes);
  p.nvls_ctr_words = static_cast<int>(counter_bank_bytes / 4);
  for (int r = 0; r < world; ++r) p.ar_ctl[r] = static_cast<uint32_t*>(ctl_all[r]);
  return launch_linear(fp8 != 0, xq, w, p, static_cast<cudaStream_t>(stream));
}

// ---- device memory shared between the processes of one node (cudaMalloc + CUDA IPC), used for the buffers above
int asq_dev_alloc(size_t bytes, void** ptr) {
  if (ptr == nullptr || bytes == 0) return fail(ASQ_ERR_INVALID, "asq_dev_alloc: bad arguments");
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ASQ_ERR_CUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
  e = cudaMemset(*ptr, 0, bytes);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ASQ_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e)); }
  return ASQ_OK;
}
int asq_dev_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ASQ_ERR_CUDA, "cudaFree: %s", cudaGetErrorString(e)); }
  return ASQ_OK;
}
int asq_ipc_export(const void* dev_ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  if (dev_ptr == nullptr || handle64 == nullptr) return fail(ASQ_ERR_INVALID, "asq_ipc_export: null argument");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr));
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ASQ_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
  memcpy(handle64, &h, 64);
  return ASQ_OK;
}
int asq_ipc_open(const void* handle64, void** dev_ptr) {
  if (dev_ptr == nullptr || handle64 == nullptr) return fail(ASQ_ERR_INVALID, "asq_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ASQ_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
  return ASQ_OK;
}
int asq_ipc_close(void* dev_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ASQ_ERR_CUDA, "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e)); }
  return ASQ_OK;
}

int asq_i8gemm_o32(const int8_t* a, const int8_t* w, int32_t* c, int64_t M, int64_t N, int64_t K,
                   void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(a, w, c, M, N, K);
  if (rc != ASQ_OK || M == 0) return rc;
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  rc = attach_streamk(p, workspace, workspace_bytes);
  if (rc != ASQ_OK) return rc;
  p.y = c; p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.y_dtype = ASQ_I32; p.epi_kind = asq::EPI_RAW_I32; p.act_mode = ASQ_ACT_ROUND;
  return launch_linear(false, a, w, p, static_cast<cudaStream_t>(stream));
}

int asq_i8gemm_epi(const int8_t* a, const int8_t* w, const void* bias, int bias_dtype, void* y, int y_dtype,
                   int64_t M, int64_t N, int64_t K, float alpha, float beta, int flags,
                   void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(a, w, y, M, N, K);
  if (rc != ASQ_OK || M == 0) return rc;
  if (y_dtype != ASQ_I8 && y_dtype != ASQ_I32 && !is_float_dtype(y_dtype)) return fail(ASQ_ERR_INVALID, "bad y_dtype %d", y_dtype);
  if (bias != nullptr && bias_dtype != ASQ_I8 && bias_dtype != ASQ_I32 && bias_dtype != ASQ_F32)
    return fail(ASQ_ERR_INVALID, "bias dtype must be i8, i32 or f32");
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  p.y = y; p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.y_dtype = y_dtype; p.epi_kind = asq::EPI_ALPHA_BETA; p.act_mode = ASQ_ACT_ROUND;
  p.alpha = alpha; p.beta = beta; p.bias_any = bias; p.bias_dtype = bias_dtype; p.flags = flags;
  rc = attach_streamk(p, workspace, workspace_bytes);
  if (rc != ASQ_OK) return rc;
  return launch_linear(false, a, w, p, static_cast<cudaStream_t>(stream));
}

int asq_quantize_act(const void* x, int x_dtype, void* q, float* row_scale, int64_t M, int64_t K, int act_mode,
                     float quant_scale, int div_mode, int fp8, void* stream) {
  if (M < 0 || K <= 0 || K % 16 != 0) return fail(ASQ_ERR_INVALID, "bad shape M=%lld K=%lld (K %% 16 == 0 required)", (long long)M, (long long)K);
  if (!is_float_dtype(x_dtype)) return fail(ASQ_ERR_INVALID, "x dtype must be f32, f16 or bf16");
  if (act_mode != ASQ_ACT_ROUND && act_mode != ASQ_ACT_SCALE && act_mode != ASQ_ACT_PER_TOKEN &&
      act_mode != ASQ_ACT_ROW_SCALE_GIVEN)
    return fail(act_mode == ASQ_ACT_PER_TENSOR_DYNAMIC ? ASQ_ERR_UNSUPPORTED : ASQ_ERR_INVALID, "unsupported act_mode %d", act_mode);
  if (M == 0) return ASQ_OK;
  if (x == nullptr || q == nullptr) return fail(ASQ_ERR_INVALID, "null pointer argument");
  if ((act_mode == ASQ_ACT_PER_TOKEN || act_mode == ASQ_ACT_ROW_SCALE_GIVEN) && row_scale == nullptr)
    return fail(ASQ_ERR_INVALID, "row_scale required for per-token modes");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(q) & 15))
    return fail(ASQ_ERR_INVALID, "x and q must be 16-byte aligned");
  int dev;
  DeviceState* st;
  int rc = get_device(&dev, &st);
  if (rc != ASQ_OK) return rc;
  asq::LinearParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.a_q = static_cast<uint8_t*>(q);
  if (act_mode == ASQ_ACT_ROW_SCALE_GIVEN) p.row_scale_in = row_scale;  // input: scales to quantise with
  else p.row_scale = row_scale;
  p.quant_scale = quant_scale; p.inv_quant_scale = 1.0f / quant_scale;
  p.qmax = fp8 ? 448.0f : 127.0f; p.inv_qmax = 1.0f / p.qmax;
  p.M = static_cast<int>(M); p.K = static_cast<int>(K);
  p.x_dtype = x_dtype; p.act_mode = act_mode; p.div_mode = div_mode;
  const int rows_per_block = 8;
  long long blocks = (M + rows_per_block - 1) / rows_per_block;
  const long long cap = static_cast<long long>(st->sm_count) * 8;
  if (blocks > cap) blocks = cap;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(blocks), 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = asq_pdl_enabled() ? 1 : 0;
  cudaError_t e = fp8 ? cudaLaunchKernelEx(&cfg, asq::asq_quantize_kernel<true>, p)
                      : cudaLaunchKernelEx(&cfg, asq::asq_quantize_kernel<false>, p);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ASQ_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return ASQ_OK;
}

}  // extern "C"
#endif  // ASQ_TU host part
