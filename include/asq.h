/*
 * asq.h — C ABI of the B200-native (sm_100a) SmoothQuant W8A8 / FP8 linear path.
 *
 * This shared library (libasq_b200.so) is the drop-in boundary for the native
 * extension of AniZpZ/AutoSmoothQuant (`autosmoothquant._CUDA`) plus the eager
 * quantize / dequantize launches the reference's Python modules run around it.
 * Plain pointers and sizes only: no torch types cross this boundary.
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   - csrc/int8gemm/bindings.cpp:69-84   I8CUGEMM::linear_a8_w8_o32_   -> asq_i8gemm_o32
 *   - csrc/int8gemm/bindings.cpp:52-67   I8CUGEMM::linear_a8_w8_o32    -> asq_i8gemm_o32
 *   - csrc/int8gemm/bindings.cpp:86-142  linear_a8_w8_o8[_], _b8_o8_   -> asq_i8gemm_epi
 *   - autosmoothquant/layers/nn/linear.py:83-106   W8A8BFP32OFP32Linear.forward
 *   - autosmoothquant/layers/nn/linear.py:172-208  W8A8BFP32OFP32QKVLinear.forward
 *   - autosmoothquant/layers/nn/linear.py:278-302  ...LinearWithQuantScale.forward
 *                                                                      -> asq_w8a8_linear
 *   - autosmoothquant/layers/nn/linear.py:336-369, 413-427, 551-566
 *       easy_fp8_gemm / FP8LinearDynamic.forward / FP8LinearStatic.forward
 *                                                                      -> asq_fp8_linear
 *   - autosmoothquant/layers/functional/quantization.py:173-191, 208-211 and the
 *     inline quantisation in linear.py:88-95,283-292                   -> asq_quantize_act
 *
 * Conventions
 *   - Every entry point returns ASQ_OK (0) or a negative asq_status and never
 *     throws; asq_last_error() returns a thread-local description of the last
 *     failure on the calling thread.
 *   - All device pointers must belong to the CUDA device that is current on the
 *     calling thread; `stream` is a cudaStream_t passed as void*. Calls are
 *     asynchronous with respect to the host, stateless and re-entrant.
 *   - Matrices are dense row-major: x [M,K], w [N,K] (K contiguous, i.e. the
 *     "TN" GEMM of cublasINT8MMWrapper.cc:246-253), y [M,N].
 *   - K must be a multiple of 16 (TMA row pitch); M, N arbitrary (M may be 0).
 *   - There is no CPU fallback: without a CUDA device every compute entry point
 *     returns ASQ_ERR_CUDA.
 */
#ifndef ASQ_H_
#define ASQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASQ_VERSION 100 /* major*100 + minor */

typedef enum asq_status {
  ASQ_OK = 0,
  ASQ_ERR_INVALID = -1,     /* bad argument (null pointer, misaligned, bad enum, K%16!=0) */
  ASQ_ERR_UNSUPPORTED = -2, /* valid request this build does not implement */
  ASQ_ERR_CUDA = -3,        /* CUDA runtime / driver error, or no sm_100 device */
  ASQ_ERR_WORKSPACE = -4    /* workspace too small or not provided */
} asq_status;

/* Element types of x / y buffers. */
typedef enum asq_dtype {
  ASQ_F32 = 0,
  ASQ_F16 = 1,
  ASQ_BF16 = 2,
  ASQ_I32 = 3, /* y only: raw accumulator (int8 path) */
  ASQ_I8 = 4   /* y only: requantised output (asq_i8gemm_o8) */
} asq_dtype;

/* Activation quantisation applied by the fused prologue.
 *   ASQ_ACT_ROUND      q = sat(rint(x))                      linear.py:95   (scale folded into the norm)
 *   ASQ_ACT_SCALE      q = sat(rint(T(x / quant_scale)))     linear.py:290-292 (division rounded to x's dtype T)
 *   ASQ_ACT_PER_TOKEN  s[m] = f32(T(max_k|x[m,k]|) / T(qmax)); q = sat(rint(f32(x)/s[m]))   linear.py:88-92
 *   ASQ_ACT_PER_TENSOR_DYNAMIC (fp8 only) s = T(max|x|)/T(448) over the whole tensor, q = T(x/s)
 *                                                            quantization.py:144-170
 * qmax is 127 for int8 and 448 for e4m3.                                                   */
typedef enum asq_act_mode {
  ASQ_ACT_ROUND = 0,
  ASQ_ACT_SCALE = 1,
  ASQ_ACT_PER_TOKEN = 2,
  ASQ_ACT_PER_TENSOR_DYNAMIC = 3,
  ASQ_ACT_ROW_SCALE_GIVEN = 4, /* per-token arithmetic with caller-supplied scales s[m] (row-parallel
                                  tensor parallelism: the global row absmax is all-reduced first) */
  ASQ_ACT_RMSNORM = 5          /* (asq_w8a8_rmsnorm_* entry points) HF RMSNorm with the folded 1/input_scale as the
                                  prologue: q = sat(rint(T(norm_w * T(x * rsqrt(mean(x^2) + eps)))))  models/llama.py:27-37 */
} asq_act_mode;

/* How `tensor / python_scalar` is evaluated.  torch on CUDA multiplies by the
 * fp32 reciprocal of a scalar divisor, torch on CPU performs a true division;
 * the two differ in the last bit.  ASQ_DIV_RECIPROCAL reproduces the reference
 * running on its only supported device (CUDA); ASQ_DIV_EXACT reproduces the
 * reference's Python executed on CPU (what the oracle is pinned against). */
typedef enum asq_div_mode { ASQ_DIV_RECIPROCAL = 0, ASQ_DIV_EXACT = 1 } asq_div_mode;

/* Epilogue flags for asq_i8gemm_epi (csrc/kernels/linear.cu variants). */
typedef enum asq_epi_flags {
  ASQ_EPI_RELU = 1 /* clamp negative results to zero before the output conversion */
} asq_epi_flags;

int asq_version(void);
const char* asq_last_error(void);

/* 1 if the current CUDA device can run the kernels (compute capability 10.0), else 0. */
int asq_device_supported(void);

/* Bytes of scratch for an [M,K] activation: the phase counters, the stream-K
 * region (handshake words + partial accumulators, ~20 MB), M fp32 row scales and
 * the int8/e4m3 copy of x.  asq_workspace_bytes(0, 0) is what the GEMM-only entry
 * points can use (their workspace is optional: without it decode-sized problems
 * are not split along K).  The buffer must be 1024-byte aligned device memory,
 * zero-filled once when allocated (the kernels leave every word they rely on at
 * zero again), and must not be shared by calls that may run concurrently. */
size_t asq_workspace_bytes(int64_t M, int64_t K);

/* y = dequant( quant(x) . w^T ) [+ bias], one launch.
 *   x         [M,K] x_dtype in {F32,F16,BF16}
 *   w         [N,K] int8
 *   bias      [N] fp32 or NULL
 *   y         [M,N] y_dtype in {F32,F16,BF16}
 *   dequant_scale   scalar applied to every column, used when col_scale == NULL
 *   col_scale [N] fp32 or NULL: per-output-column dequant scale (the QKV variant's
 *             piecewise-constant q/k/v scales, linear.py:197-200)
 *   row_scale_out [M] fp32 or NULL: receives the per-token scales (ASQ_ACT_PER_TOKEN); with
 *             ASQ_ACT_ROW_SCALE_GIVEN it is an INPUT holding the scales to quantise with
 * Arithmetic (bit-exact with the reference's fp32 epilogue):
 *   f = col_scale ? col_scale[n] : dequant_scale;  per-token: f = f * s[m]
 *   y[m,n] = T( f * f32(acc[m,n]) (+ bias[n]) )     acc = exact int32 dot product       */
int asq_w8a8_linear(const void* x, int x_dtype, const int8_t* w, const float* bias,
                    void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                    int act_mode, float quant_scale, float dequant_scale,
                    const float* col_scale, float* row_scale_out, int div_mode,
                    void* workspace, size_t workspace_bytes, void* stream);

/* FP8-e4m3 twin: y = T( (sum_k q[m,k]*w[n,k]) * (s_x * w_scale) (+ bias) ), fp32 accumulate
 * on the tensor cores (the reference dequantises both operands and calls F.linear,
 * linear.py:363-368).
 *   w  [N,K] e4m3 bytes
 *   act_mode  ASQ_ACT_PER_TOKEN (FP8LinearDynamic per-token), ASQ_ACT_SCALE (FP8LinearStatic, in_scale),
 *             ASQ_ACT_PER_TENSOR_DYNAMIC (FP8LinearDynamic's other branch: whole-tensor absmax, computed
 *             in-kernel with a grid-wide reduction), ASQ_ACT_ROW_SCALE_GIVEN
 *   out_scale != 0: the output is additionally fake-quantised through e4m3 with this scale
 *             (FP8LinearStatic with a truthy output_scale, linear.py:562-564)
 *   row_scale_out [M] or NULL receives the activation scale used for every row */
int asq_fp8_linear(const void* x, int x_dtype, const uint8_t* w_e4m3, const float* bias,
                   void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                   int act_mode, float in_scale, float w_scale, float out_scale,
                   float* row_scale_out, int div_mode,
                   void* workspace, size_t workspace_bytes, void* stream);

/* asq_fp8_linear with one dequantisation scale per OUTPUT COLUMN: w_col_scale [N] fp32 on the device replaces the
 * scalar w_scale, y = T( acc * (w_col_scale[n] * s_x[m]) (+ bias) ).  Horizontally fused FP8 projections that share
 * their input (q|k|v, gate|up: weights concatenated along N, the activation quantised once) keep the per-tensor
 * weight scale of each block — the FP8 counterpart of W8A8BFP32OFP32QKVLinear (linear.py:132-245, 197-200).
 * Dynamic activation scales only: ASQ_ACT_PER_TOKEN, ASQ_ACT_PER_TENSOR_DYNAMIC, ASQ_ACT_ROW_SCALE_GIVEN. */
int asq_fp8_linear_cs(const void* x, int x_dtype, const uint8_t* w_e4m3, const float* bias,
                      void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                      int act_mode, const float* w_col_scale,
                      float* row_scale_out, int div_mode,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Same GEMM + dequant epilogue for activations that are ALREADY int8 (emitted by a fused producer such as
 * asq_add_rmsnorm_quant / asq_silu_mul_quant): no prologue runs.  row_scale [M] fp32 or NULL supplies
 * per-token scales for the epilogue. */
int asq_w8a8_linear_q8(const int8_t* xq, const float* row_scale, const int8_t* w, const float* bias,
                       void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                       float dequant_scale, const float* col_scale,
                       void* workspace /* nullable */, size_t workspace_bytes, void* stream);

/* asq_w8a8_linear / asq_w8a8_linear_q8 with the decoder block's residual add in the epilogue:
 *   y = T(residual + T(linear output))     residual [M,N] of y's 16-bit dtype (may alias y), y_dtype F16 | BF16
 * i.e. `hidden = residual + o_proj(...)` / `residual + down_proj(...)` of the HF decoder layers the reference's
 * model classes inherit (models/llama.py:218) without the separate add (or the add inside the next norm kernel). */
int asq_w8a8_linear_res(const void* x, int x_dtype, const int8_t* w, const float* bias, const void* residual,
                        void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                        int act_mode, float quant_scale, float dequant_scale,
                        const float* col_scale, float* row_scale_out, int div_mode,
                        void* workspace, size_t workspace_bytes, void* stream);
int asq_w8a8_linear_q8_res(const int8_t* xq, const float* row_scale, const int8_t* w, const float* bias,
                           const void* residual, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                           float dequant_scale, const float* col_scale, void* stream);

/* asq_w8a8_linear_q8 for the fused q|k|v projection with HF rotate-half RoPE applied in the epilogue (what
 * asq_rope_inplace does as a second pass over the same tensor): columns [0, rope_cols) — the q and k heads —
 * are rotated per head, position = row % S; the remaining columns (v) are plain dequantised outputs.
 * head_dim must be 128, N and rope_cols multiples of 128, y_dtype F16 | BF16.  The tables hold the HF
 * [S, head_dim] cos / sin values in y's dtype, re-laid-out in 8-column blocks as [head_dim/8][S][8] (entry
 * (pos, col) at ((col/8)*S + pos)*8 + col%8) so that the 32 rows a warp owns read contiguous memory.
 * halves_equal != 0: the caller guarantees table[:, d] == table[:, d + head_dim/2] (true for HF's
 * emb = cat(freqs, freqs)), and only the first half is read. */
int asq_w8a8_linear_q8_rope(const int8_t* xq, const float* row_scale, const int8_t* w, const float* bias,
                            void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                            float dequant_scale, const float* col_scale,
                            const void* cos_table, const void* sin_table, int64_t S, int64_t rope_cols,
                            int64_t head_dim, int halves_equal, void* stream);

/* The q|k|v and gate|up launches with the preceding RMSNorm as their prologue (ASQ_ACT_RMSNORM): x [M,K] is the
 * residual stream (F16 | BF16), norm_weight [K] the norm weight with 1/input_scale folded in (models/llama.py:
 * 27-37, 326-339); q = sat(rint(T(norm_weight * T(x * rsqrt(mean(x^2) + eps))))) feeds the GEMM, bit-identical to
 * asq_add_rmsnorm_quant followed by the int8-in entry points.  Epilogues as in asq_w8a8_linear_q8_rope (cos_table
 * NULL: plain dequant, y in x's dtype) and asq_w8a8_gateup_swiglu_q8.  Workspace as for asq_w8a8_linear. */
int asq_w8a8_rmsnorm_linear_rope(const void* x, int x_dtype, const void* norm_weight, float eps, const int8_t* w,
                                 const float* bias, void* y, int64_t M, int64_t N, int64_t K, float dequant_scale,
                                 const float* col_scale, const void* cos_table, const void* sin_table, int64_t S,
                                 int64_t rope_cols, int64_t head_dim, int halves_equal, void* workspace,
                                 size_t workspace_bytes, void* stream);
int asq_w8a8_rmsnorm_gateup_swiglu(const void* x, int x_dtype, const void* norm_weight, float eps, const int8_t* w_il,
                                   const float* bias_il, void* out, int out_dtype, int64_t M, int64_t N, int64_t K,
                                   float gate_dequant_scale, float up_dequant_scale, const float* col_scale_il,
                                   float out_quant_scale, int div_mode, void* workspace, size_t workspace_bytes,
                                   void* stream);

/* gate|up projection with the SwiGLU product (and the next Linear's activation quantisation) in the GEMM
 * epilogue: the [M, 2I] gate|up tensor of HF LlamaMLP.forward (borrowed at models/llama.py:218) never
 * reaches HBM.  Replaces asq_w8a8_linear_q8 (fused gate|up) + asq_silu_mul_quant byte for byte.
 *   w_il  [N = 2I, K] int8, rows INTERLEAVED in blocks of 32: rows [64b, 64b+32) = gate rows [32b, 32b+32),
 *         rows [64b+32, 64b+64) = up rows [32b, 32b+32)   (a load-time re-layout; I % 32 == 0)
 *   gate_dequant_scale / up_dequant_scale  the two projections keep their own scalar scale (as the blocks of
 *         the fused module do, linear.py:197-200); col_scale_il [N] fp32, if not NULL, replaces both with a
 *         per-column vector; bias_il [N] fp32 or NULL; vectors are in the same interleaved order as w_il
 *   gate = T(f * acc_g (+ b)), up = T(f * acc_u (+ b)), a = T(T(silu(gate)) * up), T = mid_dtype (F16 | BF16)
 *   out   [M, I]: out_dtype == mid_dtype -> a;  ASQ_I8 -> sat(rint(T(a / out_quant_scale))), which is what
 *         W8A8BFP32OFP32LinearWithQuantScale (per-tensor) derives from a (linear.py:290-292)
 *   row_scale [M] fp32 or NULL: per-token scales of xq                                            */
int asq_w8a8_gateup_swiglu_q8(const int8_t* xq, const float* row_scale, const int8_t* w_il, const float* bias_il,
                              void* out, int out_dtype, int mid_dtype, int64_t M, int64_t N, int64_t K,
                              float gate_dequant_scale, float up_dequant_scale, const float* col_scale_il,
                              float out_quant_scale, int div_mode, void* stream);

/* ---- Grouped linear: all experts of a Mixtral sparse-MoE block in ONE launch.
 * Replaces the per-expert loop of HF MixtralSparseMoeBlock.forward (borrowed at models/mixtral.py:145) over
 * Int8MixtralBlockSparseTop2MLP (models/mixtral.py:94-121): the caller sorts the routed token rows by expert
 * into x [M_pad, K], padding every expert's segment with zero rows to a multiple of 256; group_of_blk[i]
 * (device, int32) names the expert of rows [128 i, 128 i + 128), -1 for unused trailing blocks (skipped).
 *   w_stacked [G * N, K] int8: expert g's weight in rows [g N, (g+1) N); group_dequant_scale [G] (device fp32)
 *   act_mode ROUND | SCALE (per-expert group_quant_scale [G], LinearWithQuantScale) | PER_TOKEN |
 *   ROW_SCALE_GIVEN (row_scale_out then is an INPUT: [M_pad] fp32 scales to quantise with, the tensor-parallel
 *   w2 whose row absmax spans all ranks' ffn slices)
 *   swiglu != 0: w_stacked holds every expert's w1|w3 in the interleaved layout of asq_w8a8_gateup_swiglu_q8
 *   (N = 2 * ffn), group_dequant_scale / _up are the w1 / w3 scales, y [M_pad, N/2] = T(T(silu(w1 x)) * w3 x).
 * Row results are independent of the other rows, so every real row equals what the expert's own module
 * returns for that token, bit for bit. */
int asq_w8a8_grouped_linear(const void* x, int x_dtype, const int8_t* w_stacked, void* y, int y_dtype,
                            int64_t M_pad, int64_t N, int64_t K, int num_groups, const int32_t* group_of_blk,
                            const float* group_dequant_scale, const float* group_dequant_scale_up,
                            const float* group_quant_scale, int act_mode, float* row_scale_out, int swiglu,
                            int div_mode, void* workspace, size_t workspace_bytes, void* stream);

/* ---- Row-parallel linear fused with its all-reduce over NVLink peer memory (SURVEY 8(e) "fusion target").
 * One launch per rank replaces asq_w8a8_linear_q8 + ncclAllReduce(sum): every rank computes all output tiles
 * over its K shard; tile t is owned by rank t % world; non-owners store their raw int32 accumulators into the
 * owner's receive buffer (P2P stores) and raise a flag; the owner adds them (exact integer sum, so the result
 * is bit-identical to the unsharded module), applies the fp32 dequant epilogue (+ bias) and TMA-stores the
 * finished 16-bit tile into the y buffer of every rank.  Ranks walk the tiles they do not own first, so the
 * exchange overlaps the remaining math tile by tile.  No NCCL call is involved.
 *   y_all / recv_all / ctl_all   [world] device pointers valid in THIS process: entry r is rank r's output,
 *       receive and control buffer (own buffers from asq_dev_alloc, the peers' via asq_ipc_open).  Sizes from
 *       asq_ar_buffer_bytes for the largest (M, N) used; zero-filled at allocation, never reset by the host.
 *   bias (fp32 [N]) must be given on EVERY rank (the tile owner applies it); row_scale = global per-token scales.
 *   partial16 != 0: partials travel dequantised and rounded to y's 16-bit dtype instead (half the bytes; the
 *       owner adds them in fp32 in fixed rank order and rounds once: at world 2 exactly the result of a bf16
 *       ncclAllReduce of the per-rank outputs).  In this mode bias must be given on ONE rank only.
 *   y_multicast != NULL: the NVLS multicast address bound to all ranks' y buffers (cuMulticast* / torch
 *       symmetric memory); finished tiles are then broadcast with multimem.st — one store reaches every rank
 *       through the switch — instead of one TMA store per peer.
 * Every rank of the group must issue the same sequence of these calls (same shapes); a launch returns only
 * after every peer has finished writing this rank's y.  y may be reused by the launch after the next one. */
int asq_ar_buffer_bytes(int64_t M, int64_t N, int world, size_t* recv_bytes, size_t* ctl_bytes);
int asq_w8a8_linear_q8_allreduce(const int8_t* xq, const float* row_scale, const int8_t* w, const float* bias,
                                 void* const* y_all, int y_dtype, int64_t M, int64_t N, int64_t K,
                                 float dequant_scale, const float* col_scale, void* const* recv_all,
                                 void* const* ctl_all, int rank, int world, int partial16, void* y_multicast,
                                 void* stream);

/* The same row-parallel linear with the sum taken INSIDE THE NVSWITCH (NVLS): the collective an NCCL NVLS
 * all-reduce performs, fused into the GEMM launch.  Every rank's epilogue TMA-stores its dequantised 16-bit
 * partial tiles (T(f * acc (+ bias)), the arithmetic of asq_w8a8_linear_q8) into its OWN slice of a symmetric
 * allocation and bumps a counter on the tile's owner (rank = tile % world); reducer CTAs appended to the same
 * grid wait for the counters of the tiles their rank owns, read them through the allocation's multicast address
 * with multimem.ld_reduce (the switch adds the `world` copies) and write the sums to every rank's output with
 * multimem.st.  All ranks walk the tiles in the same order, so the reduction of one round of tiles overlaps the
 * MMAs of the next; nothing but 16-byte multimem traffic crosses NVLink.  Measured on B200: the switch's 16-bit
 * sum is within one ulp of the exactly rounded sum (not always equal to it), deterministic, identical on all ranks.
 *   xq [M, K/world], w [N, K/world]: int8 (fp8 == 0) or e4m3 (fp8 != 0: kind::f8f6f4, fp32 accumulate — the
 *       tensor-parallel FP8LinearDynamic of BASELINE config 5); row_scale = GLOBAL per-token scales or NULL
 *   bias on ONE rank only (it is part of that rank's partial), as for GEMM + ncclAllReduce
 *   partial_local: this rank's partial buffer [M, N] 16-bit; partial_mc / y_mc: multicast addresses of all
 *       ranks' partial / output buffers (cuMulticast* or torch symmetric memory, same offset on every rank)
 *   ctl_all [world]: the control buffers of asq_ar_buffer_bytes (own + asq_ipc_open'ed), zero-filled once (they
 *       carry only the end-of-launch handshake here)
 *   counters_local / counters_mc: this rank's copy and the multicast address of 2 x counter_bank_bytes of "slab
 *       landed" counters in the same symmetric allocation, zero-filled once; a rank bumps only its LOCAL copy, a
 *       reducer reads the sum over all ranks with one multimem.ld_reduce; launch_parity (0 / 1, alternating per
 *       launch on this group) selects the bank, the other bank is cleared for the next launch.  A bank needs
 *       4 bytes per (64-row slab, 256-column tile) of the largest [M, N].
 * Numerics: those of asq_w8a8_linear_q8 followed by an NVLS ncclAllReduce in T.  Every rank must issue the
 * same sequence of calls; a launch returns only after every peer has finished writing this rank's output. */
int asq_q8_linear_allreduce_nvls(const void* xq, int fp8, const float* row_scale, const void* w, const float* bias,
                                 void* partial_local, const void* partial_mc, void* y_mc, int y_dtype, int64_t M,
                                 int64_t N, int64_t K, float dequant_scale, const float* col_scale,
                                 void* const* ctl_all, void* counters_local, const void* counters_mc,
                                 size_t counter_bank_bytes, int launch_parity, int rank, int world, void* stream);

/* Measurement aid, not on the product path: drives the NVLS data path with the access pattern of the kernel above
 * (16-byte multimem.ld_reduce / multimem.st, `unroll` requests in flight per thread, ctas x threads threads) over
 * `bytes` bytes of two multicast-mapped buffers.  mode 0: dst = sum over ranks of src, 1: ld_reduce only, 2:
 * multimem.st only.  Gives the measured ceiling the collective half of the fused kernel is reported against. */
int asq_nvls_probe(const void* src_mc, void* dst_mc, size_t bytes, int ctas, int threads, int unroll, int mode,
                   void* sink, void* stream);

/* Zero-filled cudaMalloc memory and CUDA IPC handles (64 bytes) to map it into the other ranks' processes. */
int asq_dev_alloc(size_t bytes, void** ptr);
int asq_dev_free(void* ptr);
int asq_ipc_export(const void* dev_ptr, void* handle64);
int asq_ipc_open(const void* handle64, void** dev_ptr);
int asq_ipc_close(void* dev_ptr);

/* c[M,N] (int32) = a[M,K] (int8) . w[N,K]^T (int8), exact.  Drop-in for
 * I8CUGEMM::linear_a8_w8_o32_ and the exactness tap of the fused kernels. */
int asq_i8gemm_o32(const int8_t* a, const int8_t* w, int32_t* c,
                   int64_t M, int64_t N, int64_t K,
                   void* workspace /* nullable */, size_t workspace_bytes, void* stream);

/* INT8-in GEMM with a scaling epilogue (the o8 methods of I8CUGEMM and the
 * csrc/kernels/linear.cu variants):
 *   v = alpha * f32(acc) + beta * bias            (bias [N]: int8, int32 or fp32 per bias_dtype, or NULL)
 *   y = y_dtype == I8 ? sat_i8(rint(v)) : y_dtype == I32 ? rint(v) : v     (ReLU first if flagged) */
int asq_i8gemm_epi(const int8_t* a, const int8_t* w, const void* bias, int bias_dtype,
                   void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                   float alpha, float beta, int flags, 
                   void* workspace /* nullable */, size_t workspace_bytes, void* stream);

/* Batched INT8 GEMM, the csrc/kernels/bmm.cu family (bmm_s8t_s8n_{s8t,f32t,s32t}, layers/nn/bmm.py):
 *   a [batch, M, K] int8 row-major, w [batch, N, K] int8 ("column-major B"), c [batch, M, N]
 *   c_dtype ASQ_I32: raw accumulators (alpha ignored); ASQ_F32: alpha * f32(acc); ASQ_I8: sat_i8(rint(alpha * f32(acc))).
 * One launch when M is a multiple of the tile height (256 rows; 128 when batch*M <= 128), else one per batch entry. */
int asq_i8bmm(const int8_t* a, const int8_t* w, void* c, int c_dtype, int64_t batch, int64_t M, int64_t N, int64_t K,
              float alpha, void* stream);

/* Debug / parity tap of the fused prologue: writes the quantised activations
 * (int8, or e4m3 bytes when fp8 != 0) and, for per-token, the row scales. */
int asq_quantize_act(const void* x, int x_dtype, void* q, float* row_scale,
                     int64_t M, int64_t K, int act_mode, float quant_scale,
                     int div_mode, int fp8, void* stream);

/* ---- producer-side fusions (reference intent: layers/nn/fused.py:2-25,
 * csrc/kernels/fused.cu:5-24, norm folding models/llama.py:27-37,326-339).  dtype in {F16, BF16}. ---- */

/* x_out = T(x + delta) (when delta != NULL); h = T(weight * T(x_out * rsqrt(mean(x_out^2) + eps)));
 * h_out (T, nullable) receives h, q_out (int8, nullable) receives sat(rint(h)) — the int8 tensor the
 * per-tensor W8A8BFP32OFP32Linear would derive from h (linear.py:95).  H % 8 == 0, H <= 8192. */
int asq_add_rmsnorm_quant(const void* x, const void* delta, const void* weight, void* x_out, void* h_out,
                          int8_t* q_out, int dtype, int64_t M, int64_t H, float eps, void* stream);

/* gate_up rows = [gate (I) | up (I)], row_stride elements apart.  a = T(T(silu(gate)) * up);
 * a_out (T, nullable) receives a, q_out (int8, nullable) receives sat(rint(T(a / quant_scale))) — what
 * W8A8BFP32OFP32LinearWithQuantScale derives from a (linear.py:290-292). */
int asq_silu_mul_quant(const void* gate_up, int dtype, int64_t M, int64_t I, int64_t row_stride,
                       float quant_scale, int div_mode, int8_t* q_out, void* a_out, void* stream);

/* HF rotate-half RoPE in place on the first n_heads*head_dim elements of each of M rows (row_stride
 * elements apart); position = row % S; cos/sin tables [S, head_dim] of T. */
int asq_rope_inplace(void* qk, int dtype, const void* cos_table, const void* sin_table, int64_t M, int64_t S,
                     int64_t row_stride, int64_t n_heads, int64_t head_dim, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ASQ_H_ */
