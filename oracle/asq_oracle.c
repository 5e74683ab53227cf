/*
 * asq_oracle.c — plain-C restatement of the integer core of the W8A8 path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/w8a8_oracle.py): used by tests/ to cross-check the numpy
 * oracle with an independent implementation and by bench.py's CPU-baseline leg.  Never linked
 * into or called from the product library.
 *
 * Restates (paths relative to the reference repo AniZpZ/AutoSmoothQuant):
 *   - csrc/int8gemm/bindings.cpp:69-84 + cublasINT8MMWrapper.cc:224-354:
 *       C[m,n] (int32) = sum_k A[m,k] (int8) * W[n,k] (int8), alpha = 1, beta = 0
 *   - autosmoothquant/layers/nn/linear.py:88-92 (per-token fp32 path):
 *       s = absmax_row / 127 ; q = clamp(rint(x / s), -128, 127)      (x fp32; NaN -> 0)
 *   - autosmoothquant/layers/nn/linear.py:104 (fp32 dequant): y = (ds * s[m]) * (float)acc + bias
 *   - autosmoothquant/layers/functional/quantization.py:173-191 (per_token_quantize_fp8, fp32 path):
 *       s = absmax_row / 448 ; q = e4m3(clamp(x / s, -448, 448))   with torch's float8_e4m3fn cast restated on the bits
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/libasq_oracle.so oracle/asq_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>

void oracle_i8gemm_o32(const int8_t* a, const int8_t* w, int32_t* c, int64_t M, int64_t N, int64_t K) {
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < M; ++m) {
    const int8_t* ar = a + m * K;
    for (int64_t n = 0; n < N; ++n) {
      const int8_t* wr = w + n * K;
      int32_t acc = 0;
      for (int64_t k = 0; k < K; ++k) acc += (int32_t)ar[k] * (int32_t)wr[k];
      c[m * N + n] = acc;
    }
  }
}

static int8_t sat_i8(float v) {
  if (isnan(v)) return 0;
  float r = nearbyintf(v); /* round-half-to-even in the default rounding mode */
  if (r < -128.f) r = -128.f;
  if (r > 127.f) r = 127.f;
  return (int8_t)r;
}

/* fp32 activations only: the 16-bit dtype roundings live in the numpy oracle. */
void oracle_quant_per_token_f32(const float* x, int8_t* q, float* scale, int64_t M, int64_t K) {
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < M; ++m) {
    const float* xr = x + m * K;
    float amax = 0.f;
    for (int64_t k = 0; k < K; ++k) amax = fmaxf(amax, fabsf(xr[k]));
    const float s = amax / 127.0f;
    scale[m] = s;
    for (int64_t k = 0; k < K; ++k) q[m * K + k] = sat_i8(xr[k] / s);
  }
}

void oracle_quant_round_f32(const float* x, int8_t* q, int64_t count) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < count; ++i) q[i] = sat_i8(x[i]);
}

void oracle_dequant_f32(const int32_t* acc, const float* row_scale /* nullable */, float ds, const float* bias /* nullable */,
                        float* y, int64_t M, int64_t N) {
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < M; ++m) {
    const float f = row_scale ? ds * row_scale[m] : ds;
    for (int64_t n = 0; n < N; ++n) {
      volatile float t = f * (float)acc[m * N + n]; /* volatile: keep the multiply and the add separately rounded */
      y[m * N + n] = bias ? t + bias[n] : t;
    }
  }
}

/* ------------------------------------------------------------------------------------------------ float8_e4m3fn
 * 1 sign, 4 exponent (bias 7), 3 mantissa bits; no infinities; S.1111.111 is NaN; largest finite 448 = 0x7E;
 * subnormals m * 2^-9.  Encoder works on the fp32 bit pattern (round to nearest, ties to even), independently of the
 * log2 / rint formulation of oracle/w8a8_oracle.py: the two are cross-checked in tests/test_c_oracle.py.
 * Inputs above 448 in magnitude are NaN for torch's cast (saturation is done by the clamp BEFORE the cast,
 * quantization.py:190); here they map to NaN as well. */
#include <string.h>

static uint8_t e4m3_from_f32(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  const uint8_t sign = (uint8_t)((u >> 24) & 0x80u);
  const uint32_t a = u & 0x7FFFFFFFu;
  if (a > 0x7F800000u) return (uint8_t)(sign | 0x7Fu);         /* NaN */
  if (a > 0x43E80000u) return (uint8_t)(sign | 0x7Fu);         /* > 464: beyond the rounding range of 448 -> NaN */
  const int32_t exp = (int32_t)(a >> 23) - 127;                /* unbiased fp32 exponent */
  if (exp < -10) {
    /* below 2^-10 = half of the smallest subnormal quantum: rounds to zero, except exactly-half ties (-> even = 0) */
    return sign;
  }
  uint32_t mant = (a & 0x7FFFFFu) | 0x800000u;                 /* 24-bit significand, value = mant * 2^(exp-23) */
  int shift;                                                   /* bits dropped to reach the e4m3 quantum */
  int32_t e_field;
  if (exp >= -6) { shift = 20; e_field = exp + 7; }            /* normal: 3 mantissa bits below the leading one */
  else { shift = 20 + (-6 - exp); e_field = 0; }               /* subnormal: fixed quantum 2^-9 */
  const uint32_t kept = mant >> shift;
  const uint32_t rem = mant & ((1u << shift) - 1u);
  const uint32_t half = 1u << (shift - 1);
  uint32_t q = kept + ((rem > half || (rem == half && (kept & 1u))) ? 1u : 0u);
  uint32_t code;
  if (e_field == 0) {
    code = q;                                                  /* 0..8: 8 carries into the first normal (0x08) */
  } else {
    if (q == 16u) { q = 8u; e_field += 1; }                    /* mantissa overflow: next binade */
    code = ((uint32_t)e_field << 3) | (q - 8u);
  }
  if (code > 0x7Eu) return (uint8_t)(sign | 0x7Fu);            /* (448, 464] rounds to 480: not representable -> NaN */
  return (uint8_t)(sign | code);
}

static float e4m3_to_f32(uint8_t b) {
  const int ex = (b >> 3) & 0xF, m = b & 7;
  float v;
  if ((b & 0x7F) == 0x7F) return NAN;
  v = ex == 0 ? ldexpf((float)m, -9) : ldexpf((float)(8 + m), ex - 10);
  return (b & 0x80) ? -v : v;
}

void oracle_e4m3_encode(const float* x, uint8_t* q, int64_t count) {
  for (int64_t i = 0; i < count; ++i) q[i] = e4m3_from_f32(x[i]);
}

void oracle_e4m3_decode(const uint8_t* q, float* x, int64_t count) {
  for (int64_t i = 0; i < count; ++i) x[i] = e4m3_to_f32(q[i]);
}

/* per_token_quantize_fp8 on fp32 activations: scale[m] = absmax / 448, q = e4m3(clamp(x / scale)) */
void oracle_fp8_quant_per_token_f32(const float* x, uint8_t* q, float* scale, int64_t M, int64_t K) {
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < M; ++m) {
    const float* xr = x + m * K;
    float amax = 0.f;
    for (int64_t k = 0; k < K; ++k) amax = fmaxf(amax, fabsf(xr[k]));
    const float s = amax / 448.0f;
    scale[m] = s;
    for (int64_t k = 0; k < K; ++k) {
      float v = xr[k] / s;                                     /* 0/0 -> NaN stays NaN through the clamp */
      if (v > 448.f) v = 448.f;
      if (v < -448.f) v = -448.f;
      q[m * K + k] = e4m3_from_f32(v);
    }
  }
}

/* y[m,n] = (sum_k e4m3(a)[m,k] * e4m3(w)[n,k]) * (a_scale[m] * w_scale) (+ bias[n]) evaluated in double */
void oracle_fp8_linear_f64(const uint8_t* a, const uint8_t* w, const float* a_scale, float w_scale, const float* bias /* nullable */,
                           double* y, int64_t M, int64_t N, int64_t K) {
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < M; ++m)
    for (int64_t n = 0; n < N; ++n) {
      double acc = 0.0;
      for (int64_t k = 0; k < K; ++k) acc += (double)e4m3_to_f32(a[m * K + k]) * (double)e4m3_to_f32(w[n * K + k]);
      y[m * N + n] = acc * ((double)a_scale[m] * (double)w_scale) + (bias ? (double)bias[n] : 0.0);
    }
}
