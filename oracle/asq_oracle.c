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
