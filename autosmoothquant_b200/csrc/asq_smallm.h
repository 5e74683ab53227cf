// Decode-sized (M <= 16) int8 linear: see asq_smallm.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

struct SmallMParams {
  const int8_t* xq;        // [M,K] int8 activations (already quantised)
  const int8_t* w;         // [N,K] int8
  const float* row_scale;  // [M] per-token scales or nullptr
  const float* col_scale;  // [N] per-column dequant scales or nullptr
  const float* bias;       // [N] fp32 or nullptr
  void* y;                 // [M,N]
  float dequant_scale;
  int M, N, K, y_dtype;
};

bool asq_smallm_supported(int64_t M, int64_t N, int64_t K, int y_dtype);
int asq_smallm_launch(const SmallMParams& p, cudaStream_t stream);
