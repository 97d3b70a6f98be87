// Device helpers shared by the training kernels (fit.cu: FP32 FFMA; fit_mma.cu: 3xTF32 tensor pipe).
#pragma once
#include "common.cuh"

namespace {

__device__ __forceinline__ float f_act(int a, float v) {
  switch (a) {
    case BORE_ACT_RELU: return fmaxf(v, 0.f);
    case BORE_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case BORE_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case BORE_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
__device__ __forceinline__ float f_act_bwd(int a, float h) {
  switch (a) {
    case BORE_ACT_RELU: return h > 0.f ? 1.f : 0.f;
    case BORE_ACT_ELU: return h > 0.f ? 1.f : h + 1.f;
    case BORE_ACT_SIGMOID: return h * (1.f - h);
    case BORE_ACT_TANH: return 1.f - h * h;
    default: return 1.f;
  }
}
__device__ __forceinline__ float stable_sigmoid(float u) {
  if (u >= 0.f) return 1.f / (1.f + expf(-u));
  const float e = expf(u);
  return e / (1.f + e);
}

__device__ __forceinline__ float block_sum(float v, float *red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// Keras-form Adam on one parameter (eps outside the bias correction); returns the new value.
// sqrt.approx / div.approx (<= 2 ulp each): the IEEE forms cost ~25 instructions per parameter,
// a fifth of the whole kernel at Dense32 sizes, for digits far below fp32 re-association noise.
__device__ __forceinline__ float adam_update(float wv, float g, float &m, float &v, float om1, float om2,
                                             float alpha, float eps) {
  m += (g - m) * om1;
  v += (g * g - v) * om2;
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return wv - __fdividef(m * alpha, r + eps);
}

// e / n for 0 <= e < 2^20 through a precomputed reciprocal (inv = 1.f / n): exact, the offset .5
// keeps the product away from integer boundaries
__device__ __forceinline__ int fdiv(int e, float inv) { return __float2int_rz(((float)e + 0.5f) * inv); }

}  // namespace
