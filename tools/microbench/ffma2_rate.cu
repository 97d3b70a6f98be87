// Packed FP32 FMA on B200: issue rate of FFMA2 (PTX fma.rn.f32x2; two IEEE fp32 FMAs per lane and
// instruction, bit-identical to two fmaf) next to scalar FFMA -- alone, and in the instruction mix of
// a register-tiled GEMM inner loop fed from shared memory (2 x LDS.128 per 16 MAC per lane).
// Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void fma2(uint64_t &c, uint64_t a, uint64_t b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}
__device__ __forceinline__ float2 upk2(uint64_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

// KIND 0: 32 scalar FFMA chains; 1: 16 FFMA2 chains (both operands packed); 2: 16 FFMA2 chains with a
// scalar broadcast operand (SASS: Rb.F32)
template <int KIND>
__global__ void __launch_bounds__(512) alu_kernel(float *out, int iters, float x, float y) {
  float c[32];
  for (int i = 0; i < 32; ++i) c[i] = threadIdx.x * 1e-3f + i;
  uint64_t c2[16];
  for (int i = 0; i < 16; ++i) c2[i] = pk2(c[2 * i], c[2 * i + 1]);
  const uint64_t x2 = pk2(x, x + 1e-3f), y2 = pk2(y, y);
  for (int it = 0; it < iters; ++it) {
    if (KIND == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) c[i] = fmaf(c[i], x, y);
    } else if (KIND == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c2[i]) : "l"(x2), "l"(y2));
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c2[i]) : "l"(pk2(x, x)), "l"(y2));
    }
  }
  float s = 0.f;
  for (int i = 0; i < 32; ++i) s += c[i];
  for (int i = 0; i < 16; ++i) { float2 v = upk2(c2[i]); s += v.x + v.y; }
  if (s == 123.456f) out[0] = s;
}

// 4 x 4 register tile per lane, operands from shared memory: per k two LDS.128, then 16 FFMA or 8 FFMA2
template <int KIND>
__global__ void __launch_bounds__(512) gemm_kernel(float *out, int iters) {
  __shared__ __align__(16) float A[64 * 36], W[64 * 68];
  for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) A[i] = 1e-3f * i;
  for (int i = threadIdx.x; i < 64 * 68; i += blockDim.x) W[i] = 1e-4f * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float *ap = A + (lane >> 4) * 4, *wp = W + (lane & 15) * 4;
  float acc[4][4] = {};
  uint64_t acc2[2][4] = {};
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int k = 0; k < 64; ++k) {
      const float4 a = *reinterpret_cast<const float4 *>(ap + k * 36);
      const float4 w = *reinterpret_cast<const float4 *>(wp + k * 68);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
      if (KIND == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[i][u] = fmaf(av[i], wv[u], acc[i][u]);
      } else {
        const uint64_t a01 = pk2(a.x, a.y), a23 = pk2(a.z, a.w);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint64_t ww = pk2(wv[u], wv[u]);
          fma2(acc2[0][u], a01, ww);
          fma2(acc2[1][u], a23, ww);
        }
      }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 4; ++i) for (int u = 0; u < 4; ++u) s += acc[i][u];
  for (int i = 0; i < 2; ++i) for (int u = 0; u < 4; ++u) { float2 v = upk2(acc2[i][u]); s += v.x + v.y; }
  if (s == 123.456f) out[0] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  float *out;
  cudaMalloc(&out, 4);
  const int grid = p.multiProcessorCount;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char *names[5] = {"ffma (32 chains)", "ffma2 packed x packed", "ffma2 packed x scalar", "tile 4x4: 2 LDS.128 + 16 FFMA",
                          "tile 4x4: 2 LDS.128 + 8 FFMA2"};
  for (int threads = 512; threads >= 128; threads /= 2) {
    printf("-- %d threads per CTA, one CTA per SM\n", threads);
    for (int kind = 0; kind < 5; ++kind) {
      float best = 1e30f;
      const int iters = kind < 3 ? 20000 : 2000;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (kind == 0) alu_kernel<0><<<grid, threads>>>(out, iters, 0.999f, 0.001f);
        else if (kind == 1) alu_kernel<1><<<grid, threads>>>(out, iters, 0.999f, 0.001f);
        else if (kind == 2) alu_kernel<2><<<grid, threads>>>(out, iters, 0.999f, 0.001f);
        else if (kind == 3) gemm_kernel<0><<<grid, threads>>>(out, iters);
        else gemm_kernel<1><<<grid, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      const double lanes = (double)grid * threads;
      const double mac = (kind < 3 ? 32.0 : 16.0 * 64) * iters * lanes;
      printf("%-34s %8.3f ms  %8.2f TFLOP/s  %8.1f MAC/clk/SM (at %d MHz nominal)\n", names[kind], best,
             2 * mac / (best * 1e-3) / 1e12, mac / (best * 1e-3) / (clk_khz * 1e3) / grid, clk_khz / 1000);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
