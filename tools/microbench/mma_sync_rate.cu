// Legacy tensor path on B200: issue rate of mma.sync (HMMA) for TF32 m16n8k8, BF16 m16n8k16 and FP16 m16n8k16,
// register operands only, 8 independent accumulator chains per warp, all SMs.  Prints TFLOP/s (2 x MAC) and
// MAC / clk / SM next to an FFMA loop of the same shape.  Standalone: nvcc -arch=sm_100a -O3 -o mma_sync_rate ...
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int KIND>
__global__ void __launch_bounds__(512) rate_kernel(float *out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x * 0x3f800000u, a1 = 0x3f800000u, a2 = 0x3f000000u, a3 = 0x3e800000u;
  uint32_t b0 = 0x3f800000u + blockIdx.x, b1 = 0x3f400000u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(512) ffma_kernel(float *out, int iters, float x, float y) {
  float c[32];
  for (int i = 0; i < 32; ++i) c[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = fmaf(c[i], x, y);
  }
  float s = 0.f;
  for (int i = 0; i < 32; ++i) s += c[i];
  if (s == 123.456f) out[0] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  float *out;
  cudaMalloc(&out, 4);
  const int grid = p.multiProcessorCount, threads = 512, iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char *names[4] = {"mma.sync m16n8k8 tf32", "mma.sync m16n8k16 bf16", "mma.sync m16n8k16 f16", "ffma"};
  for (int kind = 0; kind < 4; ++kind) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      if (kind == 0) rate_kernel<0><<<grid, threads>>>(out, iters);
      else if (kind == 1) rate_kernel<1><<<grid, threads>>>(out, iters);
      else if (kind == 2) rate_kernel<2><<<grid, threads>>>(out, iters);
      else ffma_kernel<<<grid, threads>>>(out, iters, 0.999f, 0.001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double warps = (double)grid * threads / 32;
    double mac = kind == 0 ? 16.0 * 8 * 8 : (kind < 3 ? 16.0 * 8 * 16 : 32.0);
    const double per_iter = kind < 3 ? 8 : 32;
    const double total = mac * per_iter * iters * warps;
    printf("%-24s %8.3f ms  %8.2f TFLOP/s  %8.1f MAC/clk/SM (at %d MHz nominal)\n", names[kind], best,
           2 * total / (best * 1e-3) / 1e12, total / (best * 1e-3) / (clk_khz * 1e3) / grid, clk_khz / 1000);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
