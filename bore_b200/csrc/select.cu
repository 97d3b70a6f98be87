// K4: selection kernels around the argmax (sm_100a).
//
//  * bore_topk_smallest  replaces np.argpartition(f_init, kth=num_starts-1) at
//    bore/mixins.py:56 -- which of the screening samples become L-BFGS-B starts.
//  * bore_select_best    replaces the scan of bore/mixins.py:80-87 -- the FIRST minimum of
//    `fun` over results with (success or status == 1) that pass the filter -- and packs the
//    winner into one int64 key so that ranks can agree with a single NCCL max all-reduce
//    (NCCL has no MAXLOC).
// Both are latency-bound integer work on a few thousand keys; one CTA (top-k up to 4096
// keys) or a plain grid of compare-exchange steps is all they need.
#include <algorithm>

#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned orderable(float v) {
  if (v == 0.f) v = 0.f;  // -0 and +0 compare equal on the host
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// key = (orderable(f) << 32) | index : ascending key order == ascending f, ties by lower index
__global__ void make_keys_kernel(const float *__restrict__ f, int S, int n_pow2, float sign,
                                 unsigned long long *__restrict__ keys) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pow2; i += gridDim.x * blockDim.x) {
    unsigned long long k = ~0ULL;  // padding sorts last
    if (i < S) {
      const float v = sign * f[i];
      const unsigned o = (v != v) ? 0xffffffffu : orderable(v);  // NaN last
      k = ((unsigned long long)o << 32) | (unsigned)i;
    }
    keys[i] = k;
  }
}

__global__ void bitonic_step_kernel(unsigned long long *keys, int n, int j, int kk) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ixj = i ^ j;
    if (ixj > i) {
      const unsigned long long a = keys[i], b = keys[ixj];
      const bool up = (i & kk) == 0;
      if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
    }
  }
}

// whole sort in shared memory for n <= 4096
__global__ void __launch_bounds__(1024) bitonic_smem_kernel(unsigned long long *keys, int n) {
  extern __shared__ unsigned long long sk[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) sk[i] = keys[i];
  __syncthreads();
  for (int kk = 2; kk <= n; kk <<= 1)
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = sk[i], b = sk[ixj];
          const bool up = (i & kk) == 0;
          if ((a > b) == up) { sk[i] = b; sk[ixj] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += blockDim.x) keys[i] = sk[i];
}

__global__ void take_indices_kernel(const unsigned long long *__restrict__ keys, int k,
                                    int32_t *__restrict__ idx) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x)
    idx[i] = (int32_t)(keys[i] & 0xffffffffu);
}

__global__ void iota_kernel(int32_t *idx, int k) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) idx[i] = i;
}

__global__ void __launch_bounds__(256)
select_best_kernel(const double *__restrict__ fun, const int32_t *__restrict__ status,
                   const uint8_t *__restrict__ keep, int S, long long idx_offset,
                   unsigned long long *key_out) {
  unsigned long long best = 0ULL;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) {
    const int st = status[i];
    if (!(st == 0 || st == 1)) continue;  // success or maxiter/maxfun (bore/mixins.py:83-85)
    if (keep && !keep[i]) continue;
    const float v = (float)fun[i];
    if (v != v) continue;
    const unsigned long long o = orderable(-v);
    const unsigned long long key = (o << 31) | (unsigned long long)(0x7fffffffLL - (i + idx_offset));
    best = key > best ? key : best;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(key_out, best);
}

// ---- grouped variants (batched problems: one group of samples / starts per model) ----
// k smallest of each group of P values: one CTA per group, keys sorted in shared memory
__global__ void __launch_bounds__(1024)
topk_groups_kernel(const float *__restrict__ f, int P, int n_pow2, int k, float sign,
                   int32_t *__restrict__ idx) {
  extern __shared__ unsigned long long sk[];
  const float *fg = f + (size_t)blockIdx.x * P;
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
    unsigned long long key = ~0ULL;
    if (i < P) {
      const float v = sign * fg[i];
      const unsigned o = (v != v) ? 0xffffffffu : orderable(v);
      key = ((unsigned long long)o << 32) | (unsigned)i;
    }
    sk[i] = key;
  }
  __syncthreads();
  for (int kk = 2; kk <= n_pow2; kk <<= 1)
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = sk[i], b = sk[ixj];
          const bool up = (i & kk) == 0;
          if ((a > b) == up) { sk[i] = b; sk[ixj] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < k; i += blockDim.x)
    idx[(size_t)blockIdx.x * k + i] = (int32_t)(sk[i] & 0xffffffffu);
}

// first minimum of every group of `per_group` results: one warp per group
__global__ void __launch_bounds__(128)
select_best_groups_kernel(const double *__restrict__ fun, const int32_t *__restrict__ status,
                          const uint8_t *__restrict__ keep, int n_groups, int per_group,
                          unsigned long long *__restrict__ keys) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= n_groups) return;
  unsigned long long best = 0ULL;
  for (int i = lane; i < per_group; i += 32) {
    const size_t q = (size_t)g * per_group + i;
    const int st = status[q];
    if (!(st == 0 || st == 1)) continue;
    if (keep && !keep[q]) continue;
    const float v = (float)fun[q];
    if (v != v) continue;
    const unsigned long long key = ((unsigned long long)orderable(-v) << 31) |
                                   (unsigned long long)(0x7fffffffLL - i);
    best = key > best ? key : best;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if (lane == 0) keys[g] = best;
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

extern "C" {

size_t bore_topk_workspace_bytes(int S, int k) {
  (void)k;
  if (S < 1) return 0;
  return (size_t)next_pow2(S) * sizeof(unsigned long long);
}

int bore_topk_smallest(const float *f_dev, int S, int k, int negate, int32_t *idx_dev,
                       void *work_dev, size_t work_bytes, int device, void *stream_) {
  BORE_CHECK(S >= 1 && k >= 0 && k <= S, "bore_topk_smallest: k=%d, S=%d", k, S);
  BORE_CHECK(bore_device_count() > 0, "no CUDA device visible -- bore_b200 has no CPU fallback");
  if (k == 0) return 0;
  BORE_CHECK(f_dev && idx_dev, "bore_topk_smallest: NULL buffer");
  BORE_CUDA(cudaSetDevice(device));
  cudaStream_t stream = (cudaStream_t)stream_;
  if (k == S) {  // every sample is a start: argpartition degenerates to a permutation
    iota_kernel<<<(k + 255) / 256, 256, 0, stream>>>(idx_dev, k);
    BORE_CUDA(cudaGetLastError());
    return 0;
  }
  const int n = next_pow2(S);
  BORE_CHECK(work_dev && work_bytes >= (size_t)n * sizeof(unsigned long long),
             "bore_topk_smallest: workspace too small");
  unsigned long long *keys = static_cast<unsigned long long *>(work_dev);
  const int blocks = std::min((n + 255) / 256, 2048);
  make_keys_kernel<<<blocks, 256, 0, stream>>>(f_dev, S, n, negate ? -1.f : 1.f, keys);
  if (n <= 4096) {
    const size_t smem = (size_t)n * sizeof(unsigned long long);
    bitonic_smem_kernel<<<1, std::min(1024, std::max(32, n / 2)), smem, stream>>>(keys, n);
  } else {
    for (int kk = 2; kk <= n; kk <<= 1)
      for (int j = kk >> 1; j > 0; j >>= 1) bitonic_step_kernel<<<blocks, 256, 0, stream>>>(keys, n, j, kk);
  }
  take_indices_kernel<<<(k + 255) / 256, 256, 0, stream>>>(keys, k, idx_dev);
  BORE_CUDA(cudaGetLastError());
  return 0;
}

int bore_select_best(const double *fun_dev, const int32_t *status_dev, const uint8_t *keep_dev, int S,
                     int64_t idx_offset, int64_t *key_dev, int device, void *stream_) {
  BORE_CHECK(S >= 0 && key_dev, "bore_select_best: bad arguments");
  BORE_CHECK(idx_offset >= 0 && idx_offset + S <= 0x7fffffffLL, "bore_select_best: index range");
  BORE_CHECK(bore_device_count() > 0, "no CUDA device visible -- bore_b200 has no CPU fallback");
  BORE_CUDA(cudaSetDevice(device));
  cudaStream_t stream = (cudaStream_t)stream_;
  BORE_CUDA(cudaMemsetAsync(key_dev, 0, sizeof(int64_t), stream));
  if (S > 0) {
    BORE_CHECK(fun_dev && status_dev, "bore_select_best: NULL buffer");
    select_best_kernel<<<std::min((S + 255) / 256, 1024), 256, 0, stream>>>(
        fun_dev, status_dev, keep_dev, S, (long long)idx_offset,
        reinterpret_cast<unsigned long long *>(key_dev));
    BORE_CUDA(cudaGetLastError());
  }
  return 0;
}

int bore_topk_smallest_groups(const float *f_dev, int n_groups, int per_group, int k, int negate,
                              int32_t *idx_dev, int device, void *stream_) {
  BORE_CHECK(n_groups >= 1 && per_group >= 1 && k >= 1 && k <= per_group,
             "bore_topk_smallest_groups: n_groups=%d per_group=%d k=%d", n_groups, per_group, k);
  BORE_CHECK(per_group <= 4096, "bore_topk_smallest_groups: at most 4096 samples per group");
  BORE_CHECK(f_dev && idx_dev, "bore_topk_smallest_groups: NULL buffer");
  BORE_CHECK(bore_device_count() > 0, "no CUDA device visible -- bore_b200 has no CPU fallback");
  BORE_CUDA(cudaSetDevice(device));
  const int n = next_pow2(per_group);
  topk_groups_kernel<<<n_groups, std::min(1024, std::max(32, n / 2)), (size_t)n * sizeof(unsigned long long),
                       (cudaStream_t)stream_>>>(f_dev, per_group, n, k, negate ? -1.f : 1.f, idx_dev);
  BORE_CUDA(cudaGetLastError());
  return 0;
}

int bore_select_best_groups(const double *fun_dev, const int32_t *status_dev,
                            const uint8_t *keep_dev, int n_groups, int per_group, int64_t *keys_dev,
                            int device, void *stream_) {
  BORE_CHECK(n_groups >= 1 && per_group >= 1 && fun_dev && status_dev && keys_dev,
             "bore_select_best_groups: bad arguments");
  BORE_CHECK(bore_device_count() > 0, "no CUDA device visible -- bore_b200 has no CPU fallback");
  BORE_CUDA(cudaSetDevice(device));
  select_best_groups_kernel<<<(n_groups * 32 + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(
      fun_dev, status_dev, keep_dev, n_groups, per_group,
      reinterpret_cast<unsigned long long *>(keys_dev));
  BORE_CUDA(cudaGetLastError());
  return 0;
}

// C1: the single collective of the multi-GPU argmax.  NCCL has ncclMax but no
// MAXLOC, so value and location travel in one int64 key (bore_select_best); this is
// ncclAllReduce(key, key, 1, ncclInt64, ncclMax, comm, stream) for hosts that do not go through
// torch.distributed.  NCCL is resolved at call time from the process (whatever libnccl.so.2 the
// host already loaded, else the system one), so libbore_b200.so has no link-time dependency on it.
#include <dlfcn.h>
int bore_allreduce_maxloc(void *nccl_comm, int64_t *key_dev, void *stream) {
  BORE_CHECK(nccl_comm != nullptr && key_dev != nullptr, "bore_allreduce_maxloc: NULL argument");
  typedef int (*allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
  typedef const char *(*errstr_fn)(int);
  static allreduce_fn fn = nullptr;
  static errstr_fn es = nullptr;
  if (!fn) {
    void *hnd = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!hnd) hnd = dlopen("libnccl.so.2", RTLD_NOW);
    if (!hnd) hnd = dlopen("libnccl.so", RTLD_NOW);
    BORE_CHECK(hnd != nullptr, "bore_allreduce_maxloc: cannot load libnccl.so.2 (%s)", dlerror());
    fn = reinterpret_cast<allreduce_fn>(dlsym(hnd, "ncclAllReduce"));
    es = reinterpret_cast<errstr_fn>(dlsym(hnd, "ncclGetErrorString"));
    BORE_CHECK(fn != nullptr, "bore_allreduce_maxloc: ncclAllReduce not found");
  }
  // nccl.h: ncclInt64 = 4, ncclMax = 2 (stable since NCCL 2.0)
  const int rc = fn(key_dev, key_dev, 1, 4, 2, nccl_comm, (cudaStream_t)stream);
  BORE_CHECK(rc == 0, "bore_allreduce_maxloc: ncclAllReduce -> %s", es ? es(rc) : "error");
  return 0;
}

}  // extern "C"
