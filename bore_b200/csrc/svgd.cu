// Batch argmax by Stein variational gradient descent, sm_100a.
//
// Replaces BatchMaximizableMixin.argmax_batch (bore/mixins.py:100-116) = SVGD.optimize /
// optimize_from_init (bore/optimizers/svgd/base.py:67-131) with the RBF kernel and median
// heuristic of bore/optimizers/svgd/kernels.py:4-28, the rank distortion of
// svgd/base.py:36-64 and the AdaGrad-with-decay step of svgd/base.py:103-113, for the case the
// reference ships: `func` is the model's own value-and-gradient closure (`self._func_max`,
// bore/mixins.py:98).  One iteration is two launches on one stream, no host round trip in the
// n_iter loop:
//   mlp_eval_kernel<grad>  f, df/dx of the n particles of every problem (fp32, K2)
//   svgd_step_kernel       one CTA per problem, everything else of the iteration in fp64 out of
//                          shared memory: pairwise squared distances, their exact median (radix
//                          select on the orderable bit patterns of the upper triangle; the full
//                          n x n matrix is that triangle twice plus n zeros), K = exp(-gamma r2),
//                          the kernel-weighted gradient K (zeta * f') + tau dK, the step, the
//                          clip into the box, and the fp32 copy of x that the next K2 reads.
// Summation orders are the plain index orders (numpy's differ in places: pairwise summation
// over the feature axis, BLAS for K @ V), so results agree with the reference to rounding, not
// bit for bit -- the reference's own SVGD tests compare at 1e-10 (tests/test_optimizers.py:98-105).
#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace {

struct SvgdArgs {
  int n, D, it;
  int has_bounds, use_median, use_rank, fg_is_f64;
  double length_scale, ls_eps, step_size, alpha, eps, tau, lambd, zeta_c;
  double *x;          // [P][n][D]  in/out
  float *x32;         // [P][n][D]  out (input of the next evaluation), may be NULL
  const void *f;      // [P][n]     fp32 (the MLP kernel's) or fp64 (a caller's own objective)
  const void *g;      // [P][n][D]
  double *hist;       // [P][n][D]  AdaGrad accumulator
  const double *lo, *hi;  // [D]
};

__device__ __forceinline__ double load_fg(const void *p, size_t i, int is_f64) {
  return is_f64 ? static_cast<const double *>(p)[i] : (double)static_cast<const float *>(p)[i];
}

constexpr int SVGD_THREADS = 256;

// t-th smallest (0-based) of keys[0..M): 8 passes of an 8-bit radix select, most significant
// byte first; s_hist (256 ints) and s_pick (2 ints) are shared scratch.  All threads return it.
__device__ unsigned long long radix_select(const unsigned long long *keys, int M, int t, int *s_hist,
                                           int *s_pick) {
  unsigned long long prefix = 0ULL, mask = 0ULL;
  const int tid = threadIdx.x;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int b = tid; b < 256; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    for (int i = tid; i < M; i += blockDim.x) {
      const unsigned long long k = keys[i];
      if ((k & mask) == prefix) atomicAdd(&s_hist[(int)((k >> shift) & 255ULL)], 1);
    }
    __syncthreads();
    if (tid < 32) {
      // lane l owns bins 8l..8l+7; exclusive prefix over lanes by shuffle
      int c[8], sum = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { c[q] = s_hist[tid * 8 + q]; sum += c[q]; }
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += up;
      }
      int before = incl - sum;
      if (t >= before && t < incl) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (t < before + c[q]) { s_pick[0] = tid * 8 + q; s_pick[1] = t - before; break; }
          before += c[q];
        }
      }
    }
    __syncthreads();
    prefix |= (unsigned long long)s_pick[0] << shift;
    mask |= 255ULL << shift;
    t = s_pick[1];
    __syncthreads();
  }
  return prefix;
}

// Median of the full n x n matrix of squared distances from the keys of its strict upper
// triangle (np.median over n*n entries: the middle one, or the mean of the two middle ones).
// Sorted, the full matrix is n zeros followed by every triangle value twice, so entry p is 0 for
// p < n and triangle order statistic (p - n) / 2 otherwise.  All threads return the value.
__device__ double full_matrix_median(const unsigned long long *keys, int n, int *s_hist, int *s_pick,
                                     unsigned long long *s_red, int *s_cnt) {
  const int tid = threadIdx.x, T = blockDim.x, ntri = n * (n - 1) / 2, n2 = n * n;
  if (n2 & 1) {
    const int p = (n2 - 1) / 2;
    if (p < n) return 0.0;
    return __longlong_as_double((long long)radix_select(keys, ntri, (p - n) >> 1, s_hist, s_pick));
  }
  const int p1 = n2 / 2 - 1, p2 = n2 / 2;
  if (p2 < n) return 0.0;
  const int t2 = (p2 - n) >> 1;
  const unsigned long long k2 = radix_select(keys, ntri, t2, s_hist, s_pick);
  const double hi_v = __longlong_as_double((long long)k2);
  double lo_v = 0.0;
  if (p1 >= n) {
    const int t1 = (p1 - n) >> 1;
    if (t1 == t2) {
      lo_v = hi_v;
    } else {
      // order statistic t2 - 1: the largest key below k2, unless k2 itself repeats downwards
      int cnt = 0;
      unsigned long long mx = 0ULL;
      for (int i = tid; i < ntri; i += T) {
        const unsigned long long k = keys[i];
        if (k < k2) { ++cnt; mx = k > mx ? k : mx; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        const unsigned long long om = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = om > mx ? om : mx;
      }
      if ((tid & 31) == 0) { s_cnt[tid >> 5] = cnt; s_red[tid >> 5] = mx; }
      __syncthreads();
      cnt = 0; mx = 0ULL;
      for (int w = 0; w < T / 32; ++w) { cnt += s_cnt[w]; mx = s_red[w] > mx ? s_red[w] : mx; }
      __syncthreads();
      lo_v = (cnt <= t1) ? hi_v : __longlong_as_double((long long)mx);
    }
  }
  return (lo_v + hi_v) / 2.0;
}

// gamma = .5 / length_scale^2 with the median heuristic when no length scale is given
// (kernels.py:4-10, 24): length_scale = sqrt(.5 * median / log(n + 1)), floored at eps = 1e-6
__device__ double rbf_gamma(double median_or_nan, double length_scale, int n, double ls_eps) {
  double ls = length_scale;
  if (median_or_nan == median_or_nan) ls = sqrt(0.5 * median_or_nan / log((double)n + 1.0));
  ls = fmax(ls, ls_eps);
  return 0.5 / (ls * ls);
}

__global__ void __launch_bounds__(SVGD_THREADS) svgd_step_kernel(const SvgdArgs a) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_hist[256];
  __shared__ int s_pick[2];
  __shared__ double s_gamma;
  __shared__ unsigned long long s_red[SVGD_THREADS / 32];
  __shared__ int s_cnt[SVGD_THREADS / 32];
  const int n = a.n, D = a.D, tid = threadIdx.x, T = blockDim.x;
  const int nD = n * D, ntri = n * (n - 1) / 2;
  double *xs = sm;                 // [n][D]
  double *Ks = xs + nD;            // [n][n]   r2, then K
  double *vs = Ks + n * n;         // [n][D]   zeta * f'   (the select keys live here before that)
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(vs);
  double *zeta = vs + (nD > ntri ? nD : ntri);  // [n]
  const size_t pb = (size_t)blockIdx.x;
  double *xg = a.x + pb * nD;
  const size_t f0 = pb * n, g0 = pb * nD;

  for (int e = tid; e < nD; e += T) xs[e] = xg[e];
  __syncthreads();

  // ---- pairwise squared distances (kernels.py:19-22); pair p <-> (i, j), i < j ----
  for (int i = tid; i < n; i += T) Ks[i * n + i] = 0.0;
  for (int p = tid; p < ntri; p += T) {
    // row i of the strict upper triangle starts at i*n - i*(i+1)/2
    int i = (int)((2.0 * n - 1.0 - sqrt((2.0 * n - 1.0) * (2.0 * n - 1.0) - 8.0 * p)) * 0.5);
    while (i > 0 && i * n - i * (i + 1) / 2 > p) --i;
    while ((i + 1) * n - (i + 1) * (i + 2) / 2 <= p) ++i;
    const int j = p - (i * n - i * (i + 1) / 2) + i + 1;
    double s = 0.0;
    for (int d = 0; d < D; ++d) {
      const double df = xs[i * D + d] - xs[j * D + d];
      s = __dadd_rn(s, __dmul_rn(df, df));
    }
    Ks[i * n + j] = s;
    Ks[j * n + i] = s;
    if (a.use_median) keys[p] = (unsigned long long)__double_as_longlong(s);  // s >= 0: bits are monotone
  }
  __syncthreads();

  // ---- length scale (kernels.py:4-10): median heuristic over the FULL n x n matrix ----
  {
    double med = __longlong_as_double(0x7ff8000000000000LL);
    if (a.use_median) med = full_matrix_median(keys, n, s_hist, s_pick, s_red, s_cnt);
    if (tid == 0) s_gamma = rbf_gamma(med, a.length_scale, n, a.ls_eps);
  }
  __syncthreads();
  const double gamma = s_gamma;

  // ---- K = exp(-gamma r2); zeta = distortion(rank(f)) (svgd/base.py:36-64, 97-98) ----
  for (int e = tid; e < n * n; e += T) Ks[e] = exp(-gamma * Ks[e]);
  for (int i = tid; i < n; i += T) {
    double z = a.zeta_c;
    if (a.use_rank) {
      const double fi = load_fg(a.f, f0 + i, a.fg_is_f64);
      int c = 0;
      for (int j = 0; j < n; ++j) c += load_fg(a.f, f0 + j, a.fg_is_f64) <= fi ? 1 : 0;
      z = pow((double)c / (double)n, -a.lambd);
    }
    zeta[i] = z;
  }
  __syncthreads();
  for (int e = tid; e < nD; e += T) vs[e] = zeta[e / D] * load_fg(a.g, g0 + e, a.fg_is_f64);
  __syncthreads();

  // ---- grad = (K @ (zeta f') + tau * K_grad) / n; AdaGrad step; clip (svgd/base.py:100-116) ----
  for (int e = tid; e < nD; e += T) {
    const int i = e / D, d = e - i * D;
    const double xi = xs[e];
    const double *Ki = Ks + i * n;
    double acc = 0.0, kg = 0.0;
    for (int j = 0; j < n; ++j) {
      const double k = Ki[j];
      acc = fma(k, vs[j * D + d], acc);
      kg += gamma * (xi - xs[j * D + d]) * k;
    }
    double grad = (acc + a.tau * (2.0 * kg)) / (double)n;
    double *hp = a.hist + pb * nD + e;
    double h;
    if (a.it == 0) h = grad * grad;
    else h = *hp * a.alpha + (1.0 - a.alpha) * (grad * grad);
    *hp = h;
    double xn = xi + a.step_size * (grad / (a.eps + sqrt(h)));
    if (a.has_bounds) xn = fmin(fmax(xn, a.lo[d]), a.hi[d]);
    xg[e] = xn;
    if (a.x32) a.x32[pb * nD + e] = (float)xn;
  }
}

__global__ void svgd_init_kernel(const double *__restrict__ x, float *__restrict__ x32, long total) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    x32[e] = (float)x[e];
}

size_t step_smem_bytes(int n, int D) {
  const size_t nD = (size_t)n * D, ntri = (size_t)n * (n - 1) / 2;
  return (nD + (size_t)n * n + std::max(nD, ntri) + n) * sizeof(double);
}

struct WorkLayout {
  size_t x32, f, g, hist, lo, hi, total;
};
WorkLayout work_layout(int P, int n, int D) {
  WorkLayout w;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  w.x32 = take((size_t)P * n * D * sizeof(float));
  w.f = take((size_t)P * n * sizeof(float));
  w.g = take((size_t)P * n * D * sizeof(float));
  w.hist = take((size_t)P * n * D * sizeof(double));
  w.lo = take((size_t)D * sizeof(double));
  w.hi = take((size_t)D * sizeof(double));
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

size_t bore_svgd_workspace_bytes(int n_problems, int n, int D) {
  if (n_problems < 1 || n < 1 || D < 1) return 0;
  return work_layout(n_problems, n, D).total;
}

static int svgd_launch_step(const SvgdArgs &a, int n_problems, int device, cudaStream_t stream) {
  const size_t smem = step_smem_bytes(a.n, a.D);
  BORE_CHECK(smem <= 200 * 1024, "svgd: %d particles x %d dims need %zu B of shared memory (max %d)",
             a.n, a.D, smem, 200 * 1024);
  static bool attr_done[64] = {};
  if (device >= 64 || !attr_done[device]) {
    BORE_CUDA(cudaFuncSetAttribute(svgd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (device < 64) attr_done[device] = true;
  }
  svgd_step_kernel<<<n_problems, SVGD_THREADS, smem, stream>>>(a);
  return 0;
}

static void svgd_fill_args(SvgdArgs &a, int n, int D, double length_scale, double step_size, double alpha,
                           double eps, double tau, double lambd, double zeta_c) {
  a.n = n; a.D = D; a.it = 0;
  a.use_median = !(length_scale == length_scale);  // NaN: length_scale=None (kernels.py:5-7)
  a.use_rank = lambd == lambd;                      // NaN: DistortionConstant (mixins.py:104-105)
  a.length_scale = length_scale; a.ls_eps = 1e-6;   // _check_length_scale's eps (kernels.py:4)
  a.step_size = step_size; a.alpha = alpha; a.eps = eps; a.tau = tau; a.lambd = lambd;
  a.zeta_c = zeta_c;
}

int bore_svgd_step(double *x_dev, int n_problems, int n, int D, const void *f_dev, const void *g_dev,
                   int fg_is_f64, const double *lo_dev, const double *hi_dev, double length_scale,
                   int iteration, double step_size, double alpha, double eps, double tau, double lambd,
                   double zeta_c, double *hist_dev, float *x32_dev, int device, void *stream_) {
  BORE_CHECK(bore_device_count() > 0, "no CUDA device visible -- bore_b200 has no CPU fallback");
  BORE_CHECK(n_problems >= 1 && n >= 1 && D >= 1 && iteration >= 0, "bore_svgd_step: bad sizes");
  BORE_CHECK(x_dev && f_dev && g_dev && hist_dev, "bore_svgd_step: NULL buffer");
  BORE_CHECK((lo_dev == nullptr) == (hi_dev == nullptr), "bore_svgd_step: give both bounds or none");
  BORE_CUDA(cudaSetDevice(device));
  SvgdArgs a;
  svgd_fill_args(a, n, D, length_scale, step_size, alpha, eps, tau, lambd, zeta_c);
  a.it = iteration;
  a.has_bounds = lo_dev != nullptr;
  a.fg_is_f64 = fg_is_f64;
  a.x = x_dev; a.x32 = x32_dev; a.f = f_dev; a.g = g_dev; a.hist = hist_dev; a.lo = lo_dev; a.hi = hi_dev;
  if (svgd_launch_step(a, n_problems, device, (cudaStream_t)stream_)) return -1;
  BORE_CUDA(cudaGetLastError());
  return 0;
}

int bore_svgd_maximize(bore_mlp *h, int model0, int n_problems, int transform, double *x_dev, int n,
                       const double *lo_host, const double *hi_host, double length_scale, int n_iter,
                       double step_size, double alpha, double eps, double tau, double lambd, double zeta_c,
                       void *work_dev, size_t work_bytes, void *stream_) {
  BORE_NVTX("bore:argmax_batch SVGD (K6)");
  BORE_CHECK(h != nullptr, "bore_svgd_maximize: NULL handle");
  BORE_CHECK(bore_device_count() > 0, "no CUDA device visible -- bore_b200 has no CPU fallback");
  BORE_CHECK(n_problems >= 1 && model0 >= 0 && model0 + n_problems <= h->n_models,
             "bore_svgd_maximize: models [%d, %d) out of range (%d)", model0, model0 + n_problems,
             h->n_models);
  BORE_CHECK(n >= 1 && n_iter >= 0 && x_dev, "bore_svgd_maximize: n=%d, n_iter=%d", n, n_iter);
  BORE_CHECK((lo_host == nullptr) == (hi_host == nullptr), "bore_svgd_maximize: give both bounds or none");
  const int D = h->desc.dims[0];
  const WorkLayout w = work_layout(n_problems, n, D);
  BORE_CHECK(work_dev && work_bytes >= w.total, "bore_svgd_maximize: workspace too small (%zu < %zu)",
             work_bytes, w.total);
  BORE_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  char *base = static_cast<char *>(work_dev);
  SvgdArgs a;
  svgd_fill_args(a, n, D, length_scale, step_size, alpha, eps, tau, lambd, zeta_c);
  a.has_bounds = lo_host != nullptr;
  a.fg_is_f64 = 0;
  a.x = x_dev;
  a.x32 = reinterpret_cast<float *>(base + w.x32);
  float *f = reinterpret_cast<float *>(base + w.f), *g = reinterpret_cast<float *>(base + w.g);
  a.f = f; a.g = g;
  a.hist = reinterpret_cast<double *>(base + w.hist);
  a.lo = reinterpret_cast<double *>(base + w.lo);
  a.hi = reinterpret_cast<double *>(base + w.hi);
  if (a.has_bounds) {
    BORE_CUDA(cudaMemcpyAsync(base + w.lo, lo_host, D * sizeof(double), cudaMemcpyHostToDevice, stream));
    BORE_CUDA(cudaMemcpyAsync(base + w.hi, hi_host, D * sizeof(double), cudaMemcpyHostToDevice, stream));
    BORE_CUDA(cudaStreamSynchronize(stream));  // the host arrays are the caller's (pageable) memory
  }
  if (n_iter == 0) return 0;
  const long total = (long)n_problems * n * D;
  svgd_init_kernel<<<(int)std::min<long>((total + 255) / 256, 1024), 256, 0, stream>>>(x_dev, a.x32, total);
  BORE_CUDA(cudaGetLastError());
  if (mlp_eval_prepare(h, model0, n_problems, true, stream)) return -1;
  for (int it = 0; it < n_iter; ++it) {
    // f, f' of transform(model(x)) for every particle: `self._func_max` (bore/mixins.py:98)
    if (launch_mlp_eval_multi(h, model0, n_problems, n, true, transform, 0, a.x32, f, g, nullptr, stream, 1))
      return -1;
    a.it = it;
    if (svgd_launch_step(a, n_problems, h->device, stream)) return -1;
  }
  BORE_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"

namespace {

// RadialBasis.value_and_grad (bore/optimizers/svgd/kernels.py:18-28) on its own: K [n][n] and
// K_grad [n][D] of one particle set -- the reference tests the kernel separately
// (tests/test_optimizers.py:76-127), and so do ours.
__global__ void __launch_bounds__(SVGD_THREADS)
rbf_value_and_grad_kernel(const double *__restrict__ x, int n, int D, double length_scale, int use_median,
                          double *__restrict__ K_out, double *__restrict__ Kg_out) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_hist[256];
  __shared__ int s_pick[2];
  __shared__ double s_gamma;
  __shared__ unsigned long long s_red[SVGD_THREADS / 32];
  __shared__ int s_cnt[SVGD_THREADS / 32];
  const int tid = threadIdx.x, T = blockDim.x;
  double *Ks = sm;
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(Ks + n * n);
  for (int e = tid; e < n * n; e += T) {
    const int i = e / n, j = e - i * n;
    double s = 0.0;
    for (int d = 0; d < D; ++d) {
      const double df = x[i * D + d] - x[j * D + d];
      s = __dadd_rn(s, __dmul_rn(df, df));
    }
    Ks[e] = s;
    if (use_median && i < j) keys[i * n - i * (i + 1) / 2 + (j - i - 1)] = (unsigned long long)__double_as_longlong(s);
  }
  __syncthreads();
  {
    double med = __longlong_as_double(0x7ff8000000000000LL);
    if (use_median) med = full_matrix_median(keys, n, s_hist, s_pick, s_red, s_cnt);
    if (tid == 0) s_gamma = rbf_gamma(med, length_scale, n, 1e-6);
  }
  __syncthreads();
  const double gamma = s_gamma;
  for (int e = tid; e < n * n; e += T) {
    const double k = exp(-gamma * Ks[e]);
    Ks[e] = k;
    K_out[e] = k;
  }
  __syncthreads();
  for (int e = tid; e < n * D; e += T) {
    const int i = e / D, d = e - i * D;
    double kg = 0.0;
    for (int j = 0; j < n; ++j) kg += gamma * (x[e] - x[j * D + d]) * Ks[i * n + j];
    Kg_out[e] = 2.0 * kg;
  }
}

}  // namespace

extern "C" int bore_svgd_kernel_value_and_grad(const double *x_dev, int n, int D, double length_scale,
                                               double *K_dev, double *Kgrad_dev, int device, void *stream_) {
  BORE_CHECK(bore_device_count() > 0, "no CUDA device visible -- bore_b200 has no CPU fallback");
  BORE_CHECK(n >= 1 && D >= 1 && x_dev && K_dev && Kgrad_dev, "bore_svgd_kernel_value_and_grad: bad arguments");
  const size_t smem = ((size_t)n * n + (size_t)n * (n - 1) / 2 + 2) * sizeof(double);
  BORE_CHECK(smem <= 200 * 1024, "bore_svgd_kernel_value_and_grad: n=%d too large", n);
  BORE_CUDA(cudaSetDevice(device));
  static bool attr_done[64] = {};
  if (device >= 64 || !attr_done[device]) {
    BORE_CUDA(cudaFuncSetAttribute(rbf_value_and_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024));
    if (device < 64) attr_done[device] = true;
  }
  rbf_value_and_grad_kernel<<<1, SVGD_THREADS, smem, (cudaStream_t)stream_>>>(
      x_dev, n, D, length_scale, !(length_scale == length_scale), K_dev, Kgrad_dev);
  BORE_CUDA(cudaGetLastError());
  return 0;
}
