// Shared declarations for libbore_b200.so (sm_100a).  See include/bore_b200.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/bore_b200.h"

// ---------------------------------------------------------------- error plumbing
void bore_set_error(const char *fmt, ...);

#define BORE_CHECK(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      bore_set_error(__VA_ARGS__);       \
      return -1;                         \
    }                                    \
  } while (0)

#define BORE_CUDA(call)                                                          \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      bore_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,               \
                     cudaGetErrorString(e__));                                   \
      return -2;                                                                 \
    }                                                                            \
  } while (0)

// ---------------------------------------------------------------- NVTX ranges
// Header-only NVTX 3: a no-op unless a tool (nsys, ncu --nvtx) injects itself; no link dependency.
#include <nvtx3/nvToolsExt.h>
struct BoreNvtxRange {
  explicit BoreNvtxRange(const char *name) { nvtxRangePushA(name); }
  ~BoreNvtxRange() { nvtxRangePop(); }
};
#define BORE_NVTX(name) BoreNvtxRange bore_nvtx_range__(name)

// ---------------------------------------------------------------- model descriptor
// Passed by value to kernels (lives in kernel parameter space / constant bank).
struct MlpDesc {
  int n_layers;                     // Dense layers incl. the final (out_dim 1) layer
  int dims[BORE_MAX_LAYERS + 1];    // dims[0] = D
  int act[BORE_MAX_LAYERS];
  int w_off[BORE_MAX_LAYERS];       // offsets into the flat Keras-order parameter vector
  int b_off[BORE_MAX_LAYERS];
  int n_params;
};

// K0/K2 stage their weights from a pre-packed, zero-padded image (one per kernel variant:
// grad / no grad x narrow / wide lane mapping) with a single bulk copy; the cache is shared by
// shallow copies of the handle.
struct MlpPackCache {
  float *buf[4];
  size_t cap[4];  // floats allocated
};

struct bore_mlp {
  MlpDesc desc;
  MlpPackCache *pack;
  int n_models;
  int device;
  int sm_count;
  float *params;       // [n_models][n_params]
  float *adam_m;       // [n_models][n_params]
  float *adam_v;       // [n_models][n_params]
  long long *adam_t;   // [n_models]   Keras `iterations`
  float lr, beta1, beta2, eps;  // Adam hyper-parameters
  float l2k[BORE_MAX_LAYERS], l2b[BORE_MAX_LAYERS];  // l2 regularisers per layer
  int fit_mode;  // 0 auto, 1 one CTA per model (FFMA), 2 one cluster per model, samples split (FFMA), 3 tensor
                 // pipe, 4 one cluster per model, units split (FFMA2)  (bore_mlp_set_fit_mode)
};

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------- launchers (one per .cu)
// K0 / K2.  `list` (may be NULL) is a device array of row indices: point i of the launch
// is row list[i] of X / f / g (used by the L-BFGS-B driver's compacted active list);
// `n_dev` (may be NULL) is a device int holding the number of points (<= S).
// `prepacked` != 0: the caller has run mlp_eval_prepare for these models since the parameters
// last changed (the L-BFGS-B driver packs once per call, not once per round).
int launch_mlp_eval(const bore_mlp *h, int model, bool want_grad, int transform, int negate,
                    const float *X, int S, float *f, float *g, const int *list,
                    const int *n_dev, cudaStream_t stream, int prepacked = 0);
int mlp_eval_prepare(const bore_mlp *h, int model0, int n_models, bool want_grad, cudaStream_t stream);
// batched problems (BASELINE.json configs[3]): one CTA per model; model model0+b owns points
// [b*per_model, (b+1)*per_model) of X / f / g; `flags` (may be NULL) selects the points to do.
int launch_mlp_eval_multi(const bore_mlp *h, int model0, int n_models, int per_model, bool want_grad,
                          int transform, int negate, const float *X, float *f, float *g,
                          const int *flags, cudaStream_t stream, int prepacked = 0);
// K1t (fit_mma.cu): 1 launched, 0 shape not taken (caller falls back to the FFMA kernels), < 0 error
int launch_fit_mma(const bore_mlp *h, int model0, int count, const float *X_dev, const float *z_dev, int N,
                   int shared_data, int batch_size, int epochs, const int32_t *perm_dev, int shared_perm,
                   float *loss_out_dev, cudaStream_t stream);
// K1u (fit_unit.cu): one cluster per model, hidden units split over the CTAs.  1 launched, 0 shape not
// taken (caller falls back to the sample-split cluster kernel), < 0 error
int launch_fit_unit(const bore_mlp *h, int model0, int count, const float *X_dev, const float *z_dev, int N,
                    int shared_data, int batch_size, int epochs, const int32_t *perm_dev, int shared_perm,
                    float *loss_out_dev, cudaStream_t stream);
