// Host-side interface of the fused persistent argmax kernel (lbfgsb_fused.cu), used by the
// dispatcher in lbfgsb.cu.
#pragma once
#include "common.cuh"
#include "lbfgsb_types.h"

struct FusedLaunchInfo {
  int grid, block;
  size_t smem;
};

// Resume mode: finish the starts of a lock-step run (lbfgsb.cu) that are still active -- their ids
// in list[0 .. *count), each with its persisted state block and a pending request in x_dev.
struct FusedResume {
  const int *list;
  const int *count;             // device
  char *blocks;
  size_t block_stride;
  int *qhead;                   // device int, zeroed by the launcher
  unsigned long long *evals;    // device counter of requests posted (the stepper's)
};

// bytes of device scratch the fused path needs: queue head + evaluation counter + lo / hi / nbd
size_t lbfgsb_fused_workspace_bytes(int n);
// 1 when the model's weights and at least two resident starts fit into an SM's shared memory
int lbfgsb_fused_fits(const bore_mlp *h, int m);
// P_dev: problem description whose lo / hi / nbd point into device memory.  work_dev: >= 16 bytes
// (queue head, evaluation counter), zeroed here.  evals_out != NULL synchronises the stream.
int launch_lbfgsb_fused(const bore_mlp *h, int model0, int n_models, int per_model, int transform,
                        const double *X0_dev, int S, const LbParams &P_dev, void *work_dev,
                        double *x_dev, double *fun_dev, int *nit_dev, int *nfev_dev, int *status_dev,
                        int *task_dev, long long *evals_out, FusedLaunchInfo *info, cudaStream_t stream,
                        const FusedResume *resume = nullptr);
