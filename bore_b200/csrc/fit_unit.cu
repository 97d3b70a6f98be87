// K1u: fused classifier training, ONE model = one thread-block cluster, hidden UNITS split over the CTAs.
//
// Same job as fit.cu (the whole Keras Model.fit with Adam + binary cross-entropy in one launch;
// README.rst:66,93, bore/plugins/hpbandster/base.py:156-157,184), other decomposition.  The first
// cluster kernel (fit_cluster_kernel) splits the SAMPLES of a minibatch over 8 CTAs: every CTA holds
// all weights, carries 8 samples through the net, and each step ends with a reduce-scatter of 8
// partial gradients, Adam on a parameter slice and an all-gather of the new weights into two weight
// images -- 54,600 warp instructions per CTA and step, 22 barriers, 20.5 us per step on B200
// (profiles/r02_ncu_fit_cluster_lines.txt).  Here CTA r of the cluster OWNS units [r U, (r+1) U) of every
// hidden layer:
//   * forward, layer l: h_{l+1}[:, slice] = act(h_l Wc_l + b) from the column slice Wc_l = W_l[:, slice]
//     over ALL samples of the minibatch; the slice is pushed into every CTA's copy of h_{l+1}
//     (st.shared::cluster, 128 contiguous bytes per warp and peer) and one cluster barrier publishes it;
//   * the Dense(1) output layer, the loss, dL/dlogit and the delta of the last hidden layer are
//     computed by every CTA for itself (a few thousand flops: cheaper than one exchange);
//   * reverse, layer l: delta_l[:, slice] = (delta_{l+1} Wr_l') . act'(h_l[:, slice]) from the row slice
//     Wr_l = W_l[slice, :], pushed like the activations (the delta of the first hidden layer stays local);
//   * weight gradients: a CTA computes dW only for the two slices it stores (Wc_l and Wr_l) from the
//     complete activations / deltas it holds, and applies Adam to them in place, straight from the
//     accumulators.  No gradient reduction, no weight exchange, no second weight image: the only
//     traffic between SMs is the activations (4 x 2 KB per CTA and step at cfg 3).
// W_l[k][j] is therefore updated twice, by the owner of column j (in Wc_l) and by the owner of row k
// (in Wr_l).  Both run the same instructions on the same operands in the same order (the sample sum
// is one packed FFMA2 chain over even / odd samples, added at the end), so the two copies stay
// bit-identical for the whole run; the column copies are what is written back.
//
// GEMM mapping inside a CTA (512 threads, out = U units x SP samples, K <= 64..): the K dimension is
// split over the 16 WARPS (warp w takes k = w, w + 16, ...), a lane holds an 8-sample x 2-unit register
// tile (two LDS.128 of activations, one LDS.64 of weights, 8 FFMA2 per k), the 16 partial tiles meet in
// a shared scratch [warp][unit][sample] and thread o sums the 16 partials of output o in a fixed tree,
// applies bias / activation (or act') and pushes.  One __syncthreads per pass; no shuffles.
//
// Rows of every activation / delta buffer are SP + 4 floats apart: the gradient tiles read 8 different
// rows per quarter warp with LDS.128, which this stride spreads over all 32 banks.
#include <stdlib.h>

#include "common.cuh"
#include "fit_common.cuh"

namespace {

constexpr int FU_C = 8;          // CTAs per cluster (portable maximum)
constexpr int FU_THREADS = 512;
constexpr int FU_NW = FU_THREADS / 32;
constexpr int FU_MAXG = 8;       // minibatch elements a thread prefetches into registers

__host__ __device__ inline int fu_r2(int a) { return (a + 1) & ~1; }
__host__ __device__ inline int fu_r4(int a) { return (a + 3) & ~3; }
__host__ __device__ inline int fu_r8(int a) { return (a + 7) & ~7; }

struct FuPlan {
  int SP, SPP;                      // padded batch (multiple of 8) and row stride SP + 4
  int U[BORE_MAX_LAYERS + 1];       // U[i], i = 1..L-1: units of hidden layer i per CTA (even)
  // one copy of this CTA's parameters (floats); Adam m / v follow at + npar / + 2 npar
  int wc[BORE_MAX_LAYERS];          // l = 0..L-2: Wc_l  [dims[l]][U[l+1]]
  int wr[BORE_MAX_LAYERS];          // l = 1..L-2: Wr_l' [dims[l+1]][U[l]]   (row slice, stored transposed)
  int bs[BORE_MAX_LAYERS];          // l = 0..L-2: bias slice [U[l+1]]
  int wl, bl;                       // output layer [dims[L-1]], [1] (every CTA holds and updates it)
  int npar;
  int par0;                         // offset of the parameter block
  int h0[2];                        // minibatch [dims[0]][SPP], double buffered (prefetch)
  int h[BORE_MAX_LAYERS][2];        // h_i, i = 1..L-1: [C U[i]][SPP]; second buffer for i == 1 only
  int dg[BORE_MAX_LAYERS];          // delta_i complete, i = 2..L-1 (i == L-1: computed locally), [C U[i]][SPP]
  int d1;                           // delta_1: this CTA's slice [U[1]][SPP] when L > 2, else = dg[1] complete
  int scratch, wstride;             // [NW][wstride = umax * SP]
  int lg;                           // logit partials [FU_THREADS / SP][SP]
  int dz, zb, idx;                  // dL/dlogit [SP]; labels [2][SP]; row indices [2][SP] (ints)
  int slots;                        // lsum[2][2], reg[2]
  int bars;                         // mbarriers (8 bytes each): h_i at 2 (i - 1) + buffer, delta_i at 2 L + i
  int total;
};

__host__ __device__ inline bool make_fu_plan(const MlpDesc &d, int batch, FuPlan &p) {
  const int L = d.n_layers;
  if (L < 2 || d.dims[L] != 1) return false;
  p.SP = fu_r8(batch);
  p.SPP = p.SP + 4;
  if (p.SP > FU_THREADS) return false;
  int off = 0, umax = 0;
  for (int i = 1; i <= L - 1; ++i) {
    p.U[i] = fu_r2((d.dims[i] + FU_C - 1) / FU_C);
    if (p.U[i] > umax) umax = p.U[i];
  }
  int np = 0;
  for (int l = 0; l <= L - 2; ++l) { p.wc[l] = np; np += d.dims[l] * p.U[l + 1]; }
  for (int l = 1; l <= L - 2; ++l) { p.wr[l] = np; np += d.dims[l + 1] * p.U[l]; }
  for (int l = 0; l <= L - 2; ++l) { p.bs[l] = np; np += p.U[l + 1]; }
  p.wl = np; np += d.dims[L - 1];
  p.bl = np; np += 1;
  p.npar = fu_r4(np);
  p.par0 = off; off += 3 * p.npar;
  p.h0[0] = off; off += d.dims[0] * p.SPP;
  p.h0[1] = off; off += d.dims[0] * p.SPP;
  for (int i = 1; i <= L - 1; ++i) {
    p.h[i][0] = off; off += FU_C * p.U[i] * p.SPP;
    p.h[i][1] = p.h[i][0];
    if (i == 1) { p.h[i][1] = off; off += FU_C * p.U[i] * p.SPP; }
  }
  for (int i = 2; i <= L - 1; ++i) { p.dg[i] = off; off += FU_C * p.U[i] * p.SPP; }
  if (L > 2) { p.d1 = off; off += p.U[1] * p.SPP; }
  else { p.dg[1] = off; p.d1 = off; off += FU_C * p.U[1] * p.SPP; }
  p.wstride = umax * p.SP;
  p.scratch = off; off += FU_NW * p.wstride;
  p.lg = off; off += (FU_THREADS / p.SP) * p.SP;
  p.dz = off; off += p.SP;
  p.zb = off; off += 2 * p.SP;
  p.idx = off; off += 2 * p.SP;
  p.slots = off; off += 8;
  off = fu_r2(off);  // mbarriers are 8 bytes
  p.bars = off; off += 2 * (3 * BORE_MAX_LAYERS + 2);
  p.total = fu_r4(off);
  return true;
}

struct FuArgs {
  MlpDesc d;
  FuPlan P;
  float *params, *adam_m, *adam_v;
  long long *adam_t;
  int model0;
  const float *X, *z;
  int N, shared_data, batch, epochs;
  const int *perm;
  int shared_perm;
  float l2k[BORE_MAX_LAYERS], l2b[BORE_MAX_LAYERS];
  int any_l2;
  float *loss_out;
  float lr, beta1, beta2, eps;
};

// ---------------------------------------------------------------- cluster / packed-FMA primitives
__device__ __forceinline__ uint32_t fu_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void fu_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t fu_peer(const void *p, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void fu_st(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// asynchronous remote store that signals the destination CTA's mbarrier (complete_tx of 4 bytes): data and
// signal travel together, no fence and no cluster-wide barrier on the producer side
__device__ __forceinline__ void fu_st_async(uint32_t addr, float v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr),
               "r"(__float_as_uint(v)), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fu_bar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fu_bar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fu_bar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "FU_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra FU_DONE;\n"
      "bra FU_WAIT;\n"
      "FU_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ float fu_ld(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
// FFMA2 (PTX fma.rn.f32x2): two IEEE fp32 FMAs per instruction, the same roundings as two fmaf
__device__ __forceinline__ uint64_t fu_pk(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void fu_fma2(uint64_t &c, uint64_t a, uint64_t b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}
__device__ __forceinline__ float2 fu_upk(uint64_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

// ---------------------------------------------------------------- GEMM pass, first half
// scratch[warp][u * SP + s] = sum over this warp's k of A[k][s] * Wm[k][u]     (u < U, s < SP)
// A: [K][SPP], Wm: [K][U].  Lane tile: samples {4 so .. 4 so + 3} and {SP/2 + 4 so ..}, units 2 up, 2 up + 1.
__device__ __forceinline__ void fu_partial(const float *__restrict__ A, const float *__restrict__ Wm, int K, int U,
                                           int SP, int SPP, float inv_nso, float *__restrict__ mine, int warp, int lane) {
  const int nso = SP >> 3, ntile = nso * (U >> 1), half = SP >> 1;
  for (int tile = lane; tile < ntile; tile += 32) {
    const int up = fdiv(tile, inv_nso), so = tile - up * nso;
    uint64_t acc[2][4];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[e][i] = 0ull;
    const float *ap = A + 4 * so + warp * SPP;
    const float *wp = Wm + 2 * up + warp * U;
    const int astep = FU_NW * SPP, wstep = FU_NW * U;
    // four k per trip (K <= 64: one trip), all twelve loads in flight before the first FFMA2; a k beyond K
    // reads row K - 1 again and multiplies it by zero weights (w = 0): no branch in the chain
    for (int k0 = warp; k0 < K; k0 += 4 * FU_NW) {
      float4 a0[4], a1[4];
      float2 w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = k0 + i * FU_NW < K;
        const float *ai = ok ? ap + i * astep : ap;
        a0[i] = *reinterpret_cast<const float4 *>(ai);
        a1[i] = *reinterpret_cast<const float4 *>(ai + half);
        w[i] = ok ? *reinterpret_cast<const float2 *>(wp + i * wstep) : make_float2(0.f, 0.f);
      }
      ap += 4 * astep;
      wp += 4 * wstep;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint64_t a01 = fu_pk(a0[i].x, a0[i].y), a23 = fu_pk(a0[i].z, a0[i].w);
        const uint64_t a45 = fu_pk(a1[i].x, a1[i].y), a67 = fu_pk(a1[i].z, a1[i].w);
        const uint64_t w0 = fu_pk(w[i].x, w[i].x), w1 = fu_pk(w[i].y, w[i].y);
        fu_fma2(acc[0][0], a01, w0); fu_fma2(acc[0][1], a23, w0); fu_fma2(acc[0][2], a45, w0); fu_fma2(acc[0][3], a67, w0);
        fu_fma2(acc[1][0], a01, w1); fu_fma2(acc[1][1], a23, w1); fu_fma2(acc[1][2], a45, w1); fu_fma2(acc[1][3], a67, w1);
      }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float *o = mine + (2 * up + e) * SP + 4 * so;
      const float2 p0 = fu_upk(acc[e][0]), p1 = fu_upk(acc[e][1]), p2 = fu_upk(acc[e][2]), p3 = fu_upk(acc[e][3]);
      *reinterpret_cast<float4 *>(o) = make_float4(p0.x, p0.y, p1.x, p1.y);
      *reinterpret_cast<float4 *>(o + half) = make_float4(p2.x, p2.y, p3.x, p3.y);
    }
  }
}

// ---------------------------------------------------------------- GEMM pass, second half
// epi(u, s, sum of the 16 partials of output (u, s)) for u < U, s < SP; fixed summation tree
template <class Epi>
__device__ __forceinline__ void fu_reduce(const float *__restrict__ scratch, int U, int SP, float inv_sp, int wstride,
                                          Epi epi) {
  const int n = U * SP;
  for (int o = threadIdx.x; o < n; o += FU_THREADS) {
    float p[FU_NW];
#pragma unroll
    for (int w = 0; w < FU_NW; ++w) p[w] = scratch[w * wstride + o];
#pragma unroll
    for (int st = 1; st < FU_NW; st <<= 1)
#pragma unroll
      for (int w = 0; w < FU_NW; w += 2 * st) p[w] += p[w + st];
    const int u = fdiv(o, inv_sp);
    epi(u, o - u * SP, p[0]);
  }
}

// One 4 x 2 block of a weight gradient: g[i][e] = sum_s RA_i[s] * RB_e[s] over s < SP, as an even / odd
// packed chain (FFMA2 on (s, s + 1) pairs), lo + hi at the end.  Operand order does not matter (the
// products commute), so the column copy (A = activations, B = deltas) and the row copy (A = deltas,
// B = activations) of the same weight get the same bits.
__device__ __forceinline__ void fu_grad_tile(const float *__restrict__ a0, const float *__restrict__ a1,
                                             const float *__restrict__ a2, const float *__restrict__ a3,
                                             const float *__restrict__ b0, const float *__restrict__ b1, int SP,
                                             float (&g)[4][2]) {
  uint64_t acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i][0] = 0ull; acc[i][1] = 0ull; }
#pragma unroll 2
  for (int s = 0; s < SP; s += 4) {
    const float4 x0 = *reinterpret_cast<const float4 *>(a0 + s), x1 = *reinterpret_cast<const float4 *>(a1 + s);
    const float4 x2 = *reinterpret_cast<const float4 *>(a2 + s), x3 = *reinterpret_cast<const float4 *>(a3 + s);
    const float4 y0 = *reinterpret_cast<const float4 *>(b0 + s), y1 = *reinterpret_cast<const float4 *>(b1 + s);
    const uint64_t y0a = fu_pk(y0.x, y0.y), y0b = fu_pk(y0.z, y0.w), y1a = fu_pk(y1.x, y1.y), y1b = fu_pk(y1.z, y1.w);
    const float4 xs[4] = {x0, x1, x2, x3};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint64_t xa = fu_pk(xs[i].x, xs[i].y), xb = fu_pk(xs[i].z, xs[i].w);
      fu_fma2(acc[i][0], xa, y0a);
      fu_fma2(acc[i][1], xa, y1a);
      fu_fma2(acc[i][0], xb, y0b);
      fu_fma2(acc[i][1], xb, y1b);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float2 v = fu_upk(acc[i][e]);
      g[i][e] = v.x + v.y;
    }
}
// the same chain for one row pair (output layer, biases through a row of ones are not needed: see fu_row_sum)
__device__ __forceinline__ float fu_row_dot(const float *__restrict__ a, const float *__restrict__ b, int SP) {
  uint64_t acc = 0ull;
#pragma unroll 4
  for (int s = 0; s < SP; s += 4) {
    const float4 x = *reinterpret_cast<const float4 *>(a + s), y = *reinterpret_cast<const float4 *>(b + s);
    fu_fma2(acc, fu_pk(x.x, x.y), fu_pk(y.x, y.y));
    fu_fma2(acc, fu_pk(x.z, x.w), fu_pk(y.z, y.w));
  }
  const float2 v = fu_upk(acc);
  return v.x + v.y;
}
__device__ __forceinline__ float fu_row_sum(const float *__restrict__ a, int SP) {
  float e = 0.f, o = 0.f;
#pragma unroll 4
  for (int s = 0; s < SP; s += 4) {
    const float4 x = *reinterpret_cast<const float4 *>(a + s);
    e += x.x; o += x.y; e += x.z; o += x.w;
  }
  return e + o;
}

// ASYNC: exchanges by st.async + mbarrier (a CTA waits for ITS copy to be complete, nothing else); otherwise
// plain remote stores + barrier.cluster after every exchange (kept for A/B timing, BORE_FIT_UNIT_ASYNC=0)
template <bool ASYNC>
__global__ void __cluster_dims__(FU_C, 1, 1) __launch_bounds__(FU_THREADS) fit_unit_kernel(const FuArgs a) {
  extern __shared__ __align__(16) float sm[];
  const MlpDesc &d = a.d;
  const FuPlan &P = a.P;
  const int L = d.n_layers, SP = P.SP, SPP = P.SPP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)fu_rank();
  const int cl = blockIdx.x / FU_C;  // which model of the launch
  const int model = a.model0 + cl;
  float *gp = a.params + (size_t)model * d.n_params;
  float *gm = a.adam_m + (size_t)model * d.n_params;
  float *gv = a.adam_v + (size_t)model * d.n_params;
  const float *X = a.X + (a.shared_data ? 0 : (size_t)cl * a.N * d.dims[0]);
  const float *zg = a.z + (a.shared_data ? 0 : (size_t)cl * a.N);
  const int *perm = a.perm + (a.shared_perm ? 0 : (size_t)cl * a.epochs * a.N);
  const int D = d.dims[0];
  const int steps_per_epoch = (a.N + a.batch - 1) / a.batch;
  float *W = sm + P.par0, *Mm = W + P.npar, *Vv = Mm + P.npar;
  const int wstride = P.wstride;  // floats per warp of the scratch
  const float inv_sp = 1.f / (float)SP, inv_nso = 1.f / (float)(SP >> 3), inv_nq = 1.f / (float)(SP >> 2);

  // ---- zero everything, then stage this CTA's parameter slices and their Adam slots ----
  for (int i = tid; i < P.total; i += FU_THREADS) sm[i] = 0.f;
  __syncthreads();
  const uint32_t bars = (uint32_t)__cvta_generic_to_shared(sm + P.bars);  // local mbarriers, 8 bytes apart
  if (ASYNC && tid == 0) {
    for (int i = 0; i < 3 * BORE_MAX_LAYERS + 2; ++i) fu_bar_init(bars + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int l = 0; l <= L - 2; ++l) {
    const int in = d.dims[l], out = d.dims[l + 1], U = P.U[l + 1];
    for (int e = tid; e < in * U; e += FU_THREADS) {
      const int k = e / U, u = e - k * U, j = rank * U + u;
      if (j < out) {
        const int gi = d.w_off[l] + k * out + j;
        W[P.wc[l] + e] = gp[gi]; Mm[P.wc[l] + e] = gm[gi]; Vv[P.wc[l] + e] = gv[gi];
      }
    }
    for (int u = tid; u < U; u += FU_THREADS) {
      const int j = rank * U + u;
      if (j < out) {
        const int gi = d.b_off[l] + j;
        W[P.bs[l] + u] = gp[gi]; Mm[P.bs[l] + u] = gm[gi]; Vv[P.bs[l] + u] = gv[gi];
      }
    }
  }
  for (int l = 1; l <= L - 2; ++l) {
    const int in = d.dims[l], out = d.dims[l + 1], U = P.U[l];
    for (int e = tid; e < out * U; e += FU_THREADS) {
      const int j = e / U, u = e - j * U, k = rank * U + u;
      if (k < in) {
        const int gi = d.w_off[l] + k * out + j;
        W[P.wr[l] + e] = gp[gi]; Mm[P.wr[l] + e] = gm[gi]; Vv[P.wr[l] + e] = gv[gi];
      }
    }
  }
  for (int k = tid; k <= d.dims[L - 1]; k += FU_THREADS) {  // output layer: weights, then the bias
    const int gi = k < d.dims[L - 1] ? d.w_off[L - 1] + k : d.b_off[L - 1];
    W[P.wl + k] = gp[gi]; Mm[P.wl + k] = gm[gi]; Vv[P.wl + k] = gv[gi];
  }
  long long t_step = a.adam_t[model];
  double b1p_d = pow((double)a.beta1, (double)t_step), b2p_d = pow((double)a.beta2, (double)t_step);
  uint32_t peer[FU_C];  // base of every CTA's dynamic shared memory
#pragma unroll
  for (int c = 0; c < FU_C; ++c) peer[c] = fu_peer(sm, c);

  // ---- minibatch prefetch: element e = k * SP + p of the next step -> registers -> h0[buf][k][p] ----
  // The row indices of a step are put into shared memory (idx[buf]) one exchange earlier by the threads
  // p < SP, so an element costs one LDS + one LDG (lanes walk p: conflict-free stores into h0[k][p]).
  const bool greg = SP * D <= FU_THREADS * FU_MAXG;
  float gx[FU_MAXG];
  float gz = 0.f;
  int *idxs = reinterpret_cast<int *>(sm + P.idx);
  auto row_of = [&](int ep, int st, int p) -> int {
    const int s0 = st * a.batch;
    const int nb = min(a.batch, a.N - s0);
    return p < nb ? perm[(size_t)ep * a.N + s0 + p] : -1;
  };
  auto index_stage = [&](int ep, int st, int buf) {  // visible after the next __syncthreads
    if (tid < SP) idxs[buf * SP + tid] = row_of(ep, st, tid);
  };
  auto gather_issue = [&](int buf) {
    const int *ix = idxs + buf * SP;
    if (greg) {
#pragma unroll
      for (int i = 0; i < FU_MAXG; ++i) {
        const int e = tid + i * FU_THREADS;
        gx[i] = 0.f;
        if (e < SP * D) {
          const int k = fdiv(e, inv_sp), p = e - k * SP;
          const int row = ix[p];
          if (row >= 0) gx[i] = X[(size_t)row * D + k];
        }
      }
    }
    if (tid < SP) {
      const int row = ix[tid];
      gz = row >= 0 ? zg[row] : 0.f;
    }
  };
  auto gather_commit = [&](int buf) {
    float *H0 = sm + P.h0[buf];
    if (greg) {
#pragma unroll
      for (int i = 0; i < FU_MAXG; ++i) {
        const int e = tid + i * FU_THREADS;
        if (e < SP * D) {
          const int k = fdiv(e, inv_sp), p = e - k * SP;
          H0[k * SPP + p] = gx[i];
        }
      }
    } else {
      const int *ix = idxs + buf * SP;
      for (int e = tid; e < SP * D; e += FU_THREADS) {
        const int k = fdiv(e, inv_sp), p = e - k * SP;
        const int row = ix[p];
        H0[k * SPP + p] = row >= 0 ? X[(size_t)row * D + k] : 0.f;
      }
    }
    if (tid < SP) sm[P.zb + buf * SP + tid] = gz;
  };
  index_stage(0, 0, 0);
  __syncthreads();
  gather_issue(0);
  gather_commit(0);
  fu_cluster_sync();  // everybody's shared memory exists and is initialised

  const int NLG = FU_THREADS / SP;  // k-parts of the logit
  const int HL = d.dims[L - 1];
  int n_tiles = 0, n_items = 0;
  for (int l = 0; l <= L - 2; ++l) n_tiles += ((d.dims[l] + 3) >> 2) * (P.U[l + 1] >> 1);
  for (int l = 1; l <= L - 2; ++l) n_tiles += ((d.dims[l + 1] + 3) >> 2) * (P.U[l] >> 1);
  n_items = n_tiles + HL + 1;
  for (int l = 0; l <= L - 2; ++l) n_items += P.U[l + 1];
  int par = 0;  // parity of the current step: minibatch buffer, h_1 buffer, loss slots
  uint32_t step_no = 0;  // steps done: phase parity of the mbarriers
  float epoch_tot = 0.f;
  for (int ep = 0; ep < a.epochs; ++ep) {
    for (int st = 0; st < steps_per_epoch; ++st, par ^= 1, ++step_no) {
      const int s0 = st * a.batch;
      const int nb = min(a.batch, a.N - s0);
      int nst = st + 1, nep = ep;
      if (nst == steps_per_epoch) { nst = 0; ++nep; }
      const bool more = nep < a.epochs;
      if (more) index_stage(nep, nst, par ^ 1);  // read by gather_issue after the forward pass

      // ---- forward through the hidden layers: partial products, then reduce + activation + push ----
      for (int l = 0; l <= L - 2; ++l) {
        const int U = P.U[l + 1];
        const float *A = sm + (l == 0 ? P.h0[par] : P.h[l][l == 1 ? par : 0]);
        const int act = d.act[l], outd = d.dims[l + 1];
        // this exchange's mbarrier: h_1 alternates between two buffers (one use every other step)
        const uint32_t bidx = 8u * (uint32_t)(2 * l + (l == 0 ? par : 0));
        const uint32_t bph = (l == 0 ? step_no >> 1 : step_no) & 1u;
        if (ASYNC && tid == 0) fu_bar_expect(bars + bidx, (uint32_t)(outd * SP) * 4u);
        fu_partial(A, W + P.wc[l], d.dims[l], U, SP, SPP, inv_nso, sm + P.scratch + warp * wstride, warp, lane);
        __syncthreads();
        const float *bs = W + P.bs[l];
        const uint32_t hoff = (uint32_t)(P.h[l + 1][l == 0 ? par : 0]) * 4u;
        const uint32_t boff = (uint32_t)P.bars * 4u + bidx;
        fu_reduce(sm + P.scratch, U, SP, inv_sp, wstride, [&](int u, int s, float sum) {
          const int unit = rank * U + u;
          if (unit < outd) {
            const float v = f_act(act, sum + bs[u]);
            const uint32_t off = hoff + (uint32_t)(unit * SPP + s) * 4u;
#pragma unroll
            for (int c = 0; c < FU_C; ++c) {
              if (ASYNC) fu_st_async(peer[c] + off, v, peer[c] + boff);
              else fu_st(peer[c] + off, v);
            }
          }
        });
        if (ASYNC) {
          __syncthreads();  // the scratch is free for the next pass
          fu_bar_wait(bars + bidx, bph);  // every unit of h_{l+1} has landed in THIS CTA's copy
        } else {
          fu_cluster_sync();  // h_{l+1} complete in every CTA
        }
      }

      // ---- Dense(1) logit, loss, dL/dlogit: every CTA for itself ----
      const float *HLb = sm + P.h[L - 1][L - 1 == 1 ? par : 0];
      {
        const int kp = fdiv(tid, inv_sp), s = tid - kp * SP;
        if (kp < NLG) {
          float acc = 0.f;
          for (int k = kp; k < HL; k += NLG) acc = fmaf(HLb[k * SPP + s], W[P.wl + k], acc);
          sm[P.lg + kp * SP + s] = acc;
        }
      }
      __syncthreads();
      const float inv_nb = 1.f / (float)nb;
      if (warp < 2) {  // SP <= 64: the samples sit in warps 0 and 1 (checked by the launcher)
        float dl = 0.f, lt = 0.f;
        if (tid < SP) {
          float u = W[P.bl];
          for (int kp = 0; kp < NLG; ++kp) u += sm[P.lg + kp * SP + tid];
          if (tid < nb) {
            const float zz = sm[P.zb + par * SP + tid];
            lt = fmaxf(u, 0.f) - u * zz + log1pf(expf(-fabsf(u)));
            dl = (stable_sigmoid(u) - zz) * inv_nb;
          }
          sm[P.dz + tid] = dl;
        }
        for (int o = 16; o > 0; o >>= 1) lt += __shfl_xor_sync(0xffffffffu, lt, o);
        if (lane == 0) sm[P.slots + 2 * par + warp] = lt;
      }

      // ---- prefetch the next minibatch; Adam scalars of this step (Keras: t starts at 1) ----
      if (more) gather_issue(par ^ 1);
      t_step += 1;
      b1p_d *= (double)a.beta1;
      b2p_d *= (double)a.beta2;
      const float b1p = (float)b1p_d, b2p = (float)b2p_d;
      const float alpha = a.lr * sqrtf(1.f - b2p) / (1.f - b1p);
      const float om1 = 1.f - a.beta1, om2 = 1.f - a.beta2;
      __syncthreads();  // dz and the loss slots are in place

      // ---- delta of the last hidden layer (elementwise, complete, every CTA) ----
      {
        float *DF = sm + (L - 1 >= 2 ? P.dg[L - 1] : P.dg[1]);
        const int nq = SP >> 2, actp = d.act[L - 2];
        for (int e = tid; e < HL * nq; e += FU_THREADS) {
          const int k = fdiv(e, inv_nq), q = e - k * nq;
          const float4 hv = *reinterpret_cast<const float4 *>(HLb + k * SPP + 4 * q);
          const float4 dz = *reinterpret_cast<const float4 *>(sm + P.dz + 4 * q);
          const float w = W[P.wl + k];
          float4 o;
          o.x = dz.x * w * f_act_bwd(actp, hv.x);
          o.y = dz.y * w * f_act_bwd(actp, hv.y);
          o.z = dz.z * w * f_act_bwd(actp, hv.z);
          o.w = dz.w * w * f_act_bwd(actp, hv.w);
          *reinterpret_cast<float4 *>(DF + k * SPP + 4 * q) = o;
        }
      }
      __syncthreads();

      // ---- reverse through the hidden layers: delta_i slice from delta_{i+1} and the row slice Wr_i ----
      for (int i = L - 2; i >= 1; --i) {
        const int U = P.U[i];
        if (ASYNC && i >= 2 && tid == 0) fu_bar_expect(bars + 8u * (uint32_t)(2 * L + i), (uint32_t)(d.dims[i] * SP) * 4u);
        fu_partial(sm + P.dg[i + 1], W + P.wr[i], d.dims[i + 1], U, SP, SPP, inv_nso,
                   sm + P.scratch + warp * wstride, warp, lane);
        __syncthreads();
        const float *Hi = sm + P.h[i][i == 1 ? par : 0];
        const int actp = d.act[i - 1], width = d.dims[i];
        if (i >= 2) {
          const uint32_t doff = (uint32_t)P.dg[i] * 4u;
          const uint32_t bidx = 8u * (uint32_t)(2 * L + i), boff = (uint32_t)P.bars * 4u + bidx;
          fu_reduce(sm + P.scratch, U, SP, inv_sp, wstride, [&](int u, int s, float sum) {
            const int unit = rank * U + u;
            if (unit < width) {
              const float v = sum * f_act_bwd(actp, Hi[unit * SPP + s]);
              const uint32_t off = doff + (uint32_t)(unit * SPP + s) * 4u;
#pragma unroll
              for (int c = 0; c < FU_C; ++c) {
                if (ASYNC) fu_st_async(peer[c] + off, v, peer[c] + boff);
                else fu_st(peer[c] + off, v);
              }
            }
          });
          if (ASYNC) {
            __syncthreads();
            fu_bar_wait(bars + bidx, step_no & 1u);
          } else {
            fu_cluster_sync();
          }
        } else {
          float *D1 = sm + P.d1;
          fu_reduce(sm + P.scratch, U, SP, inv_sp, wstride, [&](int u, int s, float sum) {
            const int unit = rank * U + u;
            D1[u * SPP + s] = unit < width ? sum * f_act_bwd(actp, Hi[unit * SPP + s]) : 0.f;
          });
          __syncthreads();
        }
      }

      // ---- weight gradients of the slices this CTA stores, Adam in place ----
      // Work items, one per thread (more than FU_THREADS: a second trip): 4 x 2 tiles of
      //   (a) the column copies Wc_l (rows: units of layer l, activations h_l; columns: own units of layer l + 1),
      //   (b) the row copies Wr_l' (rows: units of layer l + 1, deltas; columns: own units of layer l),
      // then (c) single outputs: output-layer weights + bias (every CTA, identical) and the bias slices.
      // An item is decoded first (cheap, divergent) and computed after (convergent per kind).
      float reg = 0.f;
      const float *dzv = sm + P.dz;
      for (int item0 = 0; item0 < n_items; item0 += FU_THREADS) {
        const int item = item0 + tid;
        int kind = -1;  // 0 tile, 1 single output
        const float *Ab = nullptr, *Bb = nullptr;
        int RA = 1, Uc = 2, nta = 1, ta = 0, tu = 0, pbase = 0, wlimit = 0, colc = 0;
        float l2 = 0.f;
        int pi1 = 0, cnt_reg = 0;
        if (item < n_tiles) {
          kind = 0;
          int base = 0;
          for (int blk = 0; blk < 2 * L - 3; ++blk) {
            const bool colcopy = blk <= L - 2;
            const int l = colcopy ? blk : blk - (L - 2);  // (a): l = 0..L-2, (b): l = 1..L-2
            const int ra_ = colcopy ? d.dims[l] : d.dims[l + 1];
            const int u_ = colcopy ? P.U[l + 1] : P.U[l];
            const int nta_ = (ra_ + 3) >> 2, nt = nta_ * (u_ >> 1);
            if (item < base + nt) {
              const int it = item - base;
              RA = ra_; Uc = u_; nta = nta_;
              tu = fdiv(it, 1.f / (float)nta); ta = it - tu * nta;
              const float *hl = sm + (l == 0 ? P.h0[par] : P.h[l][l == 1 ? par : 0]);  // h_l
              // delta_{l+1}: complete buffer, or the local slice when it belongs to the first hidden layer of a
              // deeper net (only the column copy of W_0 reads that one)
              const bool dslice = (l + 1 == 1) && L > 2;
              const float *dn = sm + (dslice ? P.d1 : P.dg[l + 1]);
              if (colcopy) {
                Ab = hl;
                Bb = dn + (dslice ? 0 : rank * Uc * SPP);
                wlimit = d.dims[l + 1] - rank * Uc;
                pbase = P.wc[l];
              } else {
                Ab = dn;
                Bb = hl + rank * Uc * SPP;
                wlimit = d.dims[l] - rank * Uc;
                pbase = P.wr[l];
              }
              colc = colcopy ? 1 : 0;
              l2 = a.l2k[l];
              break;
            }
            base += nt;
          }
        } else if (item < n_items) {
          int k = item - n_tiles;
          if (k <= HL) {  // output layer: weights 0..HL-1, bias at HL
            kind = 1;
            Ab = k < HL ? HLb + k * SPP : nullptr;
            Bb = dzv;
            pi1 = P.wl + k;
            l2 = k < HL ? a.l2k[L - 1] : a.l2b[L - 1];
            cnt_reg = rank == 0;
          } else {
            k -= HL + 1;
            for (int l = 0; l <= L - 2; ++l) {
              const int u_ = P.U[l + 1];
              if (k < u_) {
                if (rank * u_ + k < d.dims[l + 1]) {
                  kind = 1;
                  const bool dslice = (l + 1 == 1) && L > 2;
                  Bb = sm + (dslice ? P.d1 + k * SPP : P.dg[l + 1] + (rank * u_ + k) * SPP);
                  pi1 = P.bs[l] + k;
                  l2 = a.l2b[l];
                  cnt_reg = 1;
                }
                break;
              }
              k -= u_;
            }
          }
        }
        if (kind == 0) {
          int ra[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) ra[i] = min(ta + i * nta, RA - 1);
          float g[4][2];
          fu_grad_tile(Ab + ra[0] * SPP, Ab + ra[1] * SPP, Ab + ra[2] * SPP, Ab + ra[3] * SPP, Bb + (2 * tu) * SPP,
                       Bb + (2 * tu + 1) * SPP, SP, g);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (ta + i * nta >= RA) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (2 * tu + e >= wlimit) continue;
              const int pi = pbase + ra[i] * Uc + 2 * tu + e;
              float wv = W[pi], gg = g[i][e], m = Mm[pi], v = Vv[pi];
              if (l2 != 0.f) { if (colc) reg += l2 * wv * wv; gg += 2.f * l2 * wv; }
              wv = adam_update(wv, gg, m, v, om1, om2, alpha, a.eps);
              W[pi] = wv; Mm[pi] = m; Vv[pi] = v;
            }
          }
        } else if (kind == 1) {
          float gg = Ab ? fu_row_dot(Ab, Bb, SP) : fu_row_sum(Bb, SP);
          float wv = W[pi1], m = Mm[pi1], v = Vv[pi1];
          if (l2 != 0.f) { if (cnt_reg) reg += l2 * wv * wv; gg += 2.f * l2 * wv; }
          wv = adam_update(wv, gg, m, v, om1, om2, alpha, a.eps);
          W[pi1] = wv; Mm[pi1] = m; Vv[pi1] = v;
        }
      }
      if (more) gather_commit(par ^ 1);
      if (rank == 0 && tid == 0) epoch_tot += (sm[P.slots + 2 * par] + sm[P.slots + 2 * par + 1]) * inv_nb * (float)nb;
      if (a.any_l2) {
        // regulariser terms: every CTA sums those of the parameters it owns (column copies, bias slices; rank 0
        // the output layer), rank 0 collects them after one more cluster barrier (parity slots: the value is
        // overwritten two steps later, which no CTA reaches before rank 0 has passed the next barrier)
        reg = block_sum(reg, sm + P.scratch);
        if (tid == 0) sm[P.slots + 4 + par] = reg;
        fu_cluster_sync();
        if (rank == 0 && tid == 0) {
          float rg = 0.f;
          for (int c = 0; c < FU_C; ++c) rg += fu_ld(peer[c] + (uint32_t)(P.slots + 4 + par) * 4u);
          epoch_tot += rg * (float)nb;
        }
      }
      __syncthreads();  // weights updated, next minibatch in place
    }
    if (rank == 0 && tid == 0 && a.loss_out) a.loss_out[(size_t)cl * a.epochs + ep] = epoch_tot / (float)a.N;
    epoch_tot = 0.f;
  }

  // ---- write back: the column copies partition W_l; rank 0 writes the output layer ----
  for (int l = 0; l <= L - 2; ++l) {
    const int in = d.dims[l], out = d.dims[l + 1], U = P.U[l + 1];
    for (int e = tid; e < in * U; e += FU_THREADS) {
      const int k = e / U, u = e - k * U, j = rank * U + u;
      if (j < out) {
        const int gi = d.w_off[l] + k * out + j;
        gp[gi] = W[P.wc[l] + e]; gm[gi] = Mm[P.wc[l] + e]; gv[gi] = Vv[P.wc[l] + e];
      }
    }
    for (int u = tid; u < U; u += FU_THREADS) {
      const int j = rank * U + u;
      if (j < out) {
        const int gi = d.b_off[l] + j;
        gp[gi] = W[P.bs[l] + u]; gm[gi] = Mm[P.bs[l] + u]; gv[gi] = Vv[P.bs[l] + u];
      }
    }
  }
  if (rank == 0) {
    for (int k = tid; k <= HL; k += FU_THREADS) {
      const int gi = k < HL ? d.w_off[L - 1] + k : d.b_off[L - 1];
      gp[gi] = W[P.wl + k]; gm[gi] = Mm[P.wl + k]; gv[gi] = Vv[P.wl + k];
    }
    if (tid == 0) a.adam_t[model] = t_step;
  }
  fu_cluster_sync();  // nobody exits while a peer may still address its shared memory
}

}  // namespace

// 1 launched, 0 shape not taken (the caller falls back to the sample-split cluster kernel), < 0 error
int launch_fit_unit(const bore_mlp *h, int model0, int count, const float *X_dev, const float *z_dev, int N,
                    int shared_data, int batch_size, int epochs, const int32_t *perm_dev, int shared_perm,
                    float *loss_out_dev, cudaStream_t stream) {
  FuArgs a;
  a.d = h->desc;
  const int B = batch_size < N ? batch_size : N;
  if (!make_fu_plan(a.d, B, a.P)) return 0;
  if (a.P.SP > 64) return 0;  // the loss reduction assumes the samples sit in two warps
  const size_t smem = (size_t)a.P.total * sizeof(float);
  if (smem > 226 * 1024) return 0;
  a.params = h->params; a.adam_m = h->adam_m; a.adam_v = h->adam_v; a.adam_t = h->adam_t;
  a.model0 = model0;
  a.X = X_dev; a.z = z_dev; a.N = N; a.shared_data = shared_data; a.batch = batch_size;
  a.epochs = epochs; a.perm = perm_dev; a.shared_perm = shared_perm;
  a.any_l2 = 0;
  for (int l = 0; l < BORE_MAX_LAYERS; ++l) {
    a.l2k[l] = l < a.d.n_layers ? h->l2k[l] : 0.f;
    a.l2b[l] = l < a.d.n_layers ? h->l2b[l] : 0.f;
    if (a.l2k[l] != 0.f || a.l2b[l] != 0.f) a.any_l2 = 1;
  }
  a.loss_out = loss_out_dev;
  a.lr = h->lr; a.beta1 = h->beta1; a.beta2 = h->beta2; a.eps = h->eps;
  static const bool use_async = [] { const char *e = getenv("BORE_FIT_UNIT_ASYNC"); return !(e && e[0] == '0'); }();
  if (use_async) {
    BORE_CUDA(cudaFuncSetAttribute(fit_unit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fit_unit_kernel<true><<<count * FU_C, FU_THREADS, smem, stream>>>(a);
  } else {
    BORE_CUDA(cudaFuncSetAttribute(fit_unit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fit_unit_kernel<false><<<count * FU_C, FU_THREADS, smem, stream>>>(a);
  }
  BORE_CUDA(cudaGetLastError());
  return 1;
}
