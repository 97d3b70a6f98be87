// L-BFGS-B (v3.0) for ONE start -- the algorithm, written once and compiled in three variants.
//
// Replaces scipy.optimize.minimize(method="L-BFGS-B", jac=True, bounds=...) as called per
// start at bore/mixins.py:59-60 and bore/optimizers/base.py:59.  SciPy is a third-party
// dependency of the reference (scipy==1.7.0, setup.py:15) whose L-BFGS-B sources are not in
// /root/reference; this file restates the PUBLISHED algorithm (Byrd, Lu, Nocedal & Zhu 1995;
// Zhu, Byrd, Lu & Nocedal 1997, Alg. 778; Morales & Nocedal 2011 = v3.0; More' & Thuente
// 1994 line search) in the structure SciPy wraps: generalized Cauchy point, subspace
// minimisation with projection, dcsrch line search, limited-memory BFGS update, and the
// driver loop of scipy/optimize/_lbfgsb_py.py:406-443 (nit / maxiter / maxfun / status).
//
// Variants (the includer defines LB_VARIANT; lbfgsb_types.h holds what is common):
//
//   LB_VARIANT 1  "warp"    the product: one WARP owns one start, its state staged in shared
//                 memory.  Vectors of length n are spread over the 32 lanes (LB_FOR),
//                 reductions are xor-butterflies (bit-identical on every lane), the small
//                 factorisations run column-parallel, triangular solves keep their vector in
//                 registers and pass the pivot by shuffle, scalar control flow is replicated
//                 on every lane (warp-uniform).
//   LB_VARIANT 0  "host"    the same algorithm as a plain serial program compiled by g++:
//                 TEST INFRASTRUCTURE, pinned against SciPy's setulb request by request in
//                 tests/ (the product only ever runs the device variant).  Where the warp
//                 variant walks an index list, the serial one sweeps all n variables with a
//                 mask; the sums are the same up to rounding.
//
// All algebra is fp64 (SciPy's is); the objective and gradient arrive as fp32 values from
// the MLP kernel, exactly like the reference's fp32 Keras model feeding fp64 SciPy
// (bore/decorators.py:54-56).
#include "lbfgsb_types.h"

#ifndef LB_VARIANT
#error "define LB_VARIANT (0 host serial, 1 device warp) before including lbfgsb_core.h"
#endif

#undef LB_FN
#undef LB_M
#undef LB_WARP
#undef LB_LANE
#undef LB_NL
#undef LB_SYNC
#undef LB_FOR
#undef LB_UNROLL1
#undef LB_UNROLL_HOT
#undef LB_UNROLL_HOT2
#undef LB_NI
#undef LB_SHARED
// Code size is a first-order cost on the device: the stepper is ~10^4 mostly-serial
// instructions per step and every warp of an SM sits at a different place in it, so a kernel
// that does not fit the instruction caches stalls on instruction fetch (ncu: no_instruction
// was 55 % of all stall cycles with everything inlined and unrolled, profiles/r01_notes.md).
// Hence: inner loops stay rolled (LB_UNROLL1), helpers with several call sites are real
// functions (LB_NI), and their pointer arguments are declared shared (LB_SHARED) so that
// they still compile to LDS/STS.
#if LB_VARIANT == 0
#define LB_UNROLL1
#define LB_UNROLL_HOT
#define LB_UNROLL_HOT2
#define LB_NI inline
#define LB_SHARED(p) ((void)0)
#else
#define LB_UNROLL1 _Pragma("unroll 1")
#ifdef LB_COMPACT  // the includer trades dynamic instructions for instruction footprint
#define LB_UNROLL_HOT _Pragma("unroll 1")
#define LB_UNROLL_HOT2 _Pragma("unroll 1")
#else
#define LB_UNROLL_HOT _Pragma("unroll 2")   // hot inner loops: loads of 2 iterations in flight
#define LB_UNROLL_HOT2 _Pragma("unroll 2")
#endif
#define LB_NI __device__ __noinline__
#define LB_SHARED(p) __builtin_assume(__isShared(p))
#endif
#if LB_VARIANT == 0
#define LB_FN inline
#else
#define LB_FN __device__ inline
#endif
#if LB_VARIANT == 1
#define LB_WARP 1
#define LB_LANE ((int)(threadIdx.x & 31))
#define LB_NL 32
#define LB_SYNC() __syncwarp()
#else
#define LB_WARP 0
#define LB_LANE 0
#define LB_NL 1
#define LB_SYNC() ((void)0)
#endif
#define LB_FOR(i, n) LB_UNROLL1 for (int i = LB_LANE; i < (n); i += LB_NL)

// The history size m (SciPy's `maxcor`).  The includer may define LB_MCONST: the device build
// is compiled a second time with m = 10 (the default, which the reference never changes) as a
// compile-time constant, so that the i*m + k / k*ldn + ... index arithmetic of every small dense
// kernel below folds into shifts and immediates (IMAD + IMAD.IADD were 13 % of the stepper's
// instructions with a run-time m).
#ifdef LB_MCONST
#define LB_M(x) (LB_MCONST)
#else
#define LB_M(x) (x)
#endif

typedef double *LbDP;
typedef int *LbIP;

// ld = L D^-1 lives in rows m..2m-1, columns 0..m-1 of wn (leading dimension 2m): the (2,1) block
// of the K matrix, which nothing else touches -- formt's scratch is rows 0..m-1, formk and the
// factorisations only ever read or write the upper triangle.
#define LB_LDL(m) (2 * (m))

// ------------------------------------------------------------------ warp collectives
// (the butterflies are fully unrolled: rolled, the compiler shuffles the register pair of the
// double through a chain of XOR swaps -- 13 instructions per step instead of 3; the helpers are
// real functions, so there is one copy of each)
LB_NI double lb_sum(double v) {
#if LB_WARP
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}
LB_NI double lb_max(double v) {
#if LB_WARP
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
  return v;
}
LB_NI int lb_isum(int v) {
#if LB_WARP
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}
LB_FN int lb_any(int p) {
#if LB_WARP
  return __any_sync(0xffffffffu, p);
#else
  return p;
#endif
}
// minimum value and the SMALLEST index attaining it
struct LbArgMin { double v; int idx; };
LB_NI LbArgMin lb_argmin(double v, int idx) {
#if LB_WARP
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
#endif
  LbArgMin r;
  r.v = v; r.idx = idx;
  return r;
}

// scratch + state views for one start (all in fast memory while a step runs)
//
// Limited-memory matrices.  W = [Y S] holds the correction pairs; SY = S'Y, SS = S'S and
// YY = Y'Y are kept as FULL m x m matrices (the original keeps the lower triangle of SY and the
// upper triangle of SS only): with the totals at hand, a sum over the free variables is
// total - sum over the active ones, so formk only ever walks the SMALLER of the two index sets.
// The middle-matrix product of bmv is applied through precomputed operators instead of two
// triangular solves per call: tinv = T^-1 (T = theta*SS + L D^-1 L', formt) and
// ld = L D^-1 (strictly lower) with 1/D on its diagonal, so that
//   p2 = tinv (v2 + ld v1),   p1 = -D^-1 v1 + ld' p2
// is three short matrix-vector products with no serial chain.
struct LbWork {
  LbDP x, g, z, r, d, t, xp;          // [n]
  LbDP W;                             // [n][LDW]: wy cols 0..m-1, ws cols m..2m-1 (logical order)
  LbDP sy, ss, yy, tinv;              // [m][m] persisted (contiguous, in this order)
  double *ld;                         // [m][LB_LDL(m)] derived from sy (lb_prep_ld); inside wn
  double *wn;                         // [2m][2m] upper triangle (also scratch of formt)
  double *rd;                         // [2m] reciprocal diagonal of the last Cholesky factors
  double *p, *c, *wbp, *v, *q;        // [2m]
  LbIP iwhere;                        // [n]
  LbIP index;                         // [n] free variables first (count nfree), active after
                                      //     (warp variant only; the serial variants mask)
  const int *ftab;                    // formk output map (lb_formk_code), one per CTA, warp variant only
  int changed;                        // out: the request lb_advance just posted differs from the point
                                      //      evaluated last (SciPy memoises on x: a repeat is not an nfev)
};

LB_HD size_t lb_work_doubles(int n, int m) {
  return (size_t)LB_PERSIST_DOUBLES(n, m) + 2 * LB_NV(n) + 4 * m * m + 12 * m;
}
LB_HD size_t lb_work_ints(int n) { return 2 * (size_t)n; }

LB_FN void lb_carve(LbWork &w, double *dbase, int *ibase, int n, int m) {
  m = LB_M(m);
  // persisted block first (see LB_PERSIST_DOUBLES), scratch after it
  const int nv = LB_NV(n);
  double *q = dbase + LB_SCAL_DOUBLES;
  w.t = q; q += nv; w.r = q; q += nv; w.d = q; q += nv; w.z = q; q += nv;
  w.W = q; q += LB_NW(n, m);
  w.sy = q; q += m * m; w.ss = q; q += m * m; w.yy = q; q += m * m; w.tinv = q; q += m * m;
  q = dbase + LB_PERSIST_DOUBLES(n, m);  // scratch starts behind the (padded) persisted block
  w.x = q; q += nv; w.g = q; q += nv;
  w.xp = w.t;  // the Cauchy breakpoints (t) are dead by the time subsm saves the Cauchy point
  w.wn = q; q += 4 * m * m;
  w.ld = w.wn + m * LB_LDL(m);
  w.rd = q; q += 2 * m;
  w.p = q; q += 2 * m; w.c = q; q += 2 * m; w.wbp = q; q += 2 * m; w.v = q; q += 2 * m;
  w.q = q; q += 2 * m;
  w.iwhere = ibase; w.index = ibase + n;
  w.ftab = nullptr;
}

// ------------------------------------------------------------------ small dense kernels
// Cholesky A = R'R in place, R in the upper triangle (LINPACK dpofa); rd[k] = 1/R[k][k].
// 0 ok, else k+1.  Right-looking: after column k is scaled, lane j owns column k+1+j of the
// trailing block and walks its rows, so there is no index arithmetic in the inner loop.
#if LB_WARP
// 1/sqrt(a) for a Cholesky pivot a > 0 (checked by the caller, normal range): the hardware seed
// rsqrt.approx.ftz.f64 (relative error 2^-22) and two Newton steps y += y (1/2 - (a/2) y^2), i.e.
// 2^-43, then fp64 rounding level -- 9 instructions.  CUDA's rsqrt() spends ~35 on the same
// value plus the special cases (zero, denormal, inf, NaN) that cannot occur here; 30 pivots per
// iteration made it 4 % of the stepper's instructions (profiles/r01_notes.md).
__device__ __forceinline__ double lb_rsqrt(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double h = 0.5 * a;
  double e = fma(-h, y * y, 0.5);
  y = fma(y, e, y);
  e = fma(-h, y * y, 0.5);
  return fma(y, e, y);
}
#endif
LB_NI int lb_chol(double *A, int ld, int n, double *rd) {
  LB_SHARED(A); LB_SHARED(rd);
#if LB_WARP
  // n <= LB_MMAX.  Lane j keeps column j of the upper triangle in registers; the scaled pivot
  // row travels by shuffle, so the only shared-memory traffic is one load and one store of
  // the matrix.  The matrix is padded with the identity up to LB_MMAX, which makes every
  // loop bound a compile-time constant (no branches around the shuffles); below the diagonal
  // the registers fill with the mirror image, which nothing reads.  Same operations in the
  // same order as the serial form below.
  LB_SYNC();
  const int j = LB_LANE;
  double a[LB_MMAX];
#pragma unroll
  for (int i = 0; i < LB_MMAX; ++i) {
    a[i] = i == j ? 1.0 : 0.0;
    if (j < n && i <= j) a[i] = A[i * ld + j];
  }
#pragma unroll
  for (int k = 0; k < LB_MMAX; ++k) {
    const double akk = __shfl_sync(0xffffffffu, a[k], k);
    if (!(akk > 0.0)) return k + 1;
    // the factor's diagonal is only ever used through its reciprocal (rd): every consumer on
    // the device (lb_trsl_*, formt's inverse, formk's forward substitution) multiplies by rd[k],
    // so sqrt(akk) itself is not formed; the diagonal slot keeps akk * rinv
    const double rinv = lb_rsqrt(akk);
    const double akj = a[k] * rinv;
    a[k] = akj;
    if (j == k && k < n) rd[k] = rinv;
#pragma unroll
    for (int i = k + 1; i < LB_MMAX; ++i) a[i] -= __shfl_sync(0xffffffffu, akj, i) * akj;  // lane i holds R[k][i]
  }
#pragma unroll
  for (int i = 0; i < LB_MMAX; ++i)
    if (j < n && i <= j) A[i * ld + j] = a[i];
  LB_SYNC();
  return 0;
#else
  for (int k = 0; k < n; ++k) {
    const double akk = A[k * ld + k];
    if (!(akk > 0.0)) return k + 1;
    const double rkk = sqrt(akk);
    const double rinv = 1.0 / rkk;
    for (int j = k + 1; j < n; ++j) A[k * ld + j] *= rinv;
    A[k * ld + k] = rkk;
    rd[k] = rinv;
    for (int j = k + 1; j < n; ++j) {
      const double akj = A[k * ld + j];
      for (int i = k + 1; i <= j; ++i) A[i * ld + j] -= A[k * ld + i] * akj;
    }
  }
  return 0;
#endif
}

// solve R' x = b (R upper; dtrsl job 11), b overwritten; n <= 32 on the device, where lane i
// keeps b[i] in a register and the pivot travels by shuffle (no shared-memory round trip in
// the serial chain).  rd = reciprocal diagonal from lb_chol.
LB_NI void lb_trsl_t(const double *R, int ld, int n, const double *rd, double *b) {
  LB_SHARED(R); LB_SHARED(rd); LB_SHARED(b);
  LB_SYNC();
#if LB_WARP
  const int i = LB_LANE;
  double bi = i < n ? b[i] : 0.0;
  LB_UNROLL1
  for (int k = 0; k < n; ++k) {
    const double rki = (i > k && i < n) ? R[k * ld + i] : 0.0;
    const double xk = __shfl_sync(0xffffffffu, bi, k) * rd[k];
    if (i == k) bi = xk;
    bi -= rki * xk;
  }
  if (i < n) b[i] = bi;
#else
  LB_UNROLL1
  for (int k = 0; k < n; ++k) {
    const double xk = b[k] * rd[k];
    b[k] = xk;
    LB_UNROLL1
    for (int i = k + 1; i < n; ++i) b[i] -= R[k * ld + i] * xk;
  }
#endif
  LB_SYNC();
}

// solve R x = b (R upper; dtrsl job 01), b overwritten
LB_NI void lb_trsl_n(const double *R, int ld, int n, const double *rd, double *b) {
  LB_SHARED(R); LB_SHARED(rd); LB_SHARED(b);
  LB_SYNC();
#if LB_WARP
  const int i = LB_LANE;
  double bi = i < n ? b[i] : 0.0;
  LB_UNROLL1
  for (int k = n - 1; k >= 0; --k) {
    const double rik = i < k ? R[i * ld + k] : 0.0;
    const double xk = __shfl_sync(0xffffffffu, bi, k) * rd[k];
    if (i == k) bi = xk;
    bi -= rik * xk;
  }
  if (i < n) b[i] = bi;
#else
  LB_UNROLL1
  for (int k = n - 1; k >= 0; --k) {
    const double xk = b[k] * rd[k];
    b[k] = xk;
    LB_UNROLL1
    for (int i = 0; i < k; ++i) b[i] -= R[i * ld + k] * xk;
  }
#endif
  LB_SYNC();
}

// Flat work sharing for the m x m triangles: pair p < m(m+1)/2 -> (i, j) with i >= j comes from
// the CTA's table (the family-0 entries of lb_formk_code), so that all 32 lanes share the 55
// pairs of a triangle instead of <= 10 lanes walking its columns.
#define LB_PAIR_HI(cd) (((cd) >> 2) & 63)
#define LB_PAIR_LO(cd) (((cd) >> 8) & 63)

// ld = L D^-1 below the diagonal, 1/D on it (L, D = strictly lower part / diagonal of SY)
LB_FN void lb_prep_ld(LbWork &w, int m, int col) {
  m = LB_M(m);
  LB_SYNC();
  LB_FOR(k, col) w.ld[k * LB_LDL(m) + k] = 1.0 / w.sy[k * m + k];
  LB_SYNC();
#if LB_WARP
  LB_FOR(p, m * (m + 1) / 2) {
    const int cd = w.ftab[p];
    const int i = LB_PAIR_HI(cd), k = LB_PAIR_LO(cd);
    if (i < col && k < i) w.ld[i * LB_LDL(m) + k] = w.sy[i * m + k] * w.ld[k * LB_LDL(m) + k];
  }
#else
  for (int i = 1; i < col; ++i)
    for (int k = 0; k < i; ++k) w.ld[i * LB_LDL(m) + k] = w.sy[i * m + k] * w.ld[k * LB_LDL(m) + k];
#endif
  LB_SYNC();
}

// p = M v with M the inverse 2col x 2col middle matrix of the compact L-BFGS formula (bmv),
// through the precomputed operators tinv and ld (see LbWork).  v and p must not alias.
LB_NI void lb_bmv(const double *ld, const double *tinv, double *q, int m, int col,
                  const double *v, double *p) {
  m = LB_M(m);
  LB_SHARED(ld); LB_SHARED(tinv); LB_SHARED(q); LB_SHARED(v); LB_SHARED(p);
  if (col == 0) return;
  LB_SYNC();
  LB_FOR(i, col) {
    double a = v[col + i];
    LB_UNROLL_HOT
    for (int k = 0; k < i; ++k) a += ld[i * LB_LDL(m) + k] * v[k];
    q[i] = a;
  }
  LB_SYNC();
  LB_FOR(i, col) {
    double a = 0.0;
    LB_UNROLL_HOT
    for (int j = 0; j < col; ++j) a += tinv[i * m + j] * q[j];
    p[col + i] = a;
  }
  LB_SYNC();
  LB_FOR(i, col) {
    double a = -ld[i * LB_LDL(m) + i] * v[i];
    LB_UNROLL_HOT
    for (int k = i + 1; k < col; ++k) a += ld[k * LB_LDL(m) + i] * p[col + k];
    p[i] = a;
  }
  LB_SYNC();
}

// T = theta*SS + L D^-1 L' (formt), Cholesky-factored in scratch (w.wn), then inverted into
// w.tinv: every later bmv is a plain matrix-vector product.  Returns nonzero when T is not
// positive definite (the caller refreshes the memory, like the original).
LB_FN int lb_formt(LbWork &w, int m, int col, double theta) {
  m = LB_M(m);
  double *T = w.wn;             // [col][m] upper triangle -> R
  double *Ri = w.wn + m * m;    // [col][m] upper triangle: R^-1
  lb_prep_ld(w, m, col);
#if LB_WARP
  const int npair = m * (m + 1) / 2;
  LB_FOR(p, npair) {
    const int cd = w.ftab[p];
    const int j = LB_PAIR_HI(cd), i = LB_PAIR_LO(cd);  // i <= j
    if (j < col) {
      double a = theta * w.ss[i * m + j];
      LB_UNROLL_HOT
      for (int k = 0; k < i; ++k) a += w.ld[i * LB_LDL(m) + k] * w.sy[j * m + k];
      T[i * m + j] = a;
    }
  }
  if (lb_chol(T, m, col, w.rd)) return -3;
  // R^-1: lane j solves R x = e_j by back substitution with its column in registers; the
  // entries of R are the same for every lane (broadcast loads)
  {
    const int j = LB_LANE;
    double x[LB_MMAX];
#pragma unroll
    for (int k = LB_MMAX - 1; k >= 0; --k) {
      double a = k == j ? 1.0 : 0.0;
#pragma unroll
      for (int i = k + 1; i < LB_MMAX; ++i) {
        const double t = (i < col) ? T[k * m + i] : 0.0;
        a -= t * x[i];  // x[i] == 0 for i > j
      }
      x[k] = (k <= j && k < col) ? a * w.rd[k] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < LB_MMAX; ++k)
      if (j < col && k <= j) Ri[k * m + j] = x[k];
  }
  LB_SYNC();
  // T^-1 = R^-1 R^-T (symmetric, stored full)
  LB_FOR(p, npair) {
    const int cd = w.ftab[p];
    const int j = LB_PAIR_HI(cd), i = LB_PAIR_LO(cd);  // i <= j
    if (j < col) {
      double a = 0.0;
      LB_UNROLL_HOT
      for (int k = j; k < col; ++k) a += Ri[i * m + k] * Ri[j * m + k];
      w.tinv[i * m + j] = a;
      w.tinv[j * m + i] = a;
    }
  }
  LB_SYNC();
  return 0;
#else
  for (int i = 0; i < col; ++i)
    for (int j = i; j < col; ++j) {
      double a = theta * w.ss[i * m + j];
      for (int k = 0; k < i; ++k) a += w.ld[i * LB_LDL(m) + k] * w.sy[j * m + k];
      T[i * m + j] = a;
    }
  if (lb_chol(T, m, col, w.rd)) return -3;
  // R^-1 column by column (back substitution of R x = e_j; rows > j are zero)
  for (int j = 0; j < col; ++j)
    for (int k = j; k >= 0; --k) {
      double a = k == j ? 1.0 : 0.0;
      for (int i = k + 1; i <= j; ++i) a -= T[k * m + i] * Ri[i * m + j];
      Ri[k * m + j] = a * w.rd[k];
    }
  // T^-1 = R^-1 R^-T (symmetric, stored full)
  for (int i = 0; i < col; ++i)
    for (int j = i; j < col; ++j) {
      double a = 0.0;
      for (int k = j; k < col; ++k) a += Ri[i * m + k] * Ri[j * m + k];
      w.tinv[i * m + j] = a;
      w.tinv[j * m + i] = a;
    }
  return 0;
#endif
}

// ------------------------------------------------------------------ projected gradient norm
LB_NI double lb_projgr(int n, const int *nbd, const double *lo, const double *hi, const double *x,
                       const double *g) {
  LB_SHARED(x); LB_SHARED(g);
  double s = 0.0;
  LB_FOR(i, n) {
    double gi = g[i];
    const int nb = nbd[i];
    if (nb != 0) {
      if (gi < 0.0) {
        if (nb >= 2) gi = fmax(x[i] - hi[i], gi);
      } else {
        if (nb <= 2) gi = fmin(x[i] - lo[i], gi);
      }
    }
    s = fmax(s, fabs(gi));
  }
  return lb_max(s);
}

// ------------------------------------------------------------------ generalized Cauchy point
// Breakpoints live in w.t (free until the line search starts), the Cauchy direction in w.d,
// the Cauchy point in w.z; w.c receives W'(xcp - x).  Instead of the heap of the original
// (hpsolb) the next breakpoint is a warp arg-min over the remaining ones.
LB_FN int lb_cauchy(const LbParams &P, LbWork &w, const LbScal &s, int &nseg_out) {
  const int n = P.n, m = LB_M(P.m), col = s.col, col2 = 2 * col, ldw = LB_LDW(m);
  const double theta = s.theta;
  LbDP tb = w.t, d = w.d, xcp = w.z;
  nseg_out = 0;
  LB_SYNC();
  if (s.sbgnrm <= 0.0) {
    LB_FOR(i, n) xcp[i] = w.x[i];
    LB_SYNC();
    return 0;
  }
  double f1 = 0.0;
  int nbreak = 0, nmove = 0, unb = 0;
  LB_FOR(i, n) {
    const double neggi = -w.g[i];
    int iw = w.iwhere[i];
    const int nb = P.nbd[i];
    double tl = 0.0, tu = 0.0;
    if (iw != 3 && iw != -1) {
      if (nb <= 2) tl = w.x[i] - P.lo[i];
      if (nb >= 2) tu = P.hi[i] - w.x[i];
      const bool xlower = nb <= 2 && tl <= 0.0;
      const bool xupper = nb >= 2 && tu <= 0.0;
      iw = 0;
      if (xlower) { if (neggi <= 0.0) iw = 1; }
      else if (xupper) { if (neggi >= 0.0) iw = 2; }
      else { if (fabs(neggi) <= 0.0) iw = -3; }
    }
    double di = 0.0, tbi = LB_INF;
    if (iw == 0 || iw == -1) {
      di = neggi;
      f1 -= neggi * neggi;
      ++nmove;
      if (nb <= 2 && nb != 0 && neggi < 0.0) { tbi = tl / (-neggi); ++nbreak; }
      else if (nb >= 2 && neggi > 0.0) { tbi = tu / neggi; ++nbreak; }
      else if (fabs(neggi) > 0.0) unb = 1;
    }
    w.iwhere[i] = iw;
    d[i] = di;
    tb[i] = tbi;
    xcp[i] = w.x[i];
  }
  f1 = lb_sum(f1);
  nbreak = lb_isum(nbreak);
  nmove = lb_isum(nmove);
  const bool bnded = !lb_any(unb);
  LB_SYNC();
  // p = W'd (ws half scaled by theta), c = 0
  LB_UNROLL1
  for (int j = LB_LANE; j < col2; j += LB_NL) {
    const LbDP wc = w.W + (j < col ? j : m + (j - col));
    double a = 0.0;
    LB_UNROLL_HOT
    for (int i = 0; i < n; ++i) a += wc[i * ldw] * d[i];
    w.p[j] = j < col ? a : theta * a;
    w.c[j] = 0.0;
  }
  LB_SYNC();
  if (nbreak == 0 && nmove == 0) return 0;  // d == 0: xcp = x

  double f2 = -theta * f1;
  const double f2_org = f2;
  if (col > 0) {
    lb_bmv(w.ld, w.tinv, w.q, m, col, w.p, w.v);
    double a = 0.0;
    LB_UNROLL1
    for (int j = LB_LANE; j < col2; j += LB_NL) a += w.v[j] * w.p[j];
    f2 -= lb_sum(a);
  }
  double dtm = -f1 / f2, tsum = 0.0;
  int nseg = 1;
  bool all_fixed = false;
  if (nbreak > 0) {
    int nleft = nbreak;
    double tj = 0.0;
    LB_UNROLL1
    for (;;) {
      const double tj0 = tj;
      // smallest remaining breakpoint
      double bv = LB_INF;
      int bi = 0x7fffffff;
      LB_FOR(i, n) if (tb[i] < bv) { bv = tb[i]; bi = i; }
      const LbArgMin am = lb_argmin(bv, bi);
      tj = am.v;
      const int ibp = am.idx;
      const double dt = tj - tj0;
      if (dtm < dt) break;  // minimiser inside this segment
      tsum += dt;
      --nleft;
      const double dibp = d[ibp];
      double zibp;
      const double xb = dibp > 0.0 ? P.hi[ibp] : P.lo[ibp];
      zibp = xb - w.x[ibp];
      LB_SYNC();
      if (LB_LANE == 0) {
        d[ibp] = 0.0;
        tb[ibp] = LB_INF;
        xcp[ibp] = xb;
        w.iwhere[ibp] = dibp > 0.0 ? 2 : 1;
      }
      LB_SYNC();
      if (nleft == 0 && nbreak == n) {  // every variable is fixed
        dtm = dt;
        all_fixed = true;
        break;
      }
      ++nseg;
      const double dibp2 = dibp * dibp;
      f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
      f2 = f2 - theta * dibp2;
      if (col > 0) {
        LB_UNROLL1
        for (int j = LB_LANE; j < col2; j += LB_NL) {
          w.c[j] += dt * w.p[j];
          w.wbp[j] = j < col ? w.W[ibp * ldw + j] : theta * w.W[ibp * ldw + m + (j - col)];
        }
        lb_bmv(w.ld, w.tinv, w.q, m, col, w.wbp, w.v);
        double wmc = 0.0, wmp = 0.0, wmw = 0.0;
        LB_UNROLL1
        for (int j = LB_LANE; j < col2; j += LB_NL) {
          const double vj = w.v[j];
          wmc += w.c[j] * vj;
          wmp += w.p[j] * vj;
          wmw += w.wbp[j] * vj;
          w.p[j] -= dibp * w.wbp[j];
        }
        wmc = lb_sum(wmc); wmp = lb_sum(wmp); wmw = lb_sum(wmw);
        LB_SYNC();
        f1 += dibp * wmc;
        f2 += 2.0 * dibp * wmp - dibp2 * wmw;
      }
      f2 = fmax(LB_EPSMCH * f2_org, f2);
      if (nleft > 0) { dtm = -f1 / f2; continue; }
      if (bnded) { f1 = 0.0; f2 = 0.0; dtm = 0.0; }
      else dtm = -f1 / f2;
      break;
    }
  }
  if (!all_fixed) {
    if (dtm <= 0.0) dtm = 0.0;
    tsum += dtm;
    LB_FOR(i, n) xcp[i] += tsum * d[i];
  }
  if (col > 0)
    LB_UNROLL1
    for (int j = LB_LANE; j < col2; j += LB_NL) w.c[j] += dtm * w.p[j];
  LB_SYNC();
  nseg_out = nseg;
  return 0;
}

// ------------------------------------------------------------------ free / active sets at the GCP
// index[0..nfree) = free variables (iwhere <= 0) in ascending order, the rest active.
LB_FN int lb_freev(const LbParams &P, LbWork &w) {
  const int n = P.n;
  LB_SYNC();
#if LB_WARP
  int base_f = 0, base_a = 0;
  // ordered compaction, 32 variables at a time
  int nfree_total = 0;
  {
    int c = 0;
    LB_FOR(i, n) c += (w.iwhere[i] <= 0);
    nfree_total = lb_isum(c);
  }
  LB_UNROLL1
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + LB_LANE;
    const bool valid = i < n;
    const bool fr = valid && w.iwhere[i] <= 0;
    const unsigned mf = __ballot_sync(0xffffffffu, fr);
    const unsigned ma = __ballot_sync(0xffffffffu, valid && !fr);
    const unsigned lt = (1u << LB_LANE) - 1u;
    if (fr) w.index[base_f + __popc(mf & lt)] = i;
    else if (valid) w.index[nfree_total + base_a + __popc(ma & lt)] = i;
    base_f += __popc(mf);
    base_a += __popc(ma);
  }
  LB_SYNC();
  return nfree_total;
#else
  // serial variants sweep all n variables with the iwhere mask instead of an index list
  int nfree = 0;
  LB_UNROLL1
  for (int i = 0; i < n; ++i) nfree += (w.iwhere[i] <= 0);
  return nfree;
#endif
}

// ------------------------------------------------------------------ K = LEL' factorisation (formk)
// Built from scratch each time it is needed (the original updates it incrementally; the
// matrix is the same).  wn: upper triangle of the 2col x 2col matrix, leading dim 2m:
//   (1,1)  Y'ZZ'Y / theta + D        Z = free variables, A = active variables
//   (2,2)  theta * S'AA'S
//   (1,2)  R_z (s_i' Z Z' y_j, j >= i)  and  -L_a (s_i' A A' y_j, j < i)
// The three Gram sums  Gyy = Y_Q'Y_Q, Gss = S_Q'S_Q, Gsy = S_Q'Y_Q  are taken over ONE index
// set Q, the smaller of Z and A; the sums over the other set are  total - G  with the totals
// YY, SS, SY kept up to date by matupd.  Output o of the 2*tri+col^2 sums is owned by lane
// o % 32, which keeps it in a register while the warp walks the rows of Q.
#define LB_FORMK_ACC 7  // ceil((2*55 + 100) / 32) for m = 10
// The map output -> (family, i, j, column pair) is enumerated over the FULL m x m shapes
// (2*55 + 100 = 210 outputs for m = 10) so that it does not depend on col: it is tabulated
// once per CTA (lb_formk_table, LB_FORMK_ACC*32 ints in shared memory) and the entries with
// i >= col or j >= col are simply not written out.
//   code = ty | i << 2 | j << 8 | ca << 14 | cb << 22
LB_FN int lb_formk_code(int o, int m) {
  const int ntri = m * (m + 1) / 2;
  int ty, i, j, ca, cb;
  if (o >= 2 * ntri + m * m) return 0;
  if (o < 2 * ntri) {
    ty = o >= ntri;
    const int q = ty ? o - ntri : o;
    i = 0;
    while ((i + 1) * (i + 2) / 2 <= q) ++i;  // q = i*(i+1)/2 + j, i >= j
    j = q - i * (i + 1) / 2;
    ca = ty ? m + i : i;
    cb = ty ? m + j : j;
  } else {
    ty = 2;
    const int q = o - 2 * ntri;
    i = q / m;      // s index
    j = q - i * m;  // y index
    ca = m + i;
    cb = j;
  }
  return ty | i << 2 | j << 8 | ca << 14 | cb << 22;
}

#if !LB_WARP
// Serial Gram sweeps over the FREE variables (iwhere <= 0), one output family per sweep so
// that the accumulators live in registers: every loop below has compile-time bounds and the
// columns >= col contribute zeros.
//   lb_gram_sym : sum_i v_a(i) v_b(i), a >= b, for the column block starting at `coff`
//   lb_gram_sy  : sum_i s_a(i) y_b(i) for a in [A0, A0 + LB_MMAX/2)
LB_FN void lb_gram_sym(const LbParams &P, const LbWork &w, int col, int coff, double *out) {
  const int n = P.n, ldw = LB_LDW(LB_M(P.m));
  double acc[LB_MMAX * (LB_MMAX + 1) / 2];
#pragma unroll
  for (int e = 0; e < LB_MMAX * (LB_MMAX + 1) / 2; ++e) acc[e] = 0.0;
  LB_UNROLL1
  for (int i = 0; i < n; ++i) {
    if (w.iwhere[i] > 0) continue;
    const LbDP row = w.W + (i * ldw + coff);
    double v[LB_MMAX];
#pragma unroll
    for (int a = 0; a < LB_MMAX; ++a) v[a] = a < col ? row[a] : 0.0;
#pragma unroll
    for (int a = 0; a < LB_MMAX; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) acc[a * (a + 1) / 2 + b] += v[a] * v[b];
  }
#pragma unroll
  for (int e = 0; e < LB_MMAX * (LB_MMAX + 1) / 2; ++e) out[e] = acc[e];
}
template <int A0>
LB_FN void lb_gram_sy(const LbParams &P, const LbWork &w, int col, double *out) {
  const int n = P.n, m = LB_M(P.m), ldw = LB_LDW(LB_M(P.m));
  constexpr int NA = LB_MMAX / 2;
  double acc[NA * LB_MMAX];
#pragma unroll
  for (int e = 0; e < NA * LB_MMAX; ++e) acc[e] = 0.0;
  LB_UNROLL1
  for (int i = 0; i < n; ++i) {
    if (w.iwhere[i] > 0) continue;
    const LbDP row = w.W + i * ldw;
    double sv[NA], yv[LB_MMAX];
#pragma unroll
    for (int a = 0; a < NA; ++a) sv[a] = A0 + a < col ? row[m + A0 + a] : 0.0;
#pragma unroll
    for (int b = 0; b < LB_MMAX; ++b) yv[b] = b < col ? row[b] : 0.0;
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int b = 0; b < LB_MMAX; ++b) acc[a * LB_MMAX + b] += sv[a] * yv[b];
  }
#pragma unroll
  for (int e = 0; e < NA * LB_MMAX; ++e) out[(A0 + e / LB_MMAX) * LB_MMAX + e % LB_MMAX] = acc[e];
}
#endif

LB_FN int lb_formk(const LbParams &P, LbWork &w, const LbScal &s, int nfree) {
  const int n = P.n, m = LB_M(P.m), col = s.col, ldw = LB_LDW(m), ldn = 2 * m;
  const double theta = s.theta;
  LB_SYNC();
#if LB_WARP
  // Gram matrix of the compacted columns V = [Y(:, 0..col) S(:, 0..col)] over the SMALLER of
  // the free / active index sets (the other set follows from the kept totals SY, SS, YY), on the
  // FP64 tensor cores: G = V'V as 8x8 tiles of mma.sync.m8n8k4 (DMMA), four selected rows of W per
  // step.  For tile T a lane holds V[k0 + lane%4][8T + lane/4], which is the A fragment of tile
  // row T and the B fragment of tile column T alike, so three shared-memory loads feed the six
  // tiles of the upper triangle.  Entry (r, c), r <= c, of G belongs at wn[r][c] -- the position
  // of the same entry of the 2col x 2col middle matrix -- after the family's affine map:
  //   r, c <  col : Y'ZZ'Y / theta + D            (j = r, i = c)
  //   r, c >= col : theta * S'AA'S                (j = r - col, i = c - col)
  //   r < col <= c: L_a' / R_z' (sy - g or g, sign by j >= i; y index j = r, s index i = c - col)
  // (The first device version gave every lane 7 of the 210 outputs and walked the rows with two
  // loads and a DFMA per output and row: 15 % of the stepper's instructions.)
  const LbDP W = w.W;
  const bool over_free = nfree <= n - nfree;
  const LbIP ind = over_free ? w.index : w.index + nfree;
  const int nq = over_free ? nfree : n - nfree;
  const int twoc = 2 * col;
  const int kq = LB_LANE & 3, rq = LB_LANE >> 2;
  int coff[3];
#pragma unroll
  for (int T = 0; T < 3; ++T) {
    const int r = 8 * T + rq;
    coff[T] = r >= twoc ? -1 : (r < col ? r : m + (r - col));
  }
  double acc[12];
#pragma unroll
  for (int t = 0; t < 12; ++t) acc[t] = 0.0;
#define LB_DMMA(t, a, b)                                                                          \
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"   \
               : "+d"(acc[2 * (t)]), "+d"(acc[2 * (t) + 1]) : "d"(a), "d"(b))
  LB_UNROLL1
  for (int k0 = 0; k0 < nq; k0 += 4) {
    const int kk = k0 + kq;
    const bool kv = kk < nq;
    const LbDP row = W + (kv ? ind[kk] : 0) * ldw;
    const double v0 = (kv && coff[0] >= 0) ? row[coff[0] < 0 ? 0 : coff[0]] : 0.0;
    const double v1 = (kv && coff[1] >= 0) ? row[coff[1] < 0 ? 0 : coff[1]] : 0.0;
    const double v2 = (kv && coff[2] >= 0) ? row[coff[2] < 0 ? 0 : coff[2]] : 0.0;
    LB_DMMA(0, v0, v0);
    if (twoc > 8) {
      LB_DMMA(1, v0, v1);
      LB_DMMA(3, v1, v1);
    }
    if (twoc > 16) {
      LB_DMMA(2, v0, v2);
      LB_DMMA(4, v1, v2);
      LB_DMMA(5, v2, v2);
    }
  }
#undef LB_DMMA
  // write-out as a ROLLED loop (one copy of the code): the accumulators take a detour through
  // a dynamically indexed (local-memory) array.  Tile t = (R, C) of the upper triangle:
  // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2); a lane holds (8R + lane/4, 8C + 2 (lane%4) + {0, 1}).
  double accm[12];
#pragma unroll
  for (int t = 0; t < 12; ++t) accm[t] = acc[t];
  LB_UNROLL1
  for (int q = 0; q < 12; ++q) {
    const int t = q >> 1;
    const int r = 8 * ((0x211000 >> (4 * t)) & 15) + rq;
    const int c = 8 * ((0x221210 >> (4 * t)) & 15) + 2 * kq + (q & 1);
    if (r > c || c >= twoc) continue;
    const double g = accm[q];
    double a;
    if (c < col) {          // Y'ZZ'Y / theta + D
      a = over_free ? g : w.yy[c * m + r] - g;
      a /= theta;
      if (r == c) a += w.sy[c * m + c];
    } else if (r >= col) {  // theta * S'AA'S
      const int i = c - col, j = r - col;
      a = (over_free ? w.ss[i * m + j] - g : g) * theta;
    } else {                // s column i with y column j
      const int i = c - col, j = r;
      if (j >= i) a = over_free ? g : w.sy[i * m + j] - g;
      else a = -(over_free ? w.sy[i * m + j] - g : g);
    }
    w.wn[r * ldn + c] = a;
  }
#else
  (void)nfree; (void)n; (void)ldw;
  {
    double G[LB_MMAX * LB_MMAX];  // one family at a time
    lb_gram_sym(P, w, col, 0, G);  // Y'ZZ'Y
    LB_UNROLL1
    for (int i = 0; i < col; ++i)
      LB_UNROLL1
      for (int j = 0; j <= i; ++j) {
        double a = G[i * (i + 1) / 2 + j] / theta;
        if (i == j) a += w.sy[i * m + i];
        w.wn[j * ldn + i] = a;
      }
    lb_gram_sym(P, w, col, m, G);  // S'ZZ'S; the active part is the total minus it
    LB_UNROLL1
    for (int i = 0; i < col; ++i)
      LB_UNROLL1
      for (int j = 0; j <= i; ++j)
        w.wn[(col + j) * ldn + (col + i)] = (w.ss[i * m + j] - G[i * (i + 1) / 2 + j]) * theta;
    lb_gram_sy<0>(P, w, col, G);   // S'ZZ'Y, s index i, y index j
    lb_gram_sy<LB_MMAX / 2>(P, w, col, G);
    LB_UNROLL1
    for (int i = 0; i < col; ++i)
      LB_UNROLL1
      for (int j = 0; j < col; ++j) {
        const double g = G[i * LB_MMAX + j];
        w.wn[j * ldn + (col + i)] = j >= i ? g : -(w.sy[i * m + j] - g);
      }
  }
#endif
  // Cholesky of the (1,1) block
  if (lb_chol(w.wn, ldn, col, w.rd)) return -1;
#if LB_WARP
  // (1,2) block := R11^-T (1,2): lane c carries right-hand side c through the forward
  // substitution in registers; the entries of R11 are broadcast loads
  {
    const int c = LB_LANE;
    double b[LB_MMAX];
#pragma unroll
    for (int k = 0; k < LB_MMAX; ++k) b[k] = (k < col && c < col) ? w.wn[k * ldn + col + c] : 0.0;
#pragma unroll
    for (int i = 0; i < LB_MMAX; ++i) {
#pragma unroll
      for (int k = 0; k < i; ++k) {
        const double t = (i < col) ? w.wn[k * ldn + i] : 0.0;
        b[i] -= t * b[k];
      }
      b[i] *= (i < col) ? w.rd[i] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < LB_MMAX; ++k)
      if (k < col && c < col) w.wn[k * ldn + col + c] = b[k];
  }
  LB_SYNC();
  // (2,2) block += (1,2)'(1,2), upper triangle, the 55 pairs shared by all lanes
  LB_FOR(p, m * (m + 1) / 2) {
    const int cd = w.ftab[p];
    const int js = LB_PAIR_HI(cd), is = LB_PAIR_LO(cd);  // is <= js
    if (js < col) {
      double a = 0.0;
      LB_UNROLL_HOT
      for (int k = 0; k < col; ++k) a += w.wn[k * ldn + col + is] * w.wn[k * ldn + col + js];
      w.wn[(col + is) * ldn + col + js] += a;
    }
  }
#else
  // (1,2) block := R11^-T (1,2): forward elimination over all col right-hand sides
  for (int k = 0; k < col; ++k) {
    for (int c = col; c < 2 * col; ++c) w.wn[k * ldn + c] *= w.rd[k];
    for (int c = col; c < 2 * col; ++c) {
      const double xk = w.wn[k * ldn + c];
      for (int i = k + 1; i < col; ++i) w.wn[i * ldn + c] -= w.wn[k * ldn + i] * xk;
    }
  }
  // (2,2) block += (1,2)'(1,2), upper triangle
  for (int is = col; is < 2 * col; ++is)
    for (int js = is; js < 2 * col; ++js) {
      double a = 0.0;
      for (int k = 0; k < col; ++k) a += w.wn[k * ldn + is] * w.wn[k * ldn + js];
      w.wn[is * ldn + js] += a;
    }
#endif
  if (lb_chol(w.wn + col * ldn + col, ldn, col, w.rd + col)) return -2;
  return 0;
}

// ------------------------------------------------------------------ reduced gradient (cmprlb)
// r[k] = -(B(xcp - x) + g)[k] for free k (0 elsewhere); uses c = W'(xcp - x) from cauchy.
LB_FN int lb_cmprlb(const LbParams &P, LbWork &w, const LbScal &s, int nfree) {
  const int n = P.n, m = LB_M(P.m), col = s.col, ldw = LB_LDW(m);
  const double theta = s.theta;
  LB_SYNC();
  if (!P.cnstnd && col > 0) {
    LB_FOR(i, n) w.r[i] = -w.g[i];
    LB_SYNC();
    return 0;
  }
  lb_bmv(w.ld, w.tinv, w.q, m, col, w.c, w.p);
  LB_FOR(i, n) {
    if (w.iwhere[i] <= 0) {
      double a = -theta * (w.z[i] - w.x[i]) - w.g[i];
      const LbDP row = w.W + i * ldw;
      LB_UNROLL_HOT
      for (int j = 0; j < col; ++j) a += row[j] * w.p[j] + row[m + j] * (theta * w.p[col + j]);
      w.r[i] = a;
    } else {
      w.r[i] = 0.0;
    }
  }
  LB_SYNC();
  (void)nfree;
  return 0;
}

// ------------------------------------------------------------------ subspace minimisation (subsm, v3.0)
// In: w.z = Cauchy point, w.r = reduced gradient.  Out: w.z = subspace minimiser.
LB_FN int lb_subsm(const LbParams &P, LbWork &w, const LbScal &s, int nfree) {
  const int n = P.n, m = LB_M(P.m), col = s.col, col2 = 2 * col, ldw = LB_LDW(m), ldn = 2 * m;
  const double theta = s.theta;
  if (nfree <= 0) return 0;
  double *wv = w.v;
  LB_SYNC();
  // wv = W'Z d
#if LB_WARP
  LB_UNROLL1
  for (int j = LB_LANE; j < col2; j += LB_NL) {
    const LbDP wc = w.W + (j < col ? j : m + (j - col));
    double a = 0.0;
    LB_UNROLL_HOT
    for (int k = 0; k < nfree; ++k) { const int i = w.index[k]; a += wc[i * ldw] * w.r[i]; }
    wv[j] = j < col ? a : theta * a;
  }
#else
  {
    double acc[2 * LB_MMAX];
#pragma unroll
    for (int a = 0; a < 2 * LB_MMAX; ++a) acc[a] = 0.0;
    LB_UNROLL1
    for (int i = 0; i < n; ++i) {
      if (w.iwhere[i] > 0) continue;
      const double ri = w.r[i];
      const LbDP row = w.W + i * ldw;
#pragma unroll
      for (int a = 0; a < LB_MMAX; ++a)
        if (a < col) { acc[a] += row[a] * ri; acc[LB_MMAX + a] += row[m + a] * ri; }
    }
#pragma unroll
    for (int a = 0; a < LB_MMAX; ++a)
      if (a < col) { wv[a] = acc[a]; wv[col + a] = theta * acc[LB_MMAX + a]; }
  }
#endif
  // wv := K^-1 wv  (the first col entries are pre-divided by theta for the loop below)
  lb_trsl_t(w.wn, ldn, col2, w.rd, wv);
  LB_UNROLL1
  for (int j = LB_LANE; j < col; j += LB_NL) wv[j] = -wv[j];
  lb_trsl_n(w.wn, ldn, col2, w.rd, wv);
  LB_UNROLL1
  for (int j = LB_LANE; j < col; j += LB_NL) wv[j] /= theta;
  LB_SYNC();
  // d = (1/theta) d + (1/theta^2) Z'W wv ; xp = xcp ; projected Newton point
  int iword = 0;
  LB_FOR(i, n) {
    const double zi = w.z[i];
    w.xp[i] = zi;
    if (w.iwhere[i] <= 0) {
      double dk = w.r[i];
      const LbDP row = w.W + i * ldw;
      LB_UNROLL_HOT
      for (int j = 0; j < col; ++j) dk += row[j] * wv[j] + row[m + j] * wv[col + j];
      dk *= 1.0 / theta;
      w.r[i] = dk;
      const int nb = P.nbd[i];
      double xk = zi;
      if (nb != 0) {
        if (nb == 1) {
          xk = fmax(P.lo[i], zi + dk);
          if (xk == P.lo[i]) iword = 1;
        } else if (nb == 2) {
          xk = fmin(P.hi[i], fmax(P.lo[i], zi + dk));
          if (xk == P.lo[i] || xk == P.hi[i]) iword = 1;
        } else {
          xk = fmin(P.hi[i], zi + dk);
          if (xk == P.hi[i]) iword = 1;
        }
      } else {
        xk = zi + dk;
      }
      w.z[i] = xk;
    }
  }
  iword = lb_any(iword);
  LB_SYNC();
  if (!iword) return 0;
  // sign of the directional derivative along the projected step
  double ddp = 0.0;
  LB_FOR(i, n) ddp += (w.z[i] - w.x[i]) * w.g[i];
  ddp = lb_sum(ddp);
  if (ddp > 0.0) {
    // fall back to the truncated (unprojected) Newton step from xcp
    double amin = LB_INF;
    int ibd = 0x7fffffff;
    LB_FOR(i, n) {
      w.z[i] = w.xp[i];
      if (w.iwhere[i] <= 0) {
        const double dk = w.r[i];
        const int nb = P.nbd[i];
        double ratio = LB_INF;
        if (nb != 0) {
          if (dk < 0.0 && nb <= 2) {
            const double t2 = P.lo[i] - w.xp[i];
            ratio = t2 >= 0.0 ? 0.0 : t2 / dk;
          } else if (dk > 0.0 && nb >= 2) {
            const double t2 = P.hi[i] - w.xp[i];
            ratio = t2 <= 0.0 ? 0.0 : t2 / dk;
          }
        }
        if (ratio < amin) { amin = ratio; ibd = i; }
      }
    }
    {
      const LbArgMin am = lb_argmin(amin, ibd);
      amin = am.v; ibd = am.idx;
    }
    double alpha = 1.0;
    LB_SYNC();
    if (amin < 1.0) {
      alpha = amin;
      const double dk = w.r[ibd];
      LB_SYNC();
      if (LB_LANE == 0) {
        if (dk > 0.0) { w.z[ibd] = P.hi[ibd]; w.r[ibd] = 0.0; }
        else if (dk < 0.0) { w.z[ibd] = P.lo[ibd]; w.r[ibd] = 0.0; }
      }
      LB_SYNC();
    }
    LB_FOR(i, n) if (w.iwhere[i] <= 0) w.z[i] += alpha * w.r[i];
    LB_SYNC();
  }
  return 0;
}

// ------------------------------------------------------------------ BFGS memory update (matupd)
// Columns are kept in logical order (0 = oldest): when the memory is full everything is
// shifted by one instead of rotating a head pointer.  On entry w.r = y (new), w.d = s (new),
// rr = y'y, dr = y's.  SY, SS and YY are maintained as full matrices (see LbWork).
LB_FN void lb_matupd(const LbParams &P, LbWork &w, LbScal &s, double rr, double dr) {
  const int n = P.n, m = LB_M(P.m), ldw = LB_LDW(m);
  LB_SYNC();
  if (s.iupdat <= m) {
    s.col = s.iupdat;
  } else {
    LB_FOR(i, n) {
      LbDP row = w.W + i * ldw;
      LB_UNROLL_HOT
      for (int j = 0; j < m - 1; ++j) { row[j] = row[j + 1]; row[m + j] = row[m + j + 1]; }
    }
    // new(i,j) = old(i+1,j+1) for the three (m-1) x (m-1) leading blocks of sy, ss, yy
    // (contiguous, m*m apart): lane q*c1 + j owns column j of matrix q and walks down it; a row
    // is read in the iteration before the one that overwrites it, hence one sync per row
    const int c1 = s.col - 1;
#if LB_WARP
    if (c1 > 0) {
      const int q = LB_LANE / c1, j = LB_LANE - q * c1;
      LbDP A = w.sy + q * m * m;
      LB_UNROLL1
      for (int i = 0; i < c1; ++i) {
        double tmp = 0.0;
        if (q < 3) tmp = A[(i + 1) * m + j + 1];
        LB_SYNC();
        if (q < 3) A[i * m + j] = tmp;
      }
    }
#else
    LbDP mats[3] = {w.sy, w.ss, w.yy};
    for (int q = 0; q < 3; ++q) {
      LbDP A = mats[q];
      for (int i = 0; i < c1; ++i)
        for (int jj = 0; jj < c1; ++jj) A[i * m + jj] = A[(i + 1) * m + jj + 1];
    }
#endif
  }
  const int col = s.col, last = col - 1;
  LB_FOR(i, n) {
    w.W[i * ldw + m + last] = w.d[i];
    w.W[i * ldw + last] = w.r[i];
  }
  s.theta = rr / dr;
  LB_SYNC();
  // products of the new pair with every older column j: lane j < last pairs y_j with (s,y),
  // lane last + j pairs s_j with (s,y)
  LB_UNROLL1
  for (int jl = LB_LANE; jl < 2 * last; jl += LB_NL) {
    const bool is_y = jl < last;
    const int j = is_y ? jl : jl - last;
    const LbDP colp = w.W + (is_y ? j : m + j);
    double a_s = 0.0, a_y = 0.0;
    LB_UNROLL_HOT
    for (int i = 0; i < n; ++i) {
      const double c = colp[i * ldw];
      a_s += w.d[i] * c;
      a_y += w.r[i] * c;
    }
    if (is_y) {  // y_j . s_new,  y_j . y_new
      w.sy[last * m + j] = a_s;
      w.yy[last * m + j] = a_y;
      w.yy[j * m + last] = a_y;
    } else {     // s_j . s_new,  s_j . y_new
      w.ss[j * m + last] = a_s;
      w.ss[last * m + j] = a_s;
      w.sy[j * m + last] = a_y;
    }
  }
  if (LB_LANE == 0) {
    w.ss[last * m + last] = s.stp == 1.0 ? s.dtd : s.stp * s.stp * s.dtd;
    w.sy[last * m + last] = dr;
    w.yy[last * m + last] = rr;
  }
  LB_SYNC();
}

// `Mem` supplies the limited-memory matrices lazily: mem.load() is called once, right before
// they are first needed (a step that only continues a line search never touches them), and
// mem.dirty() when they were modified.  mem.phase() marks the boundaries between the big
// pieces of a new iteration (memory update | Cauchy point | K factorisation | subspace
// minimisation | line-search set-up): the device kernel re-aligns the warps of a CTA there so
// that they fetch the same instructions at the same time.
struct LbNoMem {
  // ld (derived from sy) survives between steps when the state never leaves fast memory
  LB_FN bool ld_kept() { return true; }
  LB_FN void load() {}
  LB_FN void dirty() {}
  LB_FN void dirty_vec() {}
  LB_FN void phase() {}
};

//
// `stage` splits a step in two so that the cheap and the expensive halves can be scheduled
// separately: 0 = the whole step (what the product's kernel runs); 1 = LIGHT: consume f,g, run the line-search
// logic and, if the search goes on, post the next trial point (returns 1) -- if a new
// iteration has to be set up instead it records where to resume in s.resume and returns 2
// without touching the limited-memory matrices; 2 = HEAVY: resume there.
template <class Mem>
LB_FN int lb_advance(const LbParams &P, LbWork &w, LbScal &s, Mem &mem, int stage = 0) {
  const int n = P.n, m = LB_M(P.m);
  enum { ST_TESTS, ST_ITER, ST_REQUEST, ST_FAIL };
  int st;
  int restored = 0;  // w.x was reset to the previous iterate (it no longer is the point evaluated last)
  int ld_fresh = 0;  // w.ld matches w.sy (built in this call)

  if (s.phase == LB_PH_DONE) return 0;
  if (stage != 2) {
    // Non-finite objective or gradient (transform = exp overflowing fp32, a caller's objective
    // returning NaN): SciPy's line search degenerates on such values and ends ABNORMAL, and the
    // reference's argmax drops abnormal results (bore/mixins.py:83-85).  End the start here with
    // the same status; inside a line search the previous iterate is restored first, as the
    // original does when the search fails.
    int bad = !(fabs(s.f) <= DBL_MAX);
    LB_FOR(i, n) bad |= !(fabs(w.g[i]) <= DBL_MAX);
    if (lb_any(bad)) {
      if (s.phase == LB_PH_LNSRCH) {
        LB_SYNC();
        LB_FOR(i, n) { w.x[i] = w.t[i]; w.g[i] = w.r[i]; }
        s.f = s.fold;
        LB_SYNC();
      }
      lb_finish(s, 2, 0);
      return 0;
    }
  }
  if (stage == 2) {
    st = s.resume;
  } else if (s.phase == LB_PH_START) {
    s.sbgnrm = lb_projgr(n, P.nbd, P.lo, P.hi, w.x, w.g);
    if (s.sbgnrm <= P.pgtol) { lb_finish(s, 0, 401); return 0; }
    st = ST_ITER;
  } else {
    // ---- back in the line search with f,g at the trial point (lnsrlb label 556) ----
    double gd = 0.0;
    LB_FOR(i, n) gd += w.g[i] * w.d[i];
    s.gd = lb_sum(gd);
    lb_dcsrch(s, s.f, s.gd);
    if (s.ls_task == LB_LS_CONV || s.ls_task == LB_LS_WARN) {
      // NEW_X
      s.iter += 1;
      s.sbgnrm = lb_projgr(n, P.nbd, P.lo, P.hi, w.x, w.g);
      // driver (scipy/optimize/_lbfgsb_py.py:421-434)
      s.nit += 1;
      if (s.nit >= P.maxiter) { lb_finish(s, 1, 504); return 0; }
      if (s.nfev > P.maxfun) { lb_finish(s, 1, 502); return 0; }
      st = ST_TESTS;
    } else if (s.ls_task == LB_LS_ERROR) {
      st = ST_FAIL;
    } else {
      st = ST_REQUEST;
    }
  }

  LB_UNROLL1
  for (;;) {
    if (st == ST_TESTS) {
      // ---- termination tests (mainlb label 777) ----
      if (s.sbgnrm <= P.pgtol) { lb_finish(s, 0, 401); return 0; }
      const double ddum = fmax(fabs(s.fold), fmax(fabs(s.f), 1.0));
      if (s.fold - s.f <= P.ftol * ddum) { lb_finish(s, 0, 402); return 0; }
    }
    if (stage == 1 && st != ST_REQUEST) { s.resume = st; return 2; }
    if (st == ST_TESTS || st == ST_ITER) mem.load();
    if (st == ST_TESTS) {
      // ---- y = g - gold, s = x - xold ----
      double rr = 0.0;
      LB_FOR(i, n) { const double y = w.g[i] - w.r[i]; w.r[i] = y; rr += y * y; }
      rr = lb_sum(rr);
      double dr, ddd;
      if (s.stp == 1.0) {
        dr = s.gd - s.gdold;
        ddd = -s.gdold;
      } else {
        dr = (s.gd - s.gdold) * s.stp;
        LB_FOR(i, n) w.d[i] *= s.stp;
        ddd = -s.gdold * s.stp;
      }
      if (dr <= LB_EPSMCH * ddd) {
        s.nskip += 1;
        s.updatd = 0;
      } else {
        s.updatd = 1;
        s.iupdat += 1;
        mem.dirty();
        lb_matupd(P, w, s, rr, dr);
        if (lb_formt(w, m, s.col, s.theta)) lb_reset_memory(s);
        ld_fresh = 1;
      }
      st = ST_ITER;
    }

    if (st == ST_ITER) {
      // ---- new iteration (label 222): search direction ----
      mem.dirty_vec();
      // ld = L D^-1 is not part of the persisted state: rebuild it unless formt just did, or the
      // state has been in fast memory all along
      if (!ld_fresh && !mem.ld_kept() && s.col > 0) lb_prep_ld(w, m, s.col);
      ld_fresh = 1;
      int nfree = n;
      LB_UNROLL1
      for (;;) {
        mem.phase();  // (memory update done) -> Cauchy point
        if (!P.cnstnd && s.col > 0) {
          LB_SYNC();
          LB_FOR(i, n) { w.z[i] = w.x[i]; w.iwhere[i] = -1; }
          nfree = lb_freev(P, w);
        } else {
          int nseg = 0;
          if (lb_cauchy(P, w, s, nseg)) { lb_reset_memory(s); continue; }
          s.nintol += nseg;
          nfree = lb_freev(P, w);
        }
        if (nfree != 0 && s.col != 0) {
          mem.phase();  // -> K factorisation
          if (lb_formk(P, w, s, nfree)) { lb_reset_memory(s); continue; }
          mem.phase();  // -> subspace minimisation
          int info = lb_cmprlb(P, w, s, nfree);
          if (!info) info = lb_subsm(P, w, s, nfree);
          if (info) { lb_reset_memory(s); continue; }
        }
        break;
      }
      mem.phase();  // -> line-search set-up
      LB_SYNC();
      // ---- d = z - x and line-search set-up (lnsrlb, first part) ----
      double dtd = 0.0;
      LB_FOR(i, n) { const double di = w.z[i] - w.x[i]; w.d[i] = di; dtd += di * di; }
      s.dtd = lb_sum(dtd);
      s.dnorm = sqrt(s.dtd);
      double stpmx = 1e10;
      if (P.cnstnd) {
        if (s.iter == 0) {
          stpmx = 1.0;
        } else {
          // largest feasible step; the original's sequential min, done as a reduction
          double mn = 1e10;
          LB_FOR(i, n) {
            const double a1 = w.d[i];
            const int nb = P.nbd[i];
            if (nb != 0) {
              if (a1 < 0.0 && nb <= 2) {
                const double a2 = P.lo[i] - w.x[i];
                if (a2 >= 0.0) mn = 0.0;
                else if (a1 * mn < a2) mn = a2 / a1;
              } else if (a1 > 0.0 && nb >= 2) {
                const double a2 = P.hi[i] - w.x[i];
                if (a2 <= 0.0) mn = 0.0;
                else if (a1 * mn > a2) mn = a2 / a1;
              }
            }
          }
          stpmx = -lb_max(-mn);
        }
      }
      s.stpmx = stpmx;
      s.stp = (s.iter == 0 && !P.boxed) ? fmin(1.0 / s.dnorm, stpmx) : 1.0;
      double gd = 0.0;
      LB_FOR(i, n) {
        w.t[i] = w.x[i];
        w.r[i] = w.g[i];
        gd += w.g[i] * w.d[i];
      }
      s.gd = lb_sum(gd);
      s.fold = s.f;
      s.ifun = 0;
      s.iback = 0;
      s.ls_task = LB_LS_START;
      s.gdold = s.gd;
      LB_SYNC();
      if (s.gd >= 0.0) {
        st = ST_FAIL;  // not a descent direction (info = -4)
      } else {
        lb_dcsrch_start(s, s.f, s.gd);
        st = s.ls_task == LB_LS_ERROR ? ST_FAIL : ST_REQUEST;
      }
    }

    if (st == ST_REQUEST) {
      // dcsrch wants f,g at a new step (task FG_LNSRCH)
      s.ifun += 1;
      s.iback = s.ifun - 1;
      if (s.iback >= P.maxls) {
        st = ST_FAIL;
      } else {
        LB_SYNC();
        int chg = restored;
        if (s.stp == 1.0) { LB_FOR(i, n) { const double xi = w.z[i]; chg |= (w.x[i] != xi); w.x[i] = xi; } }
        else {
          // SciPy's lnsrlb also clamps the trial point into the box, so that rounding in
          // stp*d + xold cannot step outside (scipy/optimize/tests/test_lbfgsb_setulb.py:70-113)
          LB_FOR(i, n) {
            double xi = s.stp * w.d[i] + w.t[i];
            const int nb = P.nbd[i];
            if (nb == 1 || nb == 2) xi = fmax(xi, P.lo[i]);
            if (nb == 2 || nb == 3) xi = fmin(xi, P.hi[i]);
            chg |= (w.x[i] != xi);
            w.x[i] = xi;
          }
        }
        w.changed = lb_any(chg);
        LB_SYNC();
        s.phase = LB_PH_LNSRCH;
        return 1;
      }
    }

    // ST_FAIL: restore the previous iterate
    if (stage == 1) { s.resume = ST_FAIL; return 2; }
    LB_SYNC();
    LB_FOR(i, n) { w.x[i] = w.t[i]; w.g[i] = w.r[i]; }
    s.f = s.fold;
    restored = 1;
    LB_SYNC();
    if (s.col == 0) {
      s.iter += 1;
      lb_finish(s, 2, 0);  // ABNORMAL_TERMINATION_IN_LNSRCH
      return 0;
    }
    lb_reset_memory(s);  // RESTART_FROM_LNSRCH: steepest descent from here
    st = ST_ITER;
  }
}

// Set up a start: project x0 into the box, classify the variables (active).
LB_FN void lb_init_state(const LbParams &P, LbWork &w, LbScal &s) {
  const int n = P.n;
  LB_FOR(i, n) {
    double xi = w.x[i];
    const int nb = P.nbd[i];
    if (nb > 0) {
      if (nb <= 2 && xi <= P.lo[i]) xi = P.lo[i];
      else if (nb >= 2 && xi >= P.hi[i]) xi = P.hi[i];
    }
    w.x[i] = xi;
    int iw;
    if (nb == 0) iw = -1;
    else iw = (nb == 2 && P.hi[i] - P.lo[i] <= 0.0) ? 3 : 0;
    w.iwhere[i] = iw;
  }
  LB_SYNC();
  s.f = 0; s.fold = 0; s.theta = 1.0; s.gd = 0; s.gdold = 0; s.dtd = 0; s.dnorm = 0;
  s.stp = 0; s.stpmx = 0; s.sbgnrm = 0;
  s.finit = s.ginit = s.gtest = s.gx = s.gy = s.fx = s.fy = 0;
  s.stx = s.sty = s.stmin = s.stmax = s.width = s.width1 = 0;
  s.phase = LB_PH_START; s.col = 0; s.iupdat = 0; s.iter = 0; s.nit = 0; s.nfev = 0;
  s.ifun = 0; s.iback = 0; s.updatd = 0; s.status = -1; s.task = 0;
  s.brackt = 0; s.stage = 0; s.ls_task = LB_LS_START; s.nskip = 0; s.nintol = 0; s.resume = 0;
}
