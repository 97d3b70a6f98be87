// L-BFGS-B (v3.0) for ONE start, written warp-collectively.
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
// Execution model.  On the device one WARP owns one start: vectors of length n are spread
// over the 32 lanes (`LB_FOR`), reductions are xor-butterflies (bit-identical on every lane),
// small dense factorisations run column-parallel with __syncwarp between steps, and scalar
// control flow is replicated on every lane (warp-uniform).  Compiled for the host
// (LB_NL == 1) the same code is a plain serial program, which is how it is pinned against
// SciPy's setulb request by request in tests/ -- the host build is test infrastructure; the
// product only ever runs the device build.
//
// All algebra is fp64 (SciPy's is); the objective and gradient arrive as fp32 values from
// the MLP kernel, exactly like the reference's fp32 Keras model feeding fp64 SciPy
// (bore/decorators.py:54-56).
#pragma once
#include <float.h>
#include <math.h>

#ifdef __CUDACC__
#define LB_HD __host__ __device__ inline
#else
#define LB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define LB_LANE ((int)(threadIdx.x & 31))
#define LB_NL 32
#define LB_SYNC() __syncwarp()
#else
#define LB_LANE 0
#define LB_NL 1
#define LB_SYNC() ((void)0)
#endif

#define LB_FOR(i, n) for (int i = LB_LANE; i < (n); i += LB_NL)
#define LB_MMAX 10
#define LB_INF (1.0 / 0.0)
#define LB_EPSMCH DBL_EPSILON

// ------------------------------------------------------------------ warp collectives
LB_HD double lb_sum(double v) {
#if defined(__CUDA_ARCH__)
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}
LB_HD double lb_max(double v) {
#if defined(__CUDA_ARCH__)
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
  return v;
}
LB_HD int lb_isum(int v) {
#if defined(__CUDA_ARCH__)
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}
LB_HD int lb_any(int p) {
#if defined(__CUDA_ARCH__)
  return __any_sync(0xffffffffu, p);
#else
  return p;
#endif
}
// minimum value and the SMALLEST index attaining it
LB_HD void lb_argmin(double &v, int &idx) {
#if defined(__CUDA_ARCH__)
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
#endif
}

// ------------------------------------------------------------------ problem + state
struct LbParams {
  int n, m;
  int maxiter, maxfun, maxls;
  int cnstnd, boxed;  // any bounded variable / all variables boxed
  double ftol;        // factr * epsmch
  double pgtol;
  const double *lo, *hi;  // [n]
  const int *nbd;         // [n] 0 none, 1 lower, 2 both, 3 upper
};

enum { LB_PH_START = 0, LB_PH_LNSRCH = 1, LB_PH_DONE = 2 };
enum { LB_LS_START = 0, LB_LS_FG = 1, LB_LS_CONV = 2, LB_LS_WARN = 3, LB_LS_ERROR = 4 };

// persisted per start between evaluation rounds
struct LbScal {
  double f, fold, theta, gd, gdold, dtd, dnorm, stp, stpmx, sbgnrm;
  // dcsrch
  double finit, ginit, gtest, gx, gy, fx, fy, stx, sty, stmin, stmax, width, width1;
  int phase, col, iupdat, iter, nit, nfev, ifun, iback, updatd, status, task;
  int brackt, stage, ls_task, nskip, nintol;
};

#define LB_LDW(m) (2 * (m) + 1)

// scratch + state views for one start (all in fast memory while a step runs)
struct LbWork {
  double *x, *g, *z, *r, *d, *t, *xp;  // [n]
  double *W;                          // [n][LDW]: wy cols 0..m-1, ws cols m..2m-1 (logical order)
  double *sy, *ss, *wt;               // [m][m]
  double *wn;                         // [2m][2m] upper triangle
  double *p, *c, *wbp, *v;            // [2m]
  int *iwhere;                        // [n]
  int *index;                         // [n] free variables first (count nfree), active after
};

LB_HD size_t lb_work_doubles(int n, int m) {
  return (size_t)7 * n + (size_t)n * LB_LDW(m) + 3 * m * m + 4 * m * m + 8 * m;
}
LB_HD size_t lb_work_ints(int n) { return 2 * (size_t)n; }

LB_HD void lb_carve(LbWork &w, double *dbase, int *ibase, int n, int m) {
  double *q = dbase;
  w.x = q; q += n; w.g = q; q += n; w.z = q; q += n; w.r = q; q += n;
  w.d = q; q += n; w.t = q; q += n; w.xp = q; q += n;
  w.W = q; q += (size_t)n * LB_LDW(m);
  w.sy = q; q += m * m; w.ss = q; q += m * m; w.wt = q; q += m * m;
  w.wn = q; q += 4 * m * m;
  w.p = q; q += 2 * m; w.c = q; q += 2 * m; w.wbp = q; q += 2 * m; w.v = q; q += 2 * m;
  w.iwhere = ibase; w.index = ibase + n;
}

// ------------------------------------------------------------------ small dense kernels
// Cholesky A = R'R in place, R in the upper triangle (LINPACK dpofa).  0 ok, else k+1.
LB_HD int lb_chol(double *A, int ld, int n) {
  for (int k = 0; k < n; ++k) {
    LB_SYNC();
    const double akk = A[k * ld + k];
    if (!(akk > 0.0)) return k + 1;
    const double rkk = sqrt(akk);
    for (int j = k + 1 + LB_LANE; j < n; j += LB_NL) A[k * ld + j] /= rkk;
    LB_SYNC();
    if (LB_LANE == 0) A[k * ld + k] = rkk;
    const int r = n - k - 1;  // trailing block, pairs (i<=j) in k+1..n-1
    for (int e = LB_LANE; e < r * r; e += LB_NL) {
      const int i = k + 1 + e / r, j = k + 1 + e % r;
      if (i <= j) A[i * ld + j] -= A[k * ld + i] * A[k * ld + j];
    }
  }
  LB_SYNC();
  return 0;
}

// solve R' x = b, R upper (dtrsl job 11); b overwritten
LB_HD int lb_trsl_t(const double *R, int ld, int n, double *b) {
  for (int k = 0; k < n; ++k)
    if (R[k * ld + k] == 0.0) return k + 1;
  double xprev = 0.0;
  for (int k = 0; k < n; ++k) {
    LB_SYNC();
    if (k > 0 && LB_LANE == 0) b[k - 1] = xprev;
    const double xk = b[k] / R[k * ld + k];
    for (int i = k + 1 + LB_LANE; i < n; i += LB_NL) b[i] -= R[k * ld + i] * xk;
    xprev = xk;
  }
  LB_SYNC();
  if (n > 0 && LB_LANE == 0) b[n - 1] = xprev;
  LB_SYNC();
  return 0;
}

// solve R x = b, R upper (dtrsl job 01); b overwritten
LB_HD int lb_trsl_n(const double *R, int ld, int n, double *b) {
  for (int k = 0; k < n; ++k)
    if (R[k * ld + k] == 0.0) return k + 1;
  double xprev = 0.0;
  for (int k = n - 1; k >= 0; --k) {
    LB_SYNC();
    if (k < n - 1 && LB_LANE == 0) b[k + 1] = xprev;
    const double xk = b[k] / R[k * ld + k];
    for (int i = LB_LANE; i < k; i += LB_NL) b[i] -= R[i * ld + k] * xk;
    xprev = xk;
  }
  LB_SYNC();
  if (n > 0 && LB_LANE == 0) b[0] = xprev;
  LB_SYNC();
  return 0;
}

// p = M v with M the 2col x 2col middle matrix of the compact L-BFGS formula (bmv)
LB_HD int lb_bmv(const double *sy, const double *wt, int m, int col, const double *v, double *p) {
  if (col == 0) return 0;
  LB_SYNC();
  LB_FOR(i, col) {
    double s = 0.0;
    for (int k = 0; k < i; ++k) s += sy[i * m + k] * v[k] / sy[k * m + k];
    p[col + i] = v[col + i] + s;
    p[i] = v[i] / sqrt(sy[i * m + i]);
  }
  int info = lb_trsl_t(wt, m, col, p + col);
  if (info) return info;
  info = lb_trsl_n(wt, m, col, p + col);
  if (info) return info;
  LB_FOR(i, col) {
    double pi = -p[i] / sqrt(sy[i * m + i]);
    double s = 0.0;
    for (int k = i + 1; k < col; ++k) s += sy[k * m + i] * p[col + k] / sy[i * m + i];
    p[i] = pi + s;
  }
  LB_SYNC();
  return 0;
}

// T = theta*SS + L D^-1 L' (upper), then Cholesky into wt (formt)
LB_HD int lb_formt(double *wt, const double *sy, const double *ss, int m, int col, double theta) {
  LB_SYNC();
  for (int e = LB_LANE; e < col * col; e += LB_NL) {
    const int i = e / col, j = e % col;
    if (i > j) continue;
    double s = 0.0;
    for (int k = 0; k < i; ++k) s += sy[i * m + k] * sy[j * m + k] / sy[k * m + k];
    wt[i * m + j] = s + theta * ss[i * m + j];
  }
  const int info = lb_chol(wt, m, col);
  return info ? -3 : 0;
}

// ------------------------------------------------------------------ projected gradient norm
LB_HD double lb_projgr(const LbParams &P, const double *x, const double *g) {
  double s = 0.0;
  LB_FOR(i, P.n) {
    double gi = g[i];
    const int nb = P.nbd[i];
    if (nb != 0) {
      if (gi < 0.0) {
        if (nb >= 2) gi = fmax(x[i] - P.hi[i], gi);
      } else {
        if (nb <= 2) gi = fmin(x[i] - P.lo[i], gi);
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
LB_HD int lb_cauchy(const LbParams &P, LbWork &w, const LbScal &s, int &nseg_out) {
  const int n = P.n, m = P.m, col = s.col, col2 = 2 * col, ldw = LB_LDW(m);
  const double theta = s.theta;
  double *tb = w.t, *d = w.d, *xcp = w.z;
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
  for (int j = LB_LANE; j < col2; j += LB_NL) {
    const double *wc = w.W + (j < col ? j : m + (j - col));
    double a = 0.0;
    for (int i = 0; i < n; ++i) a += wc[i * ldw] * d[i];
    w.p[j] = j < col ? a : theta * a;
    w.c[j] = 0.0;
  }
  LB_SYNC();
  if (nbreak == 0 && nmove == 0) return 0;  // d == 0: xcp = x

  double f2 = -theta * f1;
  const double f2_org = f2;
  if (col > 0) {
    const int info = lb_bmv(w.sy, w.wt, m, col, w.p, w.v);
    if (info) return info;
    double a = 0.0;
    for (int j = LB_LANE; j < col2; j += LB_NL) a += w.v[j] * w.p[j];
    f2 -= lb_sum(a);
  }
  double dtm = -f1 / f2, tsum = 0.0;
  int nseg = 1;
  bool all_fixed = false;
  if (nbreak > 0) {
    int nleft = nbreak;
    double tj = 0.0;
    for (;;) {
      const double tj0 = tj;
      // smallest remaining breakpoint
      double bv = LB_INF;
      int bi = 0x7fffffff;
      LB_FOR(i, n) if (tb[i] < bv) { bv = tb[i]; bi = i; }
      lb_argmin(bv, bi);
      tj = bv;
      const int ibp = bi;
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
        for (int j = LB_LANE; j < col2; j += LB_NL) {
          w.c[j] += dt * w.p[j];
          w.wbp[j] = j < col ? w.W[ibp * ldw + j] : theta * w.W[ibp * ldw + m + (j - col)];
        }
        const int info = lb_bmv(w.sy, w.wt, m, col, w.wbp, w.v);
        if (info) return info;
        double wmc = 0.0, wmp = 0.0, wmw = 0.0;
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
    for (int j = LB_LANE; j < col2; j += LB_NL) w.c[j] += dtm * w.p[j];
  LB_SYNC();
  nseg_out = nseg;
  return 0;
}

// ------------------------------------------------------------------ free / active sets at the GCP
// index[0..nfree) = free variables (iwhere <= 0) in ascending order, the rest active.
LB_HD int lb_freev(const LbParams &P, LbWork &w) {
  const int n = P.n;
  LB_SYNC();
#if defined(__CUDA_ARCH__)
  int base_f = 0, base_a = 0;
  // ordered compaction, 32 variables at a time
  int nfree_total = 0;
  {
    int c = 0;
    LB_FOR(i, n) c += (w.iwhere[i] <= 0);
    nfree_total = lb_isum(c);
  }
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
  int nfree = 0;
  for (int i = 0; i < n; ++i) if (w.iwhere[i] <= 0) w.index[nfree++] = i;
  int na = nfree;
  for (int i = 0; i < n; ++i) if (w.iwhere[i] > 0) w.index[na++] = i;
  return nfree;
#endif
}

// ------------------------------------------------------------------ K = LEL' factorisation (formk)
// Built from scratch each time it is needed (the original updates it incrementally; the
// matrix is the same).  wn: upper triangle of the 2col x 2col matrix, leading dim 2m.
LB_HD int lb_formk(const LbParams &P, LbWork &w, const LbScal &s, int nfree) {
  const int n = P.n, m = P.m, col = s.col, ldw = LB_LDW(m), ldn = 2 * m;
  const double theta = s.theta;
  const double *W = w.W;
  const int *ind = w.index;
  LB_SYNC();
  const int ntri = col * (col + 1) / 2;
  const int ntask = 2 * ntri + col * col;
  for (int e = LB_LANE; e < ntask; e += LB_NL) {
    if (e < 2 * ntri) {
      const bool second = e >= ntri;
      int q = second ? e - ntri : e;
      // unrank (iy >= jy) from q = iy*(iy+1)/2 + jy
      int iy = 0;
      while ((iy + 1) * (iy + 2) / 2 <= q) ++iy;
      const int jy = q - iy * (iy + 1) / 2;
      double a = 0.0;
      if (!second) {  // Y'ZZ'Y over free variables
        for (int k = 0; k < nfree; ++k) {
          const double *row = W + ind[k] * ldw;
          a += row[iy] * row[jy];
        }
        a /= theta;
        if (iy == jy) a += w.sy[iy * m + iy];
        w.wn[jy * ldn + iy] = a;
      } else {  // S'AA'S over active variables
        for (int k = nfree; k < n; ++k) {
          const double *row = W + ind[k] * ldw + m;
          a += row[iy] * row[jy];
        }
        w.wn[(col + jy) * ldn + (col + iy)] = a * theta;
      }
    } else {
      const int q = e - 2 * ntri;
      const int iy = q / col, jy = q % col;  // ws column iy with wy column jy
      double a = 0.0;
      if (jy >= iy) {  // R_z: free variables
        for (int k = 0; k < nfree; ++k) {
          const double *row = W + ind[k] * ldw;
          a += row[m + iy] * row[jy];
        }
        w.wn[jy * ldn + (col + iy)] = a;
      } else {  // L_a: active variables, negated
        for (int k = nfree; k < n; ++k) {
          const double *row = W + ind[k] * ldw;
          a += row[m + iy] * row[jy];
        }
        w.wn[jy * ldn + (col + iy)] = -a;
      }
    }
  }
  // Cholesky of the (1,1) block
  if (lb_chol(w.wn, ldn, col)) return -1;
  // (1,2) block: L^-1 (-L_a' + R_z'), one right-hand side per lane
  for (int js = col + LB_LANE; js < 2 * col; js += LB_NL) {
    for (int k = 0; k < col; ++k) {
      double b = w.wn[k * ldn + js];
      for (int i = 0; i < k; ++i) b -= w.wn[i * ldn + k] * w.wn[i * ldn + js];
      w.wn[k * ldn + js] = b / w.wn[k * ldn + k];
    }
  }
  LB_SYNC();
  // (2,2) block += (1,2)'(1,2), upper triangle
  for (int e = LB_LANE; e < col * col; e += LB_NL) {
    const int is = col + e / col, js = col + e % col;
    if (is > js) continue;
    double a = 0.0;
    for (int k = 0; k < col; ++k) a += w.wn[k * ldn + is] * w.wn[k * ldn + js];
    w.wn[is * ldn + js] += a;
  }
  if (lb_chol(w.wn + col * ldn + col, ldn, col)) return -2;
  return 0;
}

// ------------------------------------------------------------------ reduced gradient (cmprlb)
// r[k] = -(B(xcp - x) + g)[k] for free k (0 elsewhere); uses c = W'(xcp - x) from cauchy.
LB_HD int lb_cmprlb(const LbParams &P, LbWork &w, const LbScal &s, int nfree) {
  const int n = P.n, m = P.m, col = s.col, ldw = LB_LDW(m);
  const double theta = s.theta;
  LB_SYNC();
  if (!P.cnstnd && col > 0) {
    LB_FOR(i, n) w.r[i] = -w.g[i];
    LB_SYNC();
    return 0;
  }
  const int info = lb_bmv(w.sy, w.wt, m, col, w.c, w.p);
  if (info) return -8;
  LB_FOR(i, n) {
    if (w.iwhere[i] <= 0) {
      double a = -theta * (w.z[i] - w.x[i]) - w.g[i];
      const double *row = w.W + i * ldw;
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
LB_HD int lb_subsm(const LbParams &P, LbWork &w, const LbScal &s, int nfree) {
  const int n = P.n, m = P.m, col = s.col, col2 = 2 * col, ldw = LB_LDW(m), ldn = 2 * m;
  const double theta = s.theta;
  if (nfree <= 0) return 0;
  double *wv = w.v;
  LB_SYNC();
  // wv = W'Z d
  for (int j = LB_LANE; j < col2; j += LB_NL) {
    const double *wc = w.W + (j < col ? j : m + (j - col));
    double a = 0.0;
    for (int k = 0; k < nfree; ++k) { const int i = w.index[k]; a += wc[i * ldw] * w.r[i]; }
    wv[j] = j < col ? a : theta * a;
  }
  // wv := K^-1 wv
  int info = lb_trsl_t(w.wn, ldn, col2, wv);
  if (info) return info;
  for (int j = LB_LANE; j < col; j += LB_NL) wv[j] = -wv[j];
  info = lb_trsl_n(w.wn, ldn, col2, wv);
  if (info) return info;
  // d = (1/theta) d + (1/theta^2) Z'W wv ; xp = xcp ; projected Newton point
  int iword = 0;
  LB_FOR(i, n) {
    const double zi = w.z[i];
    w.xp[i] = zi;
    if (w.iwhere[i] <= 0) {
      double dk = w.r[i];
      const double *row = w.W + i * ldw;
      for (int j = 0; j < col; ++j) dk += row[j] * wv[j] / theta + row[m + j] * wv[col + j];
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
    lb_argmin(amin, ibd);
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

// ------------------------------------------------------------------ More'-Thuente step (dcstep)
LB_HD void lb_dcstep(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy,
                     double &stp, double fp, double dp, int &brackt, double stpmin, double stpmax) {
  const double sgnd = dp * (dx / fabs(dx));
  double stpf;
  if (fp > fx) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    const double p = (gamma - dx) + theta;
    const double q = ((gamma - dx) + gamma) + dp;
    const double r = p / q;
    const double stpc = stx + r * (stp - stx);
    const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    if (fabs(stpc - stx) < fabs(stpq - stx)) stpf = stpc;
    else stpf = stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta;
    const double q = ((gamma - dp) + gamma) + dx;
    const double r = p / q;
    const double stpc = stp + r * (stx - stp);
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
    else stpf = stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta;
    const double q = (gamma + (dx - dp)) + gamma;
    const double r = p / q;
    double stpc;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
      else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      stpf = fmin(stpmax, stpf);
      stpf = fmax(stpmin, stpf);
    }
  } else {
    if (brackt) {
      const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      const double s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
      double gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      const double p = (gamma - dp) + theta;
      const double q = ((gamma - dp) + gamma) + dy;
      const double r = p / q;
      stpf = stp + r * (sty - stp);
    } else if (stp > stx) {
      stpf = stpmax;
    } else {
      stpf = stpmin;
    }
  }
  if (fp > fx) {
    sty = stp; fy = fp; dy = dp;
  } else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// ------------------------------------------------------------------ line search (dcsrch)
// ftol 1e-3, gtol 0.9, xtol 0.1, stpmin 0 -- the constants lnsrlb passes.
LB_HD void lb_dcsrch(LbScal &s, double f, double g) {
  const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0;
  const double stpmax = s.stpmx;
  const double xtrapl = 1.1, xtrapu = 4.0;
  if (s.ls_task == LB_LS_START) {
    if (s.stp < stpmin || s.stp > stpmax || g >= 0.0) { s.ls_task = LB_LS_ERROR; return; }
    s.brackt = 0;
    s.stage = 1;
    s.finit = f; s.ginit = g; s.gtest = ftol * g;
    s.width = stpmax - stpmin;
    s.width1 = s.width / 0.5;
    s.stx = 0.0; s.fx = f; s.gx = g;
    s.sty = 0.0; s.fy = f; s.gy = g;
    s.stmin = 0.0;
    s.stmax = s.stp + xtrapu * s.stp;
    s.ls_task = LB_LS_FG;
    return;
  }
  const double ftest = s.finit + s.stp * s.gtest;
  if (s.stage == 1 && f <= ftest && g >= 0.0) s.stage = 2;
  int task = LB_LS_FG;
  if (s.brackt && (s.stp <= s.stmin || s.stp >= s.stmax)) task = LB_LS_WARN;
  if (s.brackt && s.stmax - s.stmin <= xtol * s.stmax) task = LB_LS_WARN;
  if (s.stp == stpmax && f <= ftest && g <= s.gtest) task = LB_LS_WARN;
  if (s.stp == stpmin && (f > ftest || g >= s.gtest)) task = LB_LS_WARN;
  if (f <= ftest && fabs(g) <= gtol * (-s.ginit)) task = LB_LS_CONV;
  if (task != LB_LS_FG) { s.ls_task = task; return; }
  if (s.stage == 1 && f <= s.fx && f > ftest) {
    const double fm = f - s.stp * s.gtest;
    double fxm = s.fx - s.stx * s.gtest, fym = s.fy - s.sty * s.gtest;
    const double gm = g - s.gtest;
    double gxm = s.gx - s.gtest, gym = s.gy - s.gtest;
    lb_dcstep(s.stx, fxm, gxm, s.sty, fym, gym, s.stp, fm, gm, s.brackt, s.stmin, s.stmax);
    s.fx = fxm + s.stx * s.gtest;
    s.fy = fym + s.sty * s.gtest;
    s.gx = gxm + s.gtest;
    s.gy = gym + s.gtest;
  } else {
    lb_dcstep(s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, s.stp, f, g, s.brackt, s.stmin, s.stmax);
  }
  if (s.brackt) {
    if (fabs(s.sty - s.stx) >= 0.66 * s.width1) s.stp = s.stx + 0.5 * (s.sty - s.stx);
    s.width1 = s.width;
    s.width = fabs(s.sty - s.stx);
  }
  if (s.brackt) {
    s.stmin = fmin(s.stx, s.sty);
    s.stmax = fmax(s.stx, s.sty);
  } else {
    s.stmin = s.stp + xtrapl * (s.stp - s.stx);
    s.stmax = s.stp + xtrapu * (s.stp - s.stx);
  }
  s.stp = fmax(s.stp, stpmin);
  s.stp = fmin(s.stp, stpmax);
  if ((s.brackt && (s.stp <= s.stmin || s.stp >= s.stmax)) ||
      (s.brackt && s.stmax - s.stmin <= xtol * s.stmax))
    s.stp = s.stx;
  s.ls_task = LB_LS_FG;
}

// ------------------------------------------------------------------ BFGS memory update (matupd)
// Columns are kept in logical order (0 = oldest): when the memory is full everything is
// shifted by one instead of rotating a head pointer.
LB_HD void lb_matupd(const LbParams &P, LbWork &w, LbScal &s, double rr, double dr) {
  const int n = P.n, m = P.m, ldw = LB_LDW(m);
  LB_SYNC();
  if (s.iupdat <= m) {
    s.col = s.iupdat;
  } else {
    LB_FOR(i, n) {
      double *row = w.W + i * ldw;
      for (int j = 0; j < m - 1; ++j) { row[j] = row[j + 1]; row[m + j] = row[m + j + 1]; }
    }
    // new(i,j) = old(i+1,j+1): one lane per diagonal, walking down it
    const int c1 = s.col - 1;
    for (int dg = LB_LANE; dg < 2 * c1; dg += LB_NL) {
      if (dg < c1) {  // ss upper: j - i = dg
        for (int i = 0; i + dg < c1; ++i) w.ss[i * m + i + dg] = w.ss[(i + 1) * m + i + dg + 1];
      } else {        // sy lower: i - j = dg - c1
        const int e = dg - c1;
        for (int j = 0; j + e < c1; ++j) w.sy[(j + e) * m + j] = w.sy[(j + e + 1) * m + j + 1];
      }
    }
  }
  const int col = s.col, last = col - 1;
  LB_FOR(i, n) {
    w.W[i * ldw + m + last] = w.d[i];
    w.W[i * ldw + last] = w.r[i];
  }
  s.theta = rr / dr;
  LB_SYNC();
  // last row of SY, last column of SS
  for (int j = LB_LANE; j < 2 * last; j += LB_NL) {
    if (j < last) {
      double a = 0.0;
      for (int i = 0; i < n; ++i) a += w.d[i] * w.W[i * ldw + j];
      w.sy[last * m + j] = a;
    } else {
      const int jj = j - last;
      double a = 0.0;
      for (int i = 0; i < n; ++i) a += w.W[i * ldw + m + jj] * w.d[i];
      w.ss[jj * m + last] = a;
    }
  }
  if (LB_LANE == 0) {
    w.ss[last * m + last] = s.stp == 1.0 ? s.dtd : s.stp * s.stp * s.dtd;
    w.sy[last * m + last] = dr;
  }
  LB_SYNC();
}

LB_HD void lb_reset_memory(LbScal &s) {
  s.col = 0;
  s.theta = 1.0;
  s.iupdat = 0;
  s.updatd = 0;
}

// ------------------------------------------------------------------ the stepper
// Advance one start until it needs f,g at a new point (returns 1: trial point is in w.x)
// or terminates (returns 0: s.status/s.task set, final iterate in w.x).
//
// On entry w.x holds the point that was just evaluated, s.f / w.g its value and gradient
// (already stored by the caller), and the persisted vectors/matrices are loaded.
// `x_eval_changed` reports whether the new request differs from the point evaluated last
// (SciPy's ScalarFunction memoises on x, so a repeated request does not count in nfev).
LB_HD void lb_finish(LbScal &s, int status, int task) {
  s.phase = LB_PH_DONE;
  s.status = status;
  s.task = task;
}

// `Mem` supplies the limited-memory matrices lazily: mem.load() is called once, right before
// they are first needed (a step that only continues a line search never touches them), and
// mem.dirty() when they were modified.
struct LbNoMem {
  LB_HD void load() {}
  LB_HD void dirty() {}
  LB_HD void dirty_vec() {}
};

template <class Mem>
LB_HD int lb_advance(const LbParams &P, LbWork &w, LbScal &s, Mem &mem) {
  const int n = P.n, m = P.m;
  enum { ST_TESTS, ST_ITER, ST_REQUEST, ST_FAIL };
  int st;

  if (s.phase == LB_PH_DONE) return 0;
  if (s.phase == LB_PH_START) {
    s.sbgnrm = lb_projgr(P, w.x, w.g);
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
      s.sbgnrm = lb_projgr(P, w.x, w.g);
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

  for (;;) {
    if (st == ST_TESTS || st == ST_ITER) mem.load();
    if (st == ST_TESTS) {
      // ---- termination tests (mainlb label 777) ----
      if (s.sbgnrm <= P.pgtol) { lb_finish(s, 0, 401); return 0; }
      const double ddum = fmax(fabs(s.fold), fmax(fabs(s.f), 1.0));
      if (s.fold - s.f <= P.ftol * ddum) { lb_finish(s, 0, 402); return 0; }
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
        if (lb_formt(w.wt, w.sy, w.ss, m, s.col, s.theta)) lb_reset_memory(s);
      }
      st = ST_ITER;
    }

    if (st == ST_ITER) {
      // ---- new iteration (label 222): search direction ----
      mem.dirty_vec();
      int nfree = n;
      for (;;) {
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
          if (lb_formk(P, w, s, nfree)) { lb_reset_memory(s); continue; }
          int info = lb_cmprlb(P, w, s, nfree);
          if (!info) info = lb_subsm(P, w, s, nfree);
          if (info) { lb_reset_memory(s); continue; }
        }
        break;
      }
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
        lb_dcsrch(s, s.f, s.gd);
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
        if (s.stp == 1.0) { LB_FOR(i, n) w.x[i] = w.z[i]; }
        else {
          // SciPy's lnsrlb also clamps the trial point into the box, so that rounding in
          // stp*d + xold cannot step outside (scipy/optimize/tests/test_lbfgsb_setulb.py:70-113)
          LB_FOR(i, n) {
            double xi = s.stp * w.d[i] + w.t[i];
            const int nb = P.nbd[i];
            if (nb == 1 || nb == 2) xi = fmax(xi, P.lo[i]);
            if (nb == 2 || nb == 3) xi = fmin(xi, P.hi[i]);
            w.x[i] = xi;
          }
        }
        LB_SYNC();
        s.phase = LB_PH_LNSRCH;
        return 1;
      }
    }

    // ST_FAIL: restore the previous iterate
    LB_SYNC();
    LB_FOR(i, n) { w.x[i] = w.t[i]; w.g[i] = w.r[i]; }
    s.f = s.fold;
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
LB_HD void lb_init_state(const LbParams &P, LbWork &w, LbScal &s) {
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
  s.brackt = 0; s.stage = 0; s.ls_task = LB_LS_START; s.nskip = 0; s.nintol = 0;
}
