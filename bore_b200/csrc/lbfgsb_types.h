// L-BFGS-B (v3.0) for ONE start: problem description, persisted per-start scalars and the
// scalar More'-Thuente line search -- the part of lbfgsb_core.h that does not depend on how a
// start is mapped onto threads.
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
#include <stddef.h>

#ifdef __CUDACC__
#define LB_HD __host__ __device__ inline
#else
#define LB_HD inline
#endif

#define LB_MMAX 10
#define LB_INF (1.0 / 0.0)
#define LB_EPSMCH DBL_EPSILON

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
  int resume;  // state to re-enter lb_advance at when a step is split (light / heavy stage)
};

#define LB_LDW(m) (2 * (m) + 1)
// persisted small matrices per start: sy ss yy tinv, each [m][m].  ld = L D^-1 is derived from
// sy (lb_prep_ld, ~70 instructions): it is rebuilt whenever a start's state is staged back into
// fast memory for a new iteration and lives in the otherwise unused (2,1) block of the K-matrix
// scratch wn -- 800 bytes per start that buy a twelfth resident warp per SM at n = 50.
#define LB_NPERSIST_MM 4
// A start's persisted block, in this order in HBM *and* at the head of its workspace, so that
// staging it in or out is ONE linear (bulk) copy:
//   scalars (LbScal, LB_SCAL_DOUBLES doubles) | t r d z (4 x LB_NV) | W (LB_NW) | sy ss yy tinv
// Vector and W extents are rounded up to even counts: every piece starts 16-byte aligned.
#define LB_SCAL_DOUBLES 32
#define LB_NV(n) (((n) + 1) & ~1)
#define LB_NW(n, m) (((n) * LB_LDW(m) + 1) & ~1)
// (rounded up to an even count: the block is moved by bulk copies whose size must be a multiple
// of 16 bytes -- 5 m^2 is odd for odd m)
#define LB_PERSIST_DOUBLES(n, m) \
  ((LB_SCAL_DOUBLES + 4 * LB_NV(n) + LB_NW(n, m) + LB_NPERSIST_MM * (m) * (m) + 1) & ~1)

// ------------------------------------------------------------------ More'-Thuente step (dcstep)
LB_HD void lb_dcstep(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy,
                     double &stp, double fp, double dp, int &brackt, double stpmin, double stpmax) {
  // The four cases of the original share one cubic-interpolation kernel
  //   theta = 3 (fa - fp) / (stp - sta) + da + dp,  s = max(|theta|, |da|, |dp|),
  //   gamma = +-s sqrt((theta/s)^2 - (da/s)(dp/s))
  // taken on the pair (sta, fa, da) = the x end point (cases 1-3) or the y end point (case 4).  It
  // is written ONCE here (same operations in the same order as the four copies of the original, so
  // the results are bit-identical): on the device the three divisions and the square root are
  // ~120 instructions, and four copies of them were 15 KB of instruction footprint for a routine
  // that runs once per step (profiles/r02_notes.md).
  const double sgnd = dp * (dx / fabs(dx));
  const int kase = fp > fx ? 1 : (sgnd < 0.0 ? 2 : (fabs(dp) < fabs(dx) ? 3 : 4));
  double stpf;
  double theta = 0.0, gamma = 0.0;
  if (kase != 4 || brackt) {
    const bool yend = kase == 4;
    const double sta = yend ? sty : stx, fa = yend ? fy : fx, da = yend ? dy : dx;
    // (cases 1-3: 3 (fx - fp) / (stp - stx); case 4: 3 (fp - fy) / (sty - stp) -- the same quotient)
    theta = (yend ? 3.0 * (fp - fa) / (sta - stp) : 3.0 * (fa - fp) / (stp - sta)) + da + dp;
    const double s = fmax(fabs(theta), fmax(fabs(da), fabs(dp)));
    double rad = (theta / s) * (theta / s) - (da / s) * (dp / s);
    if (kase == 3) rad = fmax(0.0, rad);
    gamma = s * sqrt(rad);
    const bool flip = kase == 1 ? stp < stx : (kase == 4 ? stp > sty : stp > stx);
    if (flip) gamma = -gamma;
  }
  if (kase == 1) {
    const double p = (gamma - dx) + theta;
    const double q = ((gamma - dx) + gamma) + dp;
    const double r = p / q;
    const double stpc = stx + r * (stp - stx);
    const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    if (fabs(stpc - stx) < fabs(stpq - stx)) stpf = stpc;
    else stpf = stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (kase == 2) {
    const double p = (gamma - dp) + theta;
    const double q = ((gamma - dp) + gamma) + dx;
    const double r = p / q;
    const double stpc = stp + r * (stx - stp);
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
    else stpf = stpq;
    brackt = 1;
  } else if (kase == 3) {
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
// first call of a search (task START): set up the interval, ask for f,g at the first step
LB_HD void lb_dcsrch_start(LbScal &s, double f, double g) {
  const double ftol = 1e-3, stpmin = 0.0, xtrapu = 4.0;
  const double stpmax = s.stpmx;
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
}
// every later call (task FG): f,g at the trial step have arrived
LB_HD void lb_dcsrch(LbScal &s, double f, double g) {
  const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0;
  const double stpmax = s.stpmx;
  const double xtrapl = 1.1, xtrapu = 4.0;
  (void)ftol;
  const double ftest = s.finit + s.stp * s.gtest;
  if (s.stage == 1 && f <= ftest && g >= 0.0) s.stage = 2;
  int task = LB_LS_FG;
  if (s.brackt && (s.stp <= s.stmin || s.stp >= s.stmax)) task = LB_LS_WARN;
  if (s.brackt && s.stmax - s.stmin <= xtol * s.stmax) task = LB_LS_WARN;
  if (s.stp == stpmax && f <= ftest && g <= s.gtest) task = LB_LS_WARN;
  if (s.stp == stpmin && (f > ftest || g >= s.gtest)) task = LB_LS_WARN;
  if (f <= ftest && fabs(g) <= gtol * (-s.ginit)) task = LB_LS_CONV;
  if (task != LB_LS_FG) { s.ls_task = task; return; }
  {
    // one dcstep call site for both branches (the modified-function branch works on shifted
    // copies and shifts back afterwards; the arithmetic is the original's)
    const bool mod = s.stage == 1 && f <= s.fx && f > ftest;
    double fxv = s.fx, fyv = s.fy, gxv = s.gx, gyv = s.gy, fv = f, gv = g;
    if (mod) {
      fv = f - s.stp * s.gtest;
      fxv = s.fx - s.stx * s.gtest; fyv = s.fy - s.sty * s.gtest;
      gv = g - s.gtest;
      gxv = s.gx - s.gtest; gyv = s.gy - s.gtest;
    }
    lb_dcstep(s.stx, fxv, gxv, s.sty, fyv, gyv, s.stp, fv, gv, s.brackt, s.stmin, s.stmax);
    if (mod) {
      fxv = fxv + s.stx * s.gtest;
      fyv = fyv + s.sty * s.gtest;
      gxv = gxv + s.gtest;
      gyv = gyv + s.gtest;
    }
    s.fx = fxv; s.fy = fyv; s.gx = gxv; s.gy = gyv;
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

