// Host build of bore_b200/csrc/lbfgsb_core.h (LB_NL == 1): TEST INFRASTRUCTURE ONLY.
// Lets tests/ pin the L-BFGS-B stepper against SciPy's setulb on the CPU, request by
// request, with identical f,g.  The product never links this file.
#include <stdlib.h>
#include <string.h>
#include <vector>

#define LB_VARIANT 0
#include "../../bore_b200/csrc/lbfgsb_core.h"

struct HostSim {
  int split = 0;  // 1: run every step as a LIGHT stage followed, if asked for, by a HEAVY stage
  LbParams P;
  std::vector<double> lo, hi, dbuf, xlast;
  std::vector<int> nbd, ibuf;
  LbWork w;
  LbScal s;
};

extern "C" {

void *hs_create(int n, int m, const double *lo, const double *hi, double ftol, double pgtol,
                int maxiter, int maxfun, int maxls) {
  HostSim *h = new HostSim();
  h->lo.assign(lo, lo + n);
  h->hi.assign(hi, hi + n);
  h->nbd.resize(n);
  int cnstnd = 0, boxed = 1;
  for (int i = 0; i < n; ++i) {
    const bool L = !isinf(lo[i]), U = !isinf(hi[i]);
    h->nbd[i] = L ? (U ? 2 : 1) : (U ? 3 : 0);
    if (!L) h->lo[i] = 0.0;
    if (!U) h->hi[i] = 0.0;
    if (h->nbd[i] != 2) boxed = 0;
    if (h->nbd[i] != 0) cnstnd = 1;
  }
  h->P.n = n; h->P.m = m; h->P.maxiter = maxiter; h->P.maxfun = maxfun; h->P.maxls = maxls;
  h->P.cnstnd = cnstnd; h->P.boxed = boxed; h->P.ftol = ftol; h->P.pgtol = pgtol;
  h->P.lo = h->lo.data(); h->P.hi = h->hi.data(); h->P.nbd = h->nbd.data();
  h->dbuf.assign(lb_work_doubles(n, m), 0.0);
  h->ibuf.assign(lb_work_ints(n), 0);
  h->xlast.assign(n, NAN);
  lb_carve(h->w, h->dbuf.data(), h->ibuf.data(), n, m);
  return h;
}

void hs_destroy(void *p) { delete (HostSim *)p; }
void hs_set_split(void *p, int on) { ((HostSim *)p)->split = on; }

// returns 1; xreq = first request (x0 projected)
int hs_start(void *p, const double *x0, double *xreq) {
  HostSim *h = (HostSim *)p;
  memcpy(h->w.x, x0, h->P.n * sizeof(double));
  lb_init_state(h->P, h->w, h->s);
  memcpy(xreq, h->w.x, h->P.n * sizeof(double));
  h->s.nfev = 1;
  h->xlast.assign(xreq, xreq + h->P.n);
  return 1;
}

// feed f,g for the last request; returns 1 if another request is in xreq, 0 if finished
int hs_step(void *p, double f, const double *g, double *xreq) {
  HostSim *h = (HostSim *)p;
  const int n = h->P.n;
  h->s.f = f;
  memcpy(h->w.g, g, n * sizeof(double));
  LbNoMem nomem;
  int pend;
  if (h->split) {
    pend = lb_advance(h->P, h->w, h->s, nomem, 1);
    if (pend == 2) pend = lb_advance(h->P, h->w, h->s, nomem, 2);
  } else {
    pend = lb_advance(h->P, h->w, h->s, nomem);
  }
  if (pend) {
    bool same = true;
    for (int i = 0; i < n; ++i) same = same && (h->w.x[i] == h->xlast[i]);
    if (!same) {
      h->s.nfev += 1;
      h->xlast.assign(h->w.x, h->w.x + n);
    }
  }
  memcpy(xreq, h->w.x, n * sizeof(double));
  return pend;
}

void hs_result(void *p, double *x, double *f, int *nit, int *nfev, int *status, int *task) {
  HostSim *h = (HostSim *)p;
  memcpy(x, h->w.x, h->P.n * sizeof(double));
  *f = h->s.f; *nit = h->s.nit; *nfev = h->s.nfev; *status = h->s.status; *task = h->s.task;
}

// internals for trajectory debugging: theta, stp, col, iter
void hs_peek(void *p, double *theta, double *stp, int *col, int *iter) {
  HostSim *h = (HostSim *)p;
  *theta = h->s.theta; *stp = h->s.stp; *col = h->s.col; *iter = h->s.iter;
}
}
