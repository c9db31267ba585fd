/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's fused objective.
 *
 * E(X) = (1/p) sum_k f_k(||x_i - x_j||) and dE/dX, following cvxgrp/pymde v0.2.1
 * pymde/average_distortion.py:36-80 with the penalties / losses of pymde/functions/penalties.py:112-400
 * and pymde/functions/losses.py:61-239 in closed form (same table as oracle/mde_oracle.py, which is the
 * pinned restatement; tests/test_oracle_c.py checks this file against it and against the golden fixtures).
 * Double precision throughout: this is the arbiter for full-size runs where numpy is too slow
 * (2e8 edges).  Only tests/, __graft_entry__.smoke() and bench.py may load it; the product never does.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static double sgn(double x) { return (double)(x > 0) - (double)(x < 0); }

/* f and f' for one function id (ids = include/mde_b200.h MDE_FN_*) */
static void eval_fn(int fn, const double* s, double d, double a, double b, double* f, double* fp) {
  const double s0 = s[0], s1 = s[1];
  double de, r, sg;
  switch (fn) {
    case 1: *f = a * d; *fp = a; return;                                   /* penalties.py:112 */
    case 2: *f = a * d * d; *fp = 2 * a * d; return;                       /* :123 */
    case 3: *f = a * d * d * d; *fp = 3 * a * d * d; return;               /* :163 */
    case 4: *f = a * pow(d, s0); *fp = a * s0 * pow(d, s0 - 1); return;    /* :191 */
    case 5:                                                                  /* :205 */
      if (d < s0) { *f = 0.5 * a * d * d; *fp = a * d; } else { *f = a * s0 * (d - 0.5 * s0); *fp = a * s0; }
      return;
    case 6: { double z = s1 * (d - s0);                                     /* :246 */
      *f = a * (fmax(z, 0) + log1p(exp(-fabs(z)))); *fp = a * s1 / (1 + exp(-z)); return; }
    case 7: de = pow(d, s0); *f = a * log1p(de); *fp = a * s0 * pow(d, s0 - 1) / (1 + de); return;   /* :310 */
    case 8: de = pow(d, s0); *f = a * log(-expm1(-de)); *fp = a * s0 * pow(d, s0 - 1) / expm1(de); return; /* :324 */
    case 9: *f = fabs(a) / pow(d, s0); *fp = -fabs(a) * s0 * pow(d, -s0 - 1); return;                /* :340 */
    case 10: de = pow(d, s0); *f = a * log(de / (1 + de)); *fp = a * s0 / (d * (1 + de)); return;    /* :356 */
  }
  r = fabs(a - d); sg = sgn(d - a);
  switch (fn) {
    case 20: *f = r; *fp = sg; return;                                      /* losses.py:166 */
    case 21: *f = r * r; *fp = 2 * (d - a); return;                         /* :61 */
    case 22: *f = b * r * r; *fp = 2 * b * (d - a); return;                 /* :72 */
    case 23:                                                                 /* :101 */
      if (r < s0) { *f = r * r; *fp = 2 * (d - a); } else { *f = s0 * (2 * r - s0); *fp = 2 * s0 * sg; }
      return;
    case 24: *f = r * r * r; *fp = 3 * r * r * sg; return;                  /* :128 */
    case 25: *f = pow(r, s0); *fp = s0 * pow(r, s0 - 1) * sg; return;       /* :139 */
    case 26: { double er = exp(r); *f = log(1 + er); *fp = er / (1 + er) * sg; return; }   /* :177 */
    case 27: { double u = a / d, v = d / a, du = -a / (d * d), dv = 1 / a;                  /* :189 */
      *f = fmax(u, v) - 1; *fp = v > u ? dv : (u > v ? du : 0.5 * (du + dv)); return; }
    case 28: { double u = s0 * a / d, v = s0 * d / a, mx = fmax(u, v);                      /* :203 */
      double lse = isinf(mx) ? mx : mx + log(exp(u - mx) + exp(v - mx));
      *f = (lse - (log(2.0) + s0)) / s0;
      *fp = exp(u - lse) * (-a / (d * d)) + exp(v - lse) / a; return; }
  }
  *f = 0; *fp = 0;
}

/* value (sum over this edge list / p_total) and gradient (n*m doubles, zeroed here unless NULL). */
int mde_oracle_eval(int64_t n, int64_t m, int64_t p, const int64_t* edges, int fn_att, int fn_rep,
                    const double* att, const double* rep, int push_pull, const float* par0, const float* par1,
                    const float* X, int64_t p_total, double* value, double* grad) {
  double total = 0.0;
  const double inv_p = 1.0 / (double)p_total;
  if (grad) memset(grad, 0, sizeof(double) * (size_t)(n * m));
#pragma omp parallel for reduction(+ : total) schedule(static)
  for (int64_t k = 0; k < p; ++k) {
    const int64_t i = edges[2 * k], j = edges[2 * k + 1];
    double d2 = 0.0;
    for (int64_t c = 0; c < m; ++c) { double df = (double)X[i * m + c] - (double)X[j * m + c]; d2 += df * df; }
    const double d = sqrt(d2);
    const double a = par0[k], b = par1 ? par1[k] : 0.0;
    double f, fp;
    if (push_pull && !(a >= 0)) eval_fn(fn_rep, rep, d, a, b, &f, &fp);   /* penalties.py:390 */
    else eval_fn(fn_att, att, d, a, b, &f, &fp);
    total += f;
    if (grad) {
      double g = (fp * inv_p) / d;                                        /* average_distortion.py:55 */
      if (!isfinite(g)) g = 1.0;                                          /* :57-62 */
      for (int64_t c = 0; c < m; ++c) {
        const double v = g * ((double)X[i * m + c] - (double)X[j * m + c]);
#pragma omp atomic
        grad[i * m + c] += v;                                             /* :77 */
#pragma omp atomic
        grad[j * m + c] -= v;                                             /* :78 */
      }
    }
  }
  *value = total * inv_p;
  return 0;
}
