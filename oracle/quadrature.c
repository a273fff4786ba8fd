/* TEST INFRASTRUCTURE ONLY (oracle).  See quadrature.h. */
#include <math.h>
#include "quadrature.h"

static const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                              0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                              0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                              0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
static const double wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                              0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                              0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                              0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
static const double wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                             0.381830050505118944950369775488975, 0.417959183673469387755102040816327};

static double gk15(oracle_integrand f, void *p, double a, double b, double *err)
{
  double c = 0.5 * (a + b), h = 0.5 * (b - a);
  double fc = f(c, p);
  double rk = wgk[7] * fc, rg = wg[3] * fc;
  for (int j = 0; j < 7; j++) {
    double dx = h * xgk[j];
    double f1 = f(c - dx, p), f2 = f(c + dx, p);
    rk += wgk[j] * (f1 + f2);
    if (j & 1) rg += wg[j / 2] * (f1 + f2);
  }
  *err = fabs((rk - rg) * h);
  return rk * h;
}

static double adapt(oracle_integrand f, void *p, double a, double b, double whole, double err, double tol,
                    int depth, double *errsum)
{
  if (err <= tol || depth >= 48) { *errsum += err; return whole; }
  double c = 0.5 * (a + b), e1, e2;
  double l = gk15(f, p, a, c, &e1), r = gk15(f, p, c, b, &e2);
  return adapt(f, p, a, c, l, e1, 0.5 * tol, depth + 1, errsum) + adapt(f, p, c, b, r, e2, 0.5 * tol, depth + 1, errsum);
}

double oracle_integrate(oracle_integrand f, void *params, double a, double b, double epsrel, double *abserr)
{
  double e0, errsum = 0;
  double whole = gk15(f, params, a, b, &e0);
  double tol = fabs(whole) * epsrel;
  if (tol < 1e-300) tol = 1e-300;
  double res = adapt(f, params, a, b, whole, e0, tol, 0, &errsum);
  /* one refinement round with the converged magnitude as the scale */
  if (errsum > fabs(res) * epsrel) {
    errsum = 0;
    res = adapt(f, params, a, b, whole, e0, fabs(res) * epsrel * 0.1, 0, &errsum);
  }
  if (abserr) *abserr = errsum;
  return res;
}

typedef struct { oracle_integrand f; void *p; double b; } inf_map;
static double mapped(double t, void *vp)
{
  inf_map *m = (inf_map *)vp;
  if (t <= 0) return 0;
  double x = m->b - (1 - t) / t;
  double v = m->f(x, m->p) / (t * t);
  return isfinite(v) ? v : 0;
}

double oracle_integrate_lower_inf(oracle_integrand f, void *params, double b, double epsrel, double *abserr)
{
  inf_map m = {f, params, b};
  return oracle_integrate(mapped, &m, 0.0, 1.0, epsrel, abserr);
}
