/* TEST INFRASTRUCTURE ONLY -- implements the GSL entry points declared in oracle/shim/gsl/ on top of
 * oracle/mt19937.c and oracle/quadrature.c, so the unmodified reference sources link in oracle/_ref/. */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_randist.h>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_integration.h>
#include "../mt19937.h"
#include "../quadrature.h"

static const gsl_rng_type mt_type = {"mt19937"};
static const gsl_rng_type ranlux_type = {"ranlux(unsupported)"};
const gsl_rng_type *gsl_rng_mt19937 = &mt_type;
const gsl_rng_type *gsl_rng_ranlux = &ranlux_type;

/* gsl_rng is laid out here as {unsigned long mt[624]; int mti;}; keep a private narrow state instead */
typedef struct { oracle_mt19937 g; } rng_state;

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
  if (T != gsl_rng_mt19937) { fprintf(stderr, "shim_gsl: only mt19937 is provided\n"); abort(); }
  rng_state *s = calloc(1, sizeof(gsl_rng) > sizeof(rng_state) ? sizeof(gsl_rng) : sizeof(rng_state));
  oracle_mt_seed(&s->g, 0);
  return (gsl_rng *)s;
}
void gsl_rng_set(gsl_rng *r, unsigned long seed) { oracle_mt_seed(&((rng_state *)r)->g, (uint32_t)seed); }
unsigned long gsl_rng_get(gsl_rng *r) { return oracle_mt_u32(&((rng_state *)r)->g); }
double gsl_rng_uniform(gsl_rng *r) { return oracle_mt_uniform(&((rng_state *)r)->g); }
void gsl_rng_free(gsl_rng *r) { free(r); }

unsigned int gsl_ran_poisson(gsl_rng *r, double mu)
{
  /* stand-in for GSL's sampler (same distribution, different stream): multiplication method for small means,
   * Hoermann's PTRS transformed rejection above (Insurance: Mathematics and Economics 12 (1993) 39) */
  if (!(mu > 0)) return 0;
  if (mu < 12.0) {
    double L = exp(-mu), p = 1.0;
    unsigned int k = 0;
    do { k++; p *= gsl_rng_uniform(r); } while (p > L);
    return k - 1;
  }
  const double slam = sqrt(mu), loglam = log(mu), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2);
  for (;;) {
    const double U = gsl_rng_uniform(r) - 0.5, V = gsl_rng_uniform(r);
    const double us = 0.5 - fabs(U);
    const double k = floor((2 * a / us + b) * U + mu + 0.43);
    if (us >= 0.07 && V <= vr) return (unsigned int)k;
    if (k < 0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -mu + k * loglam - lgamma(k + 1)) return (unsigned int)k;
  }
}

gsl_error_handler_t *gsl_set_error_handler_off(void) { return NULL; }

gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n)
{
  gsl_integration_workspace *w = malloc(sizeof(*w));
  w->limit = n;
  return w;
}
void gsl_integration_workspace_free(gsl_integration_workspace *w) { free(w); }
gsl_integration_qawo_table *gsl_integration_qawo_table_alloc(double omega, double L,
                                                             enum gsl_integration_qawo_enum sine, size_t n)
{
  (void)omega; (void)L; (void)sine; (void)n;
  return calloc(1, sizeof(gsl_integration_qawo_table));
}
void gsl_integration_qawo_table_free(gsl_integration_qawo_table *t) { free(t); }

int gsl_integration_qng(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        double *result, double *abserr, size_t *neval)
{
  (void)epsabs;
  double tight = epsrel * 1e-3;
  if (tight < 1e-12) tight = 1e-12;
  *result = oracle_integrate(f->function, f->params, a, b, tight, abserr);
  if (neval) *neval = 0;
  return GSL_SUCCESS;
}

int gsl_integration_qagil(gsl_function *f, double b, double epsabs, double epsrel, size_t limit,
                          gsl_integration_workspace *w, double *result, double *abserr)
{
  (void)epsabs; (void)limit; (void)w;
  double tight = epsrel * 1e-3;
  if (tight < 1e-12) tight = 1e-12;
  *result = oracle_integrate_lower_inf(f->function, f->params, b, tight, abserr);
  return GSL_SUCCESS;
}

int gsl_integration_qawf(gsl_function *f, double a, double epsabs, size_t limit, gsl_integration_workspace *w,
                         gsl_integration_workspace *cw, gsl_integration_qawo_table *wf, double *result,
                         double *abserr)
{
  (void)f; (void)a; (void)epsabs; (void)limit; (void)w; (void)cw; (void)wf; (void)result; (void)abserr;
  fprintf(stderr, "shim_gsl: gsl_integration_qawf is link-only (never reached by GetHI, cosmo.c:271)\n");
  abort();
}
