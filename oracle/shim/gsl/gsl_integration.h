/* TEST INFRASTRUCTURE ONLY -- stand-in for <gsl/gsl_integration.h>.  qng / qagil are replaced by an
 * adaptive Gauss-Kronrod (7,15) integrator run to a tolerance tighter than the caller asks for, so
 * results agree with real GSL to well inside the reference's own requested accuracy (1e-6 / 1e-4).
 * qawf / qawo are link-only: the reference never reaches them on the GetHI path (cosmo.c:271). */
#ifndef SHIM_GSL_INTEGRATION_H
#define SHIM_GSL_INTEGRATION_H
#include <stddef.h>
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)
typedef struct { size_t limit; } gsl_integration_workspace;
typedef struct { int dummy; } gsl_integration_qawo_table;
enum gsl_integration_qawo_enum { GSL_INTEG_COSINE, GSL_INTEG_SINE };
gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
gsl_integration_qawo_table *gsl_integration_qawo_table_alloc(double omega, double L,
                                                             enum gsl_integration_qawo_enum sine, size_t n);
void gsl_integration_qawo_table_free(gsl_integration_qawo_table *t);
int gsl_integration_qng(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        double *result, double *abserr, size_t *neval);
int gsl_integration_qagil(gsl_function *f, double b, double epsabs, double epsrel, size_t limit,
                          gsl_integration_workspace *w, double *result, double *abserr);
int gsl_integration_qawf(gsl_function *f, double a, double epsabs, size_t limit,
                         gsl_integration_workspace *w, gsl_integration_workspace *cw,
                         gsl_integration_qawo_table *wf, double *result, double *abserr);
#endif
