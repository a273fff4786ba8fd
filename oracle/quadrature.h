/* TEST INFRASTRUCTURE ONLY (oracle).  Adaptive Gauss-Kronrod (7,15) quadrature standing in for the
 * GSL routines the reference's host cosmology calls (gsl_integration_qng: cosmo_mad.c:283,314;
 * gsl_integration_qagil: cosmo.c:273).  Run to tolerances tighter than the callers request. */
#ifndef ORACLE_QUADRATURE_H
#define ORACLE_QUADRATURE_H
typedef double (*oracle_integrand)(double x, void *params);
double oracle_integrate(oracle_integrand f, void *params, double a, double b, double epsrel, double *abserr);
/* integral over (-inf, b] via x = b - (1-t)/t, as QAGIL's documented transformation */
double oracle_integrate_lower_inf(oracle_integrand f, void *params, double b, double epsrel, double *abserr);
#endif
