/* TEST INFRASTRUCTURE ONLY -- stand-in for <gsl/gsl_randist.h> (link-only for GetHI with do_psources=0). */
#ifndef SHIM_GSL_RANDIST_H
#define SHIM_GSL_RANDIST_H
#include <gsl/gsl_rng.h>
unsigned int gsl_ran_poisson(gsl_rng *r, double mu);
#endif
