/* TEST INFRASTRUCTURE ONLY -- stand-in for <gsl/gsl_spline.h>: opaque types only (cosmo_mad.h:59-60). */
#ifndef SHIM_GSL_SPLINE_H
#define SHIM_GSL_SPLINE_H
typedef struct shim_gsl_interp_accel gsl_interp_accel;
typedef struct shim_gsl_spline gsl_spline;
#endif
