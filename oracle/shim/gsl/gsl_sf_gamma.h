/* TEST INFRASTRUCTURE ONLY -- stand-in for <gsl/gsl_sf_gamma.h>: nothing from it is used by GetHI. */
