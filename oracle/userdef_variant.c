/* TEST INFRASTRUCTURE ONLY.  Stands in for the reference's src/user_defined.c -- the file GetHI users are told to
 * edit -- with a different HI model of the same functional family, to check that such an edit reaches the GPU path
 * (oracle/Makefile builds _ref/libgethi_ref_userdef.so = the reference sources + this file). */
#include <math.h>

double fraction_HI(double z) { return 0.012 * pow(1 + z, 0.3); }

double bias_HI(double z) { return 1.1 + 0.07 * pow(1 + z, 2.1); }
