/* TEST INFRASTRUCTURE ONLY -- stand-in for libsharp's <sharp_almhelpers.h> (link-only). */
#ifndef SHIM_SHARP_ALMHELPERS_H
#define SHIM_SHARP_ALMHELPERS_H
#include "sharp.h"
void sharp_make_triangular_alm_info(int lmax, int mmax, int stride, sharp_alm_info **alm_info);
#endif
