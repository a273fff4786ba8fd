/* TEST INFRASTRUCTURE ONLY -- stand-in for libsharp's <sharp_geomhelpers.h> (link-only). */
#ifndef SHIM_SHARP_GEOMHELPERS_H
#define SHIM_SHARP_GEOMHELPERS_H
#include "sharp.h"
void sharp_make_weighted_healpix_geom_info(int nside, int stride, const double *weight, sharp_geom_info **geom_info);
#endif
