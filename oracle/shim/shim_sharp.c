/* TEST INFRASTRUCTURE ONLY -- libsharp entry points: link-only for GetHI, abort if ever called. */
#include <stdio.h>
#include <stdlib.h>
#include "sharp.h"
#include "sharp_almhelpers.h"
#include "sharp_geomhelpers.h"
static void nope(const char *w) { fprintf(stderr, "shim_sharp: %s is link-only for GetHI\n", w); abort(); }
void sharp_execute(sharp_jobtype type, int spin, void *alm, void *map, const sharp_geom_info *geom_info,
                   const sharp_alm_info *alm_info, int ntrans, int flags, double *time, unsigned long long *opcnt)
{
  (void)type; (void)spin; (void)alm; (void)map; (void)geom_info; (void)alm_info; (void)ntrans; (void)flags;
  (void)time; (void)opcnt;
  nope("sharp_execute");
}
void sharp_destroy_alm_info(sharp_alm_info *info) { (void)info; nope("sharp_destroy_alm_info"); }
void sharp_destroy_geom_info(sharp_geom_info *info) { (void)info; nope("sharp_destroy_geom_info"); }
void sharp_make_triangular_alm_info(int lmax, int mmax, int stride, sharp_alm_info **alm_info)
{
  (void)lmax; (void)mmax; (void)stride; (void)alm_info;
  nope("sharp_make_triangular_alm_info");
}
void sharp_make_weighted_healpix_geom_info(int nside, int stride, const double *weight, sharp_geom_info **geom_info)
{
  (void)nside; (void)stride; (void)weight; (void)geom_info;
  nope("sharp_make_weighted_healpix_geom_info");
}
