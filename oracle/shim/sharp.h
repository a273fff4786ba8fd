/* TEST INFRASTRUCTURE ONLY -- stand-in for libsharp's <sharp.h>: link-only for GetHI (spherical
 * harmonic transforms are used by ForGet/JoinT only); every entry point aborts. */
#ifndef SHIM_SHARP_H
#define SHIM_SHARP_H
typedef struct shim_sharp_alm_info sharp_alm_info;
typedef struct shim_sharp_geom_info sharp_geom_info;
typedef enum { SHARP_YtW = 0, SHARP_MAP2ALM = SHARP_YtW, SHARP_Y = 1, SHARP_ALM2MAP = SHARP_Y } sharp_jobtype;
#define SHARP_DP (1 << 4)
void sharp_execute(sharp_jobtype type, int spin, void *alm, void *map, const sharp_geom_info *geom_info,
                   const sharp_alm_info *alm_info, int ntrans, int flags, double *time, unsigned long long *opcnt);
void sharp_destroy_alm_info(sharp_alm_info *info);
void sharp_destroy_geom_info(sharp_geom_info *info);
#endif
