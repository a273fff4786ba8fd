/* TEST INFRASTRUCTURE ONLY -- stand-in for <gsl/gsl_rng.h>, used solely to compile the
 * unmodified reference sources into oracle/_ref/.  Restates GSL's public mt19937 generator
 * (Matsumoto & Nishimura 2002 seeding, gsl seed 0 -> 4357, uniform = u32 / 2^32).
 * GSL itself is absent from this image and from /root/reference (version unpinned by the
 * reference: README.GetHI:48-53 says only "GSL"). */
#ifndef SHIM_GSL_RNG_H
#define SHIM_GSL_RNG_H
typedef struct shim_gsl_rng_type { const char *name; } gsl_rng_type;
typedef struct shim_gsl_rng {
  unsigned long mt[624];
  int mti;
} gsl_rng;
extern const gsl_rng_type *gsl_rng_mt19937;
extern const gsl_rng_type *gsl_rng_ranlux;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
unsigned long gsl_rng_get(gsl_rng *r);
double gsl_rng_uniform(gsl_rng *r);
void gsl_rng_free(gsl_rng *r);
#endif
