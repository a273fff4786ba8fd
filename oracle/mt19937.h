/* TEST INFRASTRUCTURE ONLY (oracle).  MT19937 exactly as exposed by GSL's gsl_rng_mt19937, which is
 * the generator the reference draws from (common.c:133-152 init_rng / rng_01).  GSL is an absent
 * third-party dependency (unpinned); the algorithm is Matsumoto & Nishimura's public mt19937ar with
 * the 2002 init_genrand seeding, GSL maps seed 0 to 4357 and returns u32 / 2^32 in [0,1). */
#ifndef ORACLE_MT19937_H
#define ORACLE_MT19937_H
#include <stdint.h>
typedef struct { uint32_t mt[624]; int mti; } oracle_mt19937;
void oracle_mt_seed(oracle_mt19937 *g, uint32_t seed);
uint32_t oracle_mt_u32(oracle_mt19937 *g);
double oracle_mt_uniform(oracle_mt19937 *g); /* u32 / 4294967296.0 */
#endif
