/* TEST INFRASTRUCTURE ONLY (oracle).  See mt19937.h. */
#include "mt19937.h"

void oracle_mt_seed(oracle_mt19937 *g, uint32_t seed)
{
  if (seed == 0) seed = 4357; /* GSL convention */
  g->mt[0] = seed;
  for (int i = 1; i < 624; i++)
    g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->mti = 624;
}

uint32_t oracle_mt_u32(oracle_mt19937 *g)
{
  if (g->mti >= 624) {
    uint32_t *mt = g->mt;
    for (int kk = 0; kk < 624; kk++) {
      uint32_t y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
      mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->mti = 0;
  }
  uint32_t k = g->mt[g->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680u;
  k ^= (k << 15) & 0xefc60000u;
  k ^= (k >> 18);
  return k;
}

double oracle_mt_uniform(oracle_mt19937 *g) { return oracle_mt_u32(g) / 4294967296.0; }
