// Philox4x32-10 (Salmon et al. 2011; Random123 known answers in tests/): the counter-based generator behind the
// k-space realisation (gh_kgen.cu) and the point-source sampling (gh_psources.cu).  Counter words c2 = c3 = 0.
#pragma once
#include <stdint.h>

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t &o0,
                                              uint32_t &o1, uint32_t &o2, uint32_t &o3, uint32_t c2 = 0u, uint32_t c3 = 0u)
{
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}
