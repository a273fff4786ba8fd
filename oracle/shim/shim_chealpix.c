/* TEST INFRASTRUCTURE ONLY -- chealpix entry points over oracle/healpix_ring.c. */
#include <stdio.h>
#include <stdlib.h>
#include <chealpix.h>
#include "../healpix_ring.h"

long nside2npix(long nside) { return oracle_nside2npix(nside); }
void vec2pix_ring(long nside, const double *vec, long *ipix) { *ipix = oracle_vec2pix_ring(nside, vec); }
void nest2ring(long nside, long ipnest, long *ipring) { *ipring = oracle_nest2ring(nside, ipnest); }
void ring2nest(long nside, long ipring, long *ipnest) { *ipnest = oracle_ring2nest(nside, ipring); }
