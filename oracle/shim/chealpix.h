/* TEST INFRASTRUCTURE ONLY -- stand-in for <chealpix.h> (HEALPix C library >= 3.10, unpinned).
 * nside2npix and vec2pix_ring restate the public HEALPix RING-scheme formulae (newer `sth`
 * variant, see oracle/healpix_ring.c); nest2ring / ring2nest (used by JoinT's he_udgrade) forward to the same file. */
#ifndef SHIM_CHEALPIX_H
#define SHIM_CHEALPIX_H
long nside2npix(long nside);
void vec2pix_ring(long nside, const double *vec, long *ipix);
void nest2ring(long nside, long ipnest, long *ipring);
void ring2nest(long nside, long ipring, long *ipnest);
#endif
