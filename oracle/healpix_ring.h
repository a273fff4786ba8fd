/* TEST INFRASTRUCTURE ONLY (oracle).  HEALPix RING-scheme pixelisation as used by the reference
 * through chealpix's vec2pix_ring / nside2npix (pixelize.c:155,174,222).  chealpix is an absent,
 * unpinned third-party dependency ("HEALPix >= 3.10", README.GetHI:48-53); this restates the public
 * HEALPix formulae (Gorski et al. 2005; chealpix.c ang2pix_ring_z_phi, the >=3.30 variant that
 * passes sin(theta) for |cos(theta)|>0.99).  PARITY UNPINNED: no reference test fixes which library
 * generation produced its maps; self-consistency is pinned by pix2vec round trips (tests/). */
#ifndef ORACLE_HEALPIX_RING_H
#define ORACLE_HEALPIX_RING_H
long oracle_nside2npix(long nside);
long oracle_vec2pix_ring(long nside, const double vec[3]);
long oracle_zphi2pix_ring(long nside, double z, double sth, double phi);
void oracle_pix2vec_ring(long nside, long ipix, double vec[3]); /* pixel centre, for KATs */
#endif
