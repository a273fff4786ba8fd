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
/* RING <-> NEST (chealpix nest2ring / ring2nest, used by he_udgrade, healpix_extra.c:318-385): the standard HEALPix
 * bijection (face number, (ix, iy) bit interleave, jrll / jpll), restated from the public definition; pinned by the
 * nside = 1, 2 tables of the HEALPix primer and by round trips (tests/test_joint_cpu.py). */
long oracle_nest2ring(long nside, long ipnest);
long oracle_ring2nest(long nside, long ipring);
/* he_udgrade (healpix_extra.c:318-385), float maps: degrade = double sum of the children in NEST order times
 * 1/ratio, upgrade = replication; nest != 0: maps in NEST ordering */
void oracle_udgrade(const float *map_in, long nside_in, float *map_out, long nside_out, int nest);
#endif
