// Exact (IEEE double, no FMA contraction) light-cone index arithmetic of mk_T_maps: comoving radius ->
// redshift -> observed frequency -> frequency shell, and direction -> HEALPix RING pixel.
// Follows reference src/pixelize.c:28-55,206-223, src/cosmo.c:52-62 and chealpix's vec2pix_ring
// (public HEALPix formulae; the >=3.30 variant that passes sin(theta) near the poles).
// The translation units that include this header are compiled with -fmad=false so that every product
// and sum rounds exactly as the reference's gcc -O3 (no -march, no -ffast-math) build does.
// __host__ __device__ so the host-side unit test can run the very same code against the oracle.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define GH_HD __host__ __device__ __forceinline__
#else
#define GH_HD static inline
#endif

struct GhIndexTables {
  const double *z_r2z, *r_r2z;
  int nz_tab;
  double glob_idr;
  const double *nu0, *nuf;
  int n_nu, irregular;
  double nu_min, inv_dnu;
  long long nside;
};

// src/cosmo.c:52-62
GH_HD double gh_z_of_r(const GhIndexTables &t, double r)
{
  if (r <= 0) return 0;
  if (r >= t.r_r2z[t.nz_tab - 1]) return t.z_r2z[t.nz_tab - 1];
  const int ir = (int)(r * t.glob_idr);
  return t.z_r2z[ir] + (t.z_r2z[ir + 1] - t.z_r2z[ir]) * (r - t.r_r2z[ir]) * t.glob_idr;
}

// src/pixelize.c:28-55 (irregular table: the walk's result does not depend on where it starts, so start
// from the uniform guess) and src/pixelize.c:216 (regular table: C truncation toward zero)
GH_HD int gh_shell_of_nu(const GhIndexTables &t, double nu)
{
  if (!t.irregular) return (int)(t.inv_dnu * (nu - t.nu_min));
  int inu = (int)(t.inv_dnu * (nu - t.nu_min));
  if (inu < 0) inu = 0;
  if (inu >= t.n_nu) inu = t.n_nu - 1;
  for (;;) {
    if (inu == -1 || inu == t.n_nu) return inu;
    if (nu < t.nu0[inu]) inu--;
    else if (nu >= t.nuf[inu]) inu++;
    else return inu;
  }
}

// chealpix vec2pix_ring -> ang2pix_ring_z_phi; vlen = sqrt(x*x+y*y+z*z) is passed in because mk_T_maps
// has just computed the identical expression as the comoving radius (src/pixelize.c:210).
GH_HD long long gh_vec2pix_ring(long long nside, double x, double y, double z, double vlen)
{
  const double twopi = 6.283185307179586476925286766559005768394;
  const double inv_halfpi = 0.6366197723675813430755350534900574;
  const double cth = z / vlen;
  const double za = fabs(cth);
  const double phi = atan2(y, x);
  // fmodulo(phi, 2 pi) for phi in [-pi, pi]
  double ph = phi;
  if (phi < 0) {
    const double tmp = phi + twopi;
    ph = (tmp == twopi) ? 0. : tmp;
  }
  const double tt = ph * inv_halfpi;  // [0,4)
  if (za <= 2.0 / 3.0) {
    const double t1 = nside * (0.5 + tt);
    const double t2 = nside * cth * 0.75;
    const long long jp = (long long)(t1 - t2);
    const long long jm = (long long)(t1 + t2);
    const long long ir = nside + 1 + jp - jm;
    const int kshift = 1 - (int)(ir & 1);
    long long ip = (jp + jm - nside + kshift + 1) / 2;
    ip %= 4 * nside;
    if (ip < 0) ip += 4 * nside;
    return nside * (nside - 1) * 2 + (ir - 1) * 4 * nside + ip;
  }
  const double tp = tt - (int)(tt);
  double tmp;
  if (za > 0.99) {
    const double sth = sqrt(x * x + y * y) / vlen;
    tmp = nside * sth / sqrt((1. + za) / 3.);
  } else {
    tmp = nside * sqrt(3 * (1 - za));
  }
  const long long jp = (long long)(tp * tmp);
  const long long jm = (long long)((1.0 - tp) * tmp);
  const long long ir = jp + jm + 1;
  long long ip = (long long)(tt * ir);
  ip %= 4 * ir;
  if (ip < 0) ip += 4 * ir;
  return (cth > 0) ? 2 * ir * (ir - 1) + ip : 12 * nside * nside - 2 * ir * (ir + 1) + ip;
}

// inner body of src/pixelize.c:206-223 for one sub-particle; returns the shell (or -1 / n_nu)
GH_HD int gh_point_to_shell_pixel(const GhIndexTables &t, double x, double y, double z, double dz_rsd, long long *ipix)
{
  const double r = sqrt(x * x + y * y + z * z);
  const double redshift = gh_z_of_r(t, r) + dz_rsd;
  const double nu = 1420.40575177 / (1 + redshift);
  const int inu = gh_shell_of_nu(t, nu);
  *ipix = -1;
  if (inu >= 0 && inu < t.n_nu) *ipix = gh_vec2pix_ring(t.nside, x, y, z, r);
  return inu;
}

// ================================================================================================
// fp32 fast path with a guaranteed-safe fallback.
//
// The exact path above costs ~350 instructions per sub-particle (fp64 sqrt, divisions, atan2).  The fast
// path evaluates the same quantities in fp32 and accepts its own answer only when every floor() /
// comparison it took is further from its decision boundary than a conservative bound on the fp32-vs-fp64
// discrepancy of that quantity; otherwise the sub-particle is re-done with the exact path.  Accepted
// answers are therefore identical to the exact path's.  Bounds (derivation in DESIGN.md; audited on
// device against the exact path by gh_cuda_fastpath_audit, which also reports the largest observed
// discrepancy/bound ratio):
//   observed frequency      : GH_FAST_EPS_NU  MHz                    (est. <= 6e-4)
//   pixel-index coordinates : GH_FAST_EPS_IDX * nside                (est. <= 1.2e-6 * nside)
//   |cos(theta)| vs 2/3     : GH_FAST_EPS_CTH                        (est. <= 4e-7)
//   phi * 2/pi vs integers  : GH_FAST_EPS_TT                         (est. <= 1e-6)
//   radius vs a shell edge  : GH_FAST_EPS_R  Mpc/h                   (est. <= 3e-3)
#define GH_FAST_EPS_NU 4e-3f
#define GH_FAST_EPS_IDX 6e-6f
#define GH_FAST_EPS_CTH 3e-6f
#define GH_FAST_EPS_TT 6e-6f
#define GH_FAST_EPS_R 2e-2f   /* Mpc/h: shell-edge radii of a cell (est. discrepancy <= 3e-3) */

enum { GH_FAST_OUT = 0, GH_FAST_IN = 1, GH_FAST_UNSURE = 2 };
