/* TEST INFRASTRUCTURE ONLY (oracle).  See healpix_ring.h. */
#include <math.h>
#include "healpix_ring.h"

static const double kPi = 3.141592653589793238462643383279502884197;
static const double kTwoPi = 6.283185307179586476925286766559005768394;
static const double kInvHalfPi = 0.6366197723675813430755350534900574;
static const double kTwoThird = 2.0 / 3.0;

long oracle_nside2npix(long nside) { return 12 * nside * nside; }

static double wrap_positive(double v1, double v2)
{
  if (v1 >= 0) return (v1 < v2) ? v1 : fmod(v1, v2);
  double tmp = fmod(v1, v2) + v2;
  return (tmp == v2) ? 0. : tmp;
}

static long imod(long v1, long v2)
{
  long v = v1 % v2;
  return (v >= 0) ? v : v + v2;
}

long oracle_zphi2pix_ring(long nside, double z, double sth, double phi)
{
  double za = fabs(z);
  double tt = wrap_positive(phi, kTwoPi) * kInvHalfPi; /* in [0,4) */

  if (za <= kTwoThird) { /* equatorial belt */
    double t1 = nside * (0.5 + tt);
    double t2 = nside * z * 0.75;
    long jp = (long)(t1 - t2); /* ascending edge line */
    long jm = (long)(t1 + t2); /* descending edge line */
    long ir = nside + 1 + jp - jm; /* ring counted from z=2/3, in 1..2n+1 */
    int kshift = 1 - (int)(ir & 1);
    long ip = (jp + jm - nside + kshift + 1) / 2;
    ip = imod(ip, 4 * nside);
    return nside * (nside - 1) * 2 + (ir - 1) * 4 * nside + ip;
  }
  /* polar caps */
  double tp = tt - (int)(tt);
  double tmp = (sth > -2.) ? nside * sth / sqrt((1. + za) / 3.) : nside * sqrt(3 * (1 - za));
  long jp = (long)(tp * tmp);
  long jm = (long)((1.0 - tp) * tmp);
  long ir = jp + jm + 1; /* ring counted from the nearest pole */
  long ip = (long)(tt * ir);
  ip = imod(ip, 4 * ir);
  if (z > 0) return 2 * ir * (ir - 1) + ip;
  return 12 * nside * nside - 2 * ir * (ir + 1) + ip;
}

long oracle_vec2pix_ring(long nside, const double vec[3])
{
  double vlen = sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
  double cth = vec[2] / vlen;
  double sth = (fabs(cth) > 0.99) ? sqrt(vec[0] * vec[0] + vec[1] * vec[1]) / vlen : -5;
  return oracle_zphi2pix_ring(nside, cth, sth, atan2(vec[1], vec[0]));
}

void oracle_pix2vec_ring(long nside, long ipix, double vec[3])
{
  long npix = 12 * nside * nside, ncap = 2 * nside * (nside - 1);
  double z, phi, fact2 = 4. / npix;
  if (ipix < ncap) { /* north cap */
    long iring = (long)(0.5 * (1 + sqrt(1.5 + 2. * ipix)));
    while (2 * iring * (iring - 1) > ipix) iring--;
    while (2 * iring * (iring + 1) <= ipix) iring++;
    long iphi = ipix + 1 - 2 * iring * (iring - 1);
    z = 1.0 - (double)(iring * iring) * fact2;
    phi = (iphi - 0.5) * kPi / (2. * iring);
  } else if (ipix < npix - ncap) { /* belt */
    double fact1 = (double)(nside << 1) * fact2;
    long ip = ipix - ncap;
    long iring = ip / (4 * nside) + nside;
    long iphi = ip % (4 * nside) + 1;
    double fodd = ((iring + nside) & 1) ? 1 : 0.5;
    z = (double)(2 * nside - iring) * fact1;
    phi = (iphi - fodd) * kPi / (2. * nside);
  } else { /* south cap */
    long ip = npix - ipix;
    long iring = (long)(0.5 * (1 + sqrt(2. * ip - 1)));
    while (2 * iring * (iring - 1) >= ip) iring--;
    while (2 * iring * (iring + 1) < ip) iring++;
    long iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    z = -1.0 + (double)(iring * iring) * fact2;
    phi = (iphi - 0.5) * kPi / (2. * iring);
  }
  double st = sqrt((1. - z) * (1. + z));
  vec[0] = st * cos(phi);
  vec[1] = st * sin(phi);
  vec[2] = z;
}

/* ---- RING <-> NEST ------------------------------------------------------------------------------------- */
static const int kJrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
static const int kJpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};

static long compress_bits(long v) /* every second bit of v, packed */
{
  long r = 0;
  for (int b = 0; b < 31; ++b) r |= ((v >> (2 * b)) & 1L) << b;
  return r;
}
static long spread_bits(long v) /* bit b of v -> bit 2b */
{
  long r = 0;
  for (int b = 0; b < 31; ++b) r |= ((v >> b) & 1L) << (2 * b);
  return r;
}

static long isqrt_long(long v)
{
  long r = (long)sqrt((double)v + 0.5);
  while (r * r > v) --r;
  while ((r + 1) * (r + 1) <= v) ++r;
  return r;
}

long oracle_nest2ring(long nside, long ipnest)
{
  const long npface = nside * nside, npix = 12 * npface, nl4 = 4 * nside, ncap = 2 * nside * (nside - 1);
  const long face = ipnest / npface, ipf = ipnest % npface;
  const long ix = compress_bits(ipf), iy = compress_bits(ipf >> 1);
  const long jr = kJrll[face] * nside - ix - iy - 1;
  long nr = nside, n_before = ncap + nl4 * (jr - nside), kshift = (jr - nside) & 1;
  if (jr < nside) { nr = jr; n_before = 2 * nr * (nr - 1); kshift = 0; }
  else if (jr > 3 * nside) { nr = nl4 - jr; n_before = npix - 2 * (nr + 1) * nr; kshift = 0; }
  long jp = (kJpll[face] * nr + ix - iy + 1 + kshift) / 2;
  if (jp > nl4) jp -= nl4;
  if (jp < 1) jp += nl4;
  return n_before + jp - 1;
}

long oracle_ring2nest(long nside, long ipring)
{
  const long npix = 12 * nside * nside, nl2 = 2 * nside, nl4 = 4 * nside, ncap = 2 * nside * (nside - 1);
  long irn, iphi, nr, kshift, face;
  if (ipring < ncap) { /* north polar cap */
    irn = (1 + isqrt_long(1 + 2 * ipring)) / 2; /* ring counted from the north pole */
    iphi = ipring + 1 - 2 * irn * (irn - 1);
    kshift = 0;
    nr = irn;
    face = (iphi - 1) / nr;
  } else if (ipring < npix - ncap) { /* equatorial belt */
    const long ip = ipring - ncap;
    irn = ip / nl4 + nside;
    iphi = ip % nl4 + 1;
    kshift = (irn + nside) & 1;
    nr = nside;
    const long ire = irn - nside + 1, irm = nl2 + 2 - ire;
    const long ifm = (iphi - ire / 2 + nside - 1) / nside, ifp = (iphi - irm / 2 + nside - 1) / nside;
    if (ifp == ifm) face = (ifp == 4) ? 4 : ifp + 4;
    else if (ifp < ifm) face = ifp;
    else face = ifm + 8;
  } else { /* south polar cap */
    const long ip = npix - ipring;
    const long irs = (1 + isqrt_long(2 * ip - 1)) / 2; /* ring counted from the south pole */
    iphi = 4 * irs + 1 - (ip - 2 * irs * (irs - 1));
    kshift = 0;
    nr = irs;
    irn = nl4 - irs;
    face = (iphi - 1) / nr + 8;
  }
  const long irt = irn - kJrll[face] * nside + 1;
  long ipt = 2 * iphi - kJpll[face] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  const long ix = (ipt - irt) / 2, iy = (-(ipt + irt)) / 2;
  return face * nside * nside + spread_bits(ix) + 2 * spread_bits(iy);
}

void oracle_udgrade(const float *map_in, long nside_in, float *map_out, long nside_out, int nest)
{
  const long npix_in = 12 * nside_in * nside_in, npix_out = 12 * nside_out * nside_out;
  if (nside_in == nside_out) {
    for (long i = 0; i < npix_out; ++i) map_out[i] = map_in[i];
  } else if (nside_in > nside_out) {
    const long ratio = npix_in / npix_out;
    const double inv = 1. / ((double)ratio);
    for (long i = 0; i < npix_out; ++i) {
      double tot = 0;
      const long base = ratio * (nest ? i : oracle_ring2nest(nside_out, i));
      for (long j = 0; j < ratio; ++j) tot += map_in[nest ? base + j : oracle_nest2ring(nside_in, base + j)];
      map_out[i] = tot * inv;
    }
  } else {
    const long ratio = npix_out / npix_in;
    for (long i = 0; i < npix_in; ++i) {
      const float v = map_in[i];
      const long base = ratio * (nest ? i : oracle_ring2nest(nside_in, i));
      for (long j = 0; j < ratio; ++j) map_out[nest ? base + j : oracle_nest2ring(nside_out, base + j)] = v;
    }
  }
}
