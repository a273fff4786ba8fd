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
