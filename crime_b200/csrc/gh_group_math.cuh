// Grouped Taylor pixelisation for mk_T_maps (reference src/pixelize.c:206-227): the arithmetic shared by
// accumulate_kernel (gh_pixelize.cu) and the host-side test harness (tests/native/group_math_host.cpp), which
// runs the very same code on the CPU against the oracle.
//
// A thread owns a 2 x 2 x 2 block of cells.  All 80 sub-particles of the block lie within hg = sqrt(3) dx of the
// block centre C, so the quantities whose floor() gives the HEALPix RING pixel are expanded to second order about C
// once per block:
//   equatorial belt (|cos theta| < 2/3):  U = ns (tt + 1/2) - 1/2 - KA,  V = (3/4) ns z/r;   jp = KA + round(U - V), jm = KA + round(U + V)
//   polar caps      (|cos theta| > 2/3):  U = tt - ntt = tp,             V = ns sqrt(3 (1 - |z|/r));
//                                          jp = round(U V - 1/2), jm = round(V - U V - 1/2), ip = ntt ir + round(U ir - 1/2)
// (tt = azimuth / (pi/2); chealpix ang2pix_ring_z_phi, restated exactly in gh_index_math.cuh).  The constant terms
// come from double-precision atan2 / sqrt of the exact centre, reduced by an integer (KA, ntt) so that the float
// evaluation keeps ~1e-7 absolute accuracy; gradients and Hessians are float.  A sub-particle then costs 5 + 9
// FMAs for (U, V) and 4 for r^2 - |C|^2, with the ten offsets as compile-time-uniform constants.  floor() is taken
// by the 1.5 * 2^23 rounding trick on the FP32 pipe; an answer is accepted only if it is further than the block's
// margin e from the rounding boundary: e = GH_GRP_EPS_IDX * ns (float evaluation) + the third-order remainders
// (bounds fitted in tools/taylor_proto.py and tools/polar_proto.py, checked by tests/test_group_pixelisation_cpu.py).
// Everything else (blocks near the polar axis, astride |cos theta| = 2/3, a quadrant boundary or the tt = 0 seam)
// takes the per-cell fp32 path of gh_pixelize.cu, and unsure sub-particles the exact fp64 path.
#pragma once
#include "gh_index_math.cuh"
#include <string.h>

#define GH_GRP_EPS_IDX 2e-6f   /* pixel-coordinate units per nside: float evaluation of the reduced expansions */
#define GH_GRP_E_MAX 0.02f     /* blocks whose margin would exceed this are not expanded */
#define GH_GRP_MAGIC 12582912.0f
#define GH_GRP_MAGIC_BITS 0x4B400000

#ifdef __CUDA_ARCH__
#define GH_RSQRTF(a) rsqrtf(a)
#define GH_F2I_BITS(f) __float_as_int(f)
#else
#define GH_RSQRTF(a) (1.0f / sqrtf(a))
static inline int gh_f2i_bits_host(float f) { int i; memcpy(&i, &f, 4); return i; }
#define GH_F2I_BITS(f) gh_f2i_bits_host(f)
#endif

enum { GH_GRP_NONE = 0, GH_GRP_EQ = 1, GH_GRP_NORTH = 2, GH_GRP_SOUTH = 3 };

struct GhGroupExp {
  // U(o) = U0 + ox (Ux + Uxx ox + Uxy oy) + oy (Uy - Uxx oy)            (harmonic in x, y)
  float U0, Ux, Uy, Uxx, Uxy;
  // V(o) = V0 + ox (Vx + Vxx ox + Vxy oy + Vxz oz) + oy (Vy + Vyy oy + Vyz oz) + oz (Vz + Vzz oz)
  float V0, Vx, Vy, Vz, Vxx, Vyy, Vzz, Vxy, Vxz, Vyz;
  // S(o) = |C + o|^2 - |C|^2 = ox (Sx + ox) + oy (Sy + oy) + oz (Sz + oz)
  float Sx, Sy, Sz;
  float e;      // margin of every rounding decision, pixel-coordinate units
  int kind;     // GH_GRP_*
  int kbase;    // equatorial: KA; polar: ntt
  double rc;    // |C|
};

// expansions about one cell of the block: offsets are now relative to the cell centre C + (dx, dy, dz)
struct GhCellExp {
  float U0, Ux, Uy, V0, Vx, Vy, Vz, S0, Sx, Sy, Sz;
};

// hg: bound on |o| for every sub-particle of the block (incl. slack); eps_scale scales the float-evaluation margin
// and the regime bounds (the on-device audit runs with 1, 1/2, 1/4)
GH_HD void gh_group_expand(double X, double Y, double Z, float hg, float fns, float eps_scale, GhGroupExp &g)
{
  const float xh = (float)X, yh = (float)Y, zh = (float)Z;
  const float rp2 = fmaf(xh, xh, yh * yh), r2 = fmaf(zh, zh, rp2);
  const float inv_r = GH_RSQRTF(r2), inv_rho = GH_RSQRTF(rp2);
  const float dr = hg * inv_r, drho = hg * inv_rho;
  const float cz = fabsf(zh) * inv_r;
  g.kind = GH_GRP_NONE;
  g.Sx = 2.0f * xh; g.Sy = 2.0f * yh; g.Sz = 2.0f * zh;
  if (!(drho < 0.2f)) return;  // close to the polar axis (also rho = 0 / NaN)
  const float e_tt = 0.2123f * drho * drho * drho;  // remainder of tt, per unit of its multiplier
  const bool eq = cz + dr < (2.0f / 3.0f) - GH_FAST_EPS_CTH * eps_scale;
  const bool pol = cz - dr > (2.0f / 3.0f) + GH_FAST_EPS_CTH * eps_scale;
  const float ax = fabsf(xh), ay = fabsf(yh), hq = 1.01f * hg;
  float e;
  if (eq) {
    // no sub-particle on the tt = 0 / 4 seam (y = 0, x > 0)
    if (!(xh < -hq || ay > hq)) return;
    e = GH_GRP_EPS_IDX * eps_scale * fns + fns * fmaf(0.375f * dr, dr * dr, e_tt);
  } else if (pol) {
    // every sub-particle in the same quadrant
    if (!(ax > hq && ay > hq)) return;
    e = GH_GRP_EPS_IDX * eps_scale * fns + fns * (e_tt + 0.45f * drho * drho * drho * (inv_r / inv_rho) + 0.4f * dr * dr * dr);
  } else {
    return;
  }
  if (!(e < GH_GRP_E_MAX)) return;
  g.e = e;
  // ---- constants in double from the exact centre ----
  const double inv_halfpi = 0.6366197723675813430755350534900574;
  double tt = atan2(Y, X) * inv_halfpi;
  if (tt < 0) tt += 4.0;
  const double R2 = X * X + Y * Y + Z * Z, RC = sqrt(R2);
  g.rc = RC;
  const float irho2 = inv_rho * inv_rho, ir2 = inv_r * inv_r, ir3 = inv_r * ir2, ir5 = ir3 * ir2;
  float k, a, b;
  if (eq) {
    const double A0 = (double)fns * (tt + 0.5) - 0.5;
    const double KA = rint(A0);
    g.kbase = (int)KA;
    g.U0 = (float)(A0 - KA);
    g.V0 = (float)(0.75 * (double)fns * Z / RC);
    k = 0.63661977236758134308f * fns;
    a = 0.75f * fns;
    b = 0.f;
    g.kind = GH_GRP_EQ;
  } else {
    const int ntt = (yh > 0.f) ? (xh > 0.f ? 0 : 1) : (xh > 0.f ? 3 : 2);
    g.kbase = ntt;
    g.U0 = (float)(tt - (double)ntt);
    const double u0 = (X * X + Y * Y) / (RC * (RC + fabs(Z)));  // 1 - |z|/r without the cancellation
    const double su = sqrt(u0), K = 1.7320508075688772935 * (double)fns;
    g.V0 = (float)(K * su);
    k = 0.63661977236758134308f;
    const float sgn = (zh > 0.f) ? 1.0f : -1.0f;
    a = (float)(-0.5 * K / su) * sgn;     // dV = a d(z/r) + b (d(z/r))^2 ...
    b = (float)(-0.25 * K / (u0 * su));
    g.kind = (zh > 0.f) ? GH_GRP_NORTH : GH_GRP_SOUTH;
  }
  // tt-type expansion (gradient and Hessian of atan2(y, x) times k)
  const float kq = k * irho2 * irho2;
  g.Ux = -k * yh * irho2;
  g.Uy = k * xh * irho2;
  g.Uxx = kq * xh * yh;
  g.Uxy = kq * fmaf(yh, yh, -xh * xh);
  // gradient gq and Hessian Hq of z/r
  const float zi3 = zh * ir3, t3 = 3.0f * zh * ir5;
  const float gx = -xh * zi3, gy = -yh * zi3, gz = rp2 * ir3;
  const float Hxx = fmaf(t3 * xh, xh, -zi3), Hyy = fmaf(t3 * yh, yh, -zi3), Hzz = fmaf(t3 * zh, zh, -3.0f * zi3);
  const float Hxy = t3 * xh * yh, Hxz = fmaf(t3 * xh, zh, -xh * ir3), Hyz = fmaf(t3 * yh, zh, -yh * ir3);
  g.Vx = a * gx; g.Vy = a * gy; g.Vz = a * gz;
  g.Vxx = 0.5f * fmaf(a, Hxx, b * gx * gx);
  g.Vyy = 0.5f * fmaf(a, Hyy, b * gy * gy);
  g.Vzz = 0.5f * fmaf(a, Hzz, b * gz * gz);
  g.Vxy = fmaf(a, Hxy, b * gx * gy);
  g.Vxz = fmaf(a, Hxz, b * gx * gz);
  g.Vyz = fmaf(a, Hyz, b * gy * gz);
}

// shift the expansion point by (px, py, pz): exact for second-order polynomials
GH_HD void gh_group_recentre(const GhGroupExp &g, float px, float py, float pz, GhCellExp &c)
{
  c.U0 = fmaf(px, fmaf(g.Uxy, py, fmaf(g.Uxx, px, g.Ux)), fmaf(py, fmaf(-g.Uxx, py, g.Uy), g.U0));
  c.Ux = fmaf(2.0f * g.Uxx, px, fmaf(g.Uxy, py, g.Ux));
  c.Uy = fmaf(-2.0f * g.Uxx, py, fmaf(g.Uxy, px, g.Uy));
  c.V0 = fmaf(px, fmaf(g.Vxz, pz, fmaf(g.Vxy, py, fmaf(g.Vxx, px, g.Vx))),
              fmaf(py, fmaf(g.Vyz, pz, fmaf(g.Vyy, py, g.Vy)), fmaf(pz, fmaf(g.Vzz, pz, g.Vz), g.V0)));
  c.Vx = fmaf(2.0f * g.Vxx, px, fmaf(g.Vxy, py, fmaf(g.Vxz, pz, g.Vx)));
  c.Vy = fmaf(2.0f * g.Vyy, py, fmaf(g.Vxy, px, fmaf(g.Vyz, pz, g.Vy)));
  c.Vz = fmaf(2.0f * g.Vzz, pz, fmaf(g.Vxz, px, fmaf(g.Vyz, py, g.Vz)));
  c.S0 = fmaf(px, g.Sx + px, fmaf(py, g.Sy + py, pz * (g.Sz + pz)));
  c.Sx = fmaf(2.0f, px, g.Sx);
  c.Sy = fmaf(2.0f, py, g.Sy);
  c.Sz = fmaf(2.0f, pz, g.Sz);
}

GH_HD float gh_cell_U(const GhGroupExp &g, const GhCellExp &c, float ox, float oy)
{
  return fmaf(ox, fmaf(g.Uxy, oy, fmaf(g.Uxx, ox, c.Ux)), fmaf(oy, fmaf(-g.Uxx, oy, c.Uy), c.U0));
}
GH_HD float gh_cell_V(const GhGroupExp &g, const GhCellExp &c, float ox, float oy, float oz)
{
  return fmaf(ox, fmaf(g.Vxz, oz, fmaf(g.Vxy, oy, fmaf(g.Vxx, ox, c.Vx))),
              fmaf(oy, fmaf(g.Vyz, oz, fmaf(g.Vyy, oy, c.Vy)), fmaf(oz, fmaf(g.Vzz, oz, c.Vz), c.V0)));
}
// o2 = |o|^2
GH_HD float gh_cell_S(const GhCellExp &c, float ox, float oy, float oz, float o2)
{
  return fmaf(ox, c.Sx, fmaf(oy, c.Sy, fmaf(oz, c.Sz, c.S0 + o2)));
}

// round(v) and its integer as the low bits of v + 1.5 * 2^23; |v| < 2^22
GH_HD float gh_magic_round(float v, int &bits)
{
  const float s = v + GH_GRP_MAGIC;
  bits = GH_F2I_BITS(s);
  return s - GH_GRP_MAGIC;
}

// Equatorial belt.  ok: both floors are clear of their boundaries by more than the margin.
// pix0 = 2 ns (ns - 1) + 4 ns^2, c_sum = 2 KA - ns + 1 - 2 GH_GRP_MAGIC_BITS (wrapping), ns4 = 4 ns, hm = 1/2 - e
GH_HD bool gh_sub_eq(float U, float V, float hm, int ns4, int pix0, int c_sum, int &pix)
{
  const float a = U - V, b = U + V;
  int ba, bb;
  const float ra = gh_magic_round(a, ba), rb = gh_magic_round(b, bb);
  const bool ok = (fabsf(a - ra) < hm) && (fabsf(b - rb) < hm);
  const int jd = ba - bb;                                  // jp - jm
  int ip = (int)((unsigned)ba + (unsigned)bb + (unsigned)c_sum) >> 1;  // (jp + jm - ns + 1) >> 1, see test_taylor_pixelisation_cpu.py
  const unsigned ipw = (unsigned)(ip - ns4);  // wraps to a huge value unless ip >= 4 ns
  ip = (int)((unsigned)ip < ipw ? (unsigned)ip : ipw);
  pix = pix0 + jd * ns4 + ip;
  return ok;
}

// Polar caps.  sg2 = +2 (north) / -2 (south), c_t = ntt - 2, pbase = -GH_GRP_MAGIC_BITS (+ npix in the south)
GH_HD bool gh_sub_polar(float tp, float F, float hm, int sg2, int c_t, int pbase, int &pix)
{
  const float P = fmaf(tp, F, -0.5f);      // tp tmp - 1/2
  const float Q = (F - 1.0f) - P;          // (1 - tp) tmp - 1/2
  int bp, bq, bc;
  const float rp = gh_magic_round(P, bp), rq = gh_magic_round(Q, bq);
  const float irf = (rp + 1.0f) + rq;      // ir = jp + jm + 1, exact in float
  const float C = fmaf(tp, irf, -0.5f);    // (tt - ntt) ir - 1/2
  const float rc = gh_magic_round(C, bc);
  const bool ok = (fabsf(P - rp) < hm) && (fabsf(Q - rq) < hm) && (fabsf(C - rc) < hm);
  const int ir = (int)((unsigned)bp + (unsigned)bq + (1u - 2u * (unsigned)GH_GRP_MAGIC_BITS));
  const int t = sg2 * ir + c_t;
  pix = ir * t + bc + pbase;
  return ok;
}
