// mk_T_maps on sm_100a (reference src/pixelize.c:150-262): every cell is split into 10 sub-particles at
// fixed offsets, each is sent to (frequency shell, HEALPix RING pixel) and deposits a tenth of the cell's
// HI mass.  Shell and pixel indices must equal the reference's bit for bit: the exact arithmetic (IEEE double,
// this translation unit is built with -fmad=false) lives in gh_index_math.cuh; two fp32 fast paths in front of
// it -- the 2x2x2-block expansion of gh_group_math.cuh and, for the few cells it does not cover, a per-cell
// path -- accept their own answer only when every decision is clear of its error bound, and hand the rest to
// the exact path.
//
// The stage is bound by instruction issue and L2-atomic throughput, not by HBM (8 B/cell of grid traffic
// against ~60 instructions per sub-particle); blocks whose whole extent lies outside the shells' redshift
// window are culled by a conservative radial bound first (about 44 % of the box for the shipped frequency
// table).  DESIGN.md section 4 has the measurements.
#include "gh_internal.cuh"
#include "gh_index_math.cuh"
#include "gh_group_math.cuh"

namespace {

__device__ __forceinline__ GhIndexTables tables_of(const GhDev &d)
{
  GhIndexTables t;
  t.z_r2z = d.z_r2z; t.r_r2z = d.r_r2z; t.nz_tab = d.nz_tab; t.glob_idr = d.glob_idr;
  t.nu0 = d.nu0; t.nuf = d.nuf; t.n_nu = d.n_nu; t.irregular = d.irregular;
  t.nu_min = d.nu_min; t.inv_dnu = d.inv_dnu; t.nside = d.nside;
  return t;
}

// ------------------------------------------------------------------------------------------------
// fp32 fast path (device only, hand-trimmed: this loop is instruction-issue bound).  It evaluates the
// same quantities as gh_point_to_shell_pixel in fp32 and accepts its own answer only when every
// floor() / comparison it took is further from its decision boundary than a conservative bound on the
// fp32-vs-fp64 discrepancy of that quantity (GH_FAST_EPS_*); otherwise the sub-particle is re-done with
// the exact path.  Accepted answers are therefore identical to the exact path's; the *_audit entry points
// measure that on the device, also with the bounds scaled down.
__device__ __forceinline__ float rsqrt_ftz(float a)
{
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float rcp_ftz(float a)
{
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}

struct FastCtx {
  const float *ztab, *edges;
  float idr, r_tab_max, z_last, nu_min, inv_dnu;
  float eps_nu, cth_lo, cth_hi, eps_tt, fns, eidx;
  float nu_out_lo, nu_out_hi;  // certainly outside every shell below / above these
  int ns, n_nu, ir_max;
  const float *rtab;           // r(z), uniform in z
  float inv_dz, z_tab_max, r_last, eps_r;
  int iz_max;
};

__device__ __forceinline__ FastCtx fast_ctx_of(const GhDev &d, float eps_scale)
{
  FastCtx f;
  f.ztab = d.z_r2z_f; f.edges = d.nu_edges_f;
  f.idr = (float)d.glob_idr; f.r_tab_max = (float)d.r_tab_max; f.z_last = __ldg(d.z_r2z_f + d.nz_tab - 1);
  f.nu_min = (float)d.nu_min; f.inv_dnu = (float)d.inv_dnu;
  f.eps_nu = GH_FAST_EPS_NU * eps_scale;
  f.cth_lo = (2.0f / 3.0f) - GH_FAST_EPS_CTH * eps_scale; f.cth_hi = (2.0f / 3.0f) + GH_FAST_EPS_CTH * eps_scale;
  f.eps_tt = GH_FAST_EPS_TT * eps_scale;
  f.ns = (int)d.nside; f.fns = (float)f.ns; f.eidx = GH_FAST_EPS_IDX * eps_scale * f.fns;
  f.n_nu = d.n_nu; f.ir_max = d.nz_tab - 2;
  f.nu_out_lo = __ldg(d.nu_edges_f) - f.eps_nu; f.nu_out_hi = __ldg(d.nu_edges_f + d.n_nu) + f.eps_nu;
  f.rtab = d.r_z2r_f; f.inv_dz = d.inv_dz_tab; f.z_tab_max = d.z_tab_max; f.r_last = __ldg(d.r_z2r_f + d.nz_tab - 1);
  f.eps_r = GH_FAST_EPS_R * eps_scale; f.iz_max = d.nz_tab - 2;
  return f;
}

// float z_of_r (src/cosmo.c:52-62)
__device__ __forceinline__ float z_of_r_f(const FastCtx &f, float r)
{
  const float s = fmaxf(r, 0.f) * f.idr;
  const int ir = min((int)s, f.ir_max);
  const float a = __ldg(f.ztab + ir), b = __ldg(f.ztab + ir + 1);
  const float zr = fmaf(b - a, s - (float)ir, a);
  return (r >= f.r_tab_max) ? f.z_last : zr;
}

// float r_of_z (src/cosmo.c:40-50)
__device__ __forceinline__ float r_of_z_f(const FastCtx &f, float z)
{
  const float s = fmaxf(z, 0.f) * f.inv_dz;
  const int iz = min((int)s, f.iz_max);
  const float a = __ldg(f.rtab + iz), b = __ldg(f.rtab + iz + 1);
  const float r = fmaf(b - a, s - (float)iz, a);
  return (z >= f.z_tab_max) ? f.r_last : r;
}

// Shell boundaries of one cell, as squared radii.  Every sub-particle of the cell has the same Delta z_RSD, so
// a shell edge nu_j is crossed at one radius r_j(dz) for the whole cell: r < r_j <=> nu > nu_j (z_of_r is
// monotone).  The cull's redshift bracket bounds which edges any of the cell's sub-particles can reach; with at
// most two of them the per-sub-particle shell is two comparisons of r^2 instead of a table look-up, a division
// and an edge search.  shell[] = shell below the inner edge, between the two, beyond the outer one (-1 = outside
// every shell); within eps_r of an edge the sub-particle is unsure.  ok = false: more than two edges in reach
// (shells thinner than cells) -> the caller uses the per-sub-particle frequency path.
struct CellShells {
  float lo_a, hi_a, lo_b, hi_b;
  int shell[3];
  int j_in;  // shell[k] == shell_or_out(j_in - k): the shells in reach are consecutive, decreasing with radius
  bool ok;
};

__device__ __forceinline__ CellShells cell_shells(const FastCtx &f, float zs_lo, float zs_hi, float dz)
{
  CellShells c;
  c.ok = false;
  c.lo_a = c.hi_a = c.lo_b = c.hi_b = 3.0e38f;
  c.shell[0] = c.shell[1] = c.shell[2] = -1;
  c.j_in = -1;
  const float nu_hi = 1420.40575177f * rcp_ftz(1.0f + zs_lo), nu_lo = 1420.40575177f * rcp_ftz(1.0f + zs_hi);
  // j_min = first edge >= nu_lo
  int j = (int)ceilf((nu_lo - f.nu_min) * f.inv_dnu);
  j = max(0, min(j, f.n_nu));
  for (int it = 0; it < 4 && j > 0 && __ldg(f.edges + j - 1) >= nu_lo; ++it) --j;
  for (int it = 0; it < 4 && j <= f.n_nu && __ldg(f.edges + j) < nu_lo; ++it) ++j;
  if (j > 0 && __ldg(f.edges + j - 1) >= nu_lo) return c;        // search did not converge (very uneven table)
  if (j <= f.n_nu && __ldg(f.edges + j) < nu_lo) return c;
  // edges in reach: j, j+1, ... while <= nu_hi
  int ne = 0;
  while (j + ne <= f.n_nu && __ldg(f.edges + j + ne) <= nu_hi) {
    if (++ne > 2) return c;
  }
  // shell just below edge j is j-1 (or outside when j == 0); nu decreases with r, so the innermost radii see
  // the highest shell
  auto shell_or_out = [&](int g) { return (g >= 0 && g < f.n_nu) ? g : -1; };
  if (ne == 0) {
    c.shell[0] = shell_or_out(j - 1);
    c.j_in = j - 1;
  } else {
    const int j_in = j + ne - 1;  // highest-frequency edge in reach = smallest radius
    c.j_in = j_in;
    {
      const float r = r_of_z_f(f, 1420.40575177f * rcp_ftz(__ldg(f.edges + j_in)) - 1.0f - dz);
      const float a = fmaxf(r - f.eps_r, 0.f), b = r + f.eps_r;
      c.lo_a = a * a; c.hi_a = b * b;
    }
    c.shell[0] = shell_or_out(j_in);       // r < r_edge: nu >= edge j_in
    c.shell[1] = shell_or_out(j_in - 1);
    if (ne == 2) {
      const float r = r_of_z_f(f, 1420.40575177f * rcp_ftz(__ldg(f.edges + j)) - 1.0f - dz);
      const float a = fmaxf(r - f.eps_r, 0.f), b = r + f.eps_r;
      c.lo_b = a * a; c.hi_b = b * b;
      c.shell[2] = shell_or_out(j - 1);
    }
  }
  c.ok = true;
  return c;
}

// shell of a frequency: GH_FAST_IN (sure, shell in inu), GH_FAST_OUT (surely outside), GH_FAST_UNSURE
__device__ __forceinline__ int fast_shell(const FastCtx &f, float nu, int &inu)
{
  int g = (int)((nu - f.nu_min) * f.inv_dnu);
  g = max(0, min(g, f.n_nu - 1));
  const float lo = __ldg(f.edges + g), hi = __ldg(f.edges + g + 1);
  inu = g;
  if (nu > lo + f.eps_nu && nu < hi - f.eps_nu) return GH_FAST_IN;
  if (nu < f.nu_out_lo || nu > f.nu_out_hi) return GH_FAST_OUT;
  return GH_FAST_UNSURE;  // next to an edge (or a non-uniform table where the uniform guess is off)
}

// RING pixel from cth = z/r, tt = azimuth/(pi/2) in [0,4], and (only when |cth| > 0.99) x^2+y^2 and 1/r
__device__ __forceinline__ bool fast_pixel(const FastCtx &f, float cth, float tt, float q, float inv_r, int &pix)
{
  const float za = fabsf(cth);
  const int ns = f.ns;
  const float lo = f.eidx, hi = 1.0f - f.eidx;
  if (za < f.cth_lo) {
    const float t1 = fmaf(f.fns, tt, 0.5f * f.fns), t2 = (0.75f * f.fns) * cth;
    const float a = t1 - t2, b = t1 + t2;
    const float fa = floorf(a), fb = floorf(b);
    const float ra = a - fa, rb = b - fb;
    const bool ok = (ra > lo) & (ra < hi) & (rb > lo) & (rb < hi) & (tt > f.eps_tt) & (tt < 4.0f - f.eps_tt);
    const int jp = (int)fa, jm = (int)fb;
    const int ir = ns + 1 + jp - jm;
    int ip = (jp + jm - ns + 2 - (ir & 1)) >> 1;  // (jp+jm-ns+kshift+1)/2, operand >= 0 for tt >= eps_tt
    ip -= (ip >= 4 * ns) ? 4 * ns : 0;
    pix = 2 * ns * (ns - 1) + (ir - 1) * 4 * ns + ip;
    return ok;
  }
  if (za > f.cth_hi) {
    const float ft = floorf(tt);
    const float tp = tt - ft;
    float tmp;
    if (za > 0.99f) {
      tmp = f.fns * (q * rsqrt_ftz(fmaxf(q, 1e-30f)) * inv_r) * rsqrt_ftz((1.0f + za) * (1.0f / 3.0f));
    } else {
      const float u = 3.0f * (1.0f - za);
      tmp = f.fns * (u * rsqrt_ftz(u));
    }
    const float a = tp * tmp, b = (1.0f - tp) * tmp;
    const float fa = floorf(a), fb = floorf(b);
    const float ra = a - fa, rb = b - fb;
    const int jp = (int)fa, jm = (int)fb;
    const int ir = jp + jm + 1;
    const float c = tt * (float)ir;
    const float fc = floorf(c);
    const float rc = c - fc;
    const bool ok = (tp > f.eps_tt) & (tp < 1.0f - f.eps_tt) & (ft < 3.5f) & (ra > lo) & (ra < hi) & (rb > lo) & (rb < hi) &
                    (rc > lo) & (rc < hi);
    int ip = (int)fc;
    ip -= (ip >= 4 * ir) ? 4 * ir : 0;
    pix = (cth > 0.f) ? 2 * ir * (ir - 1) + ip : 12 * ns * ns - 2 * ir * (ir + 1) + ip;
    return ok;
  }
  return false;
}

// ---- shells of a 2 x 2 x 2 block of cells ---------------------------------------------------------------
// The cull's redshift bracket of the block bounds which shell edges any of its sub-particles can reach (at most
// three, else the block takes the per-cell path).  A sub-particle of a cell with Delta z_RSD = dz is beyond the
// edge at redshift ze iff r > r_of_z(ze - dz) (z_of_r is monotone).  r_of_z is piecewise linear, so per edge the
// block keeps r_k = r_of_z(ze_k - dz_mid) and the slope s_k of that table interval; a cell's threshold is
// r_k - (dz - dz_mid) s_k, exact within the interval and off by at most rz_slope_var |dz - dz_mid| beyond it, which
// goes into the cell's margin.  k = 0..2 by increasing radius, unused ones at 1e30 (never crossed).
// j_in = shell inside the innermost edge in reach (may be outside the table: -1 or n_nu).
struct GroupShells {
  bool ok;
  int j_in;
  float r0, r1, r2, s0, s1, s2;
};

__device__ __forceinline__ void edge_radius(const FastCtx &f, float z, float &r, float &slope, bool &ok)
{
  const float s = z * f.inv_dz;
  const int iz = (int)s;
  ok = ok && (z > 0.f) && (iz >= 1) && (iz < f.iz_max - 1);
  const int i = max(0, min(iz, f.iz_max));
  const float a = __ldg(f.rtab + i), b = __ldg(f.rtab + i + 1);
  r = fmaf(b - a, s - (float)i, a);
  slope = (b - a) * f.inv_dz;
}

__device__ __forceinline__ GroupShells group_shells(const FastCtx &f, float zs_lo, float zs_hi, float dz_mid)
{
  GroupShells g;
  g.ok = false;
  g.j_in = -1;
  g.r0 = g.r1 = g.r2 = 1.0e30f;
  g.s0 = g.s1 = g.s2 = 0.f;
  const float nu_hi = 1420.40575177f * rcp_ftz(1.0f + zs_lo), nu_lo = 1420.40575177f * rcp_ftz(1.0f + zs_hi);
  int j = (int)ceilf((nu_lo - f.nu_min) * f.inv_dnu);  // first edge >= nu_lo
  j = max(0, min(j, f.n_nu));
  for (int it = 0; it < 4 && j > 0 && __ldg(f.edges + j - 1) >= nu_lo; ++it) --j;
  for (int it = 0; it < 4 && j <= f.n_nu && __ldg(f.edges + j) < nu_lo; ++it) ++j;
  if (j > 0 && __ldg(f.edges + j - 1) >= nu_lo) return g;  // search did not converge (very uneven table)
  if (j <= f.n_nu && __ldg(f.edges + j) < nu_lo) return g;
  int ne = 0;
  while (j + ne <= f.n_nu && __ldg(f.edges + j + ne) <= nu_hi) {
    if (++ne > 3) return g;
  }
  g.j_in = j + ne - 1;  // ne == 0: the shell below edge j
  bool ok = true;
  if (ne > 0) edge_radius(f, 1420.40575177f * rcp_ftz(__ldg(f.edges + g.j_in)) - 1.0f - dz_mid, g.r0, g.s0, ok);
  if (ne > 1) edge_radius(f, 1420.40575177f * rcp_ftz(__ldg(f.edges + g.j_in - 1)) - 1.0f - dz_mid, g.r1, g.s1, ok);
  if (ne > 2) edge_radius(f, 1420.40575177f * rcp_ftz(__ldg(f.edges + g.j_in - 2)) - 1.0f - dz_mid, g.r2, g.s2, ok);
  g.ok = ok;
  return g;
}

// (lo, hi) thresholds on S = r^2 - |C|^2 (C = block centre, |C| = rc_hi + rc_lo) for the edge at radius r +- eps
__device__ __forceinline__ void edge_thresholds(float r, float eps, float rc_hi, float rc_lo, float &lo, float &hi)
{
  const float a = fmaxf(r - eps, 0.f), b = r + eps;
  lo = ((a - rc_hi) - rc_lo) * (a + rc_hi);
  hi = ((b - rc_hi) - rc_lo) * (b + rc_hi);
}

struct AuditCounts {
  unsigned long long out, in, unsure, wrong;
};

// ---- per-cell fp32 path ---------------------------------------------------------------------------------
// For the cells the block expansion does not cover (blocks near the polar axis, astride |cos theta| = 2/3, a
// quadrant boundary or the tt = 0 seam; a few per cent): each of the 10 sub-particles goes through the fp32 fast
// path, which either proves it misses every shell, or proves (shell, pixel) with all decisions clear of their
// error bounds and deposits at once, or marks it unsure (bit in the returned mask).  The azimuth is atan2 of the
// cell centre plus the small rotation to the sub-particle (series in the tangent of the rotation angle; cells
// close to the polar axis use atan2f per sub-particle).
template <bool AUDIT>
__device__ __noinline__ unsigned generic_cell(const GhDev &d, float eps_scale, double x0, double y0, double z0, float w, float dzf,
                                              float *__restrict__ maps, AuditCounts *acp)
{
  // rebuilt from the (constant-bank) parameter block instead of being passed in: rarely executed, and passing them by
  // reference would put the caller's copies on the local-memory stack
  const FastCtx f = fast_ctx_of(d, eps_scale);
  const GhIndexTables t = tables_of(d);
  unsigned need = 0u;
  // cell centre as hi + lo floats: positions are xh + (xl + offset), one rounding of the full coordinate
  const float xh = (float)x0, yh = (float)y0, zh = (float)z0;
  const float xl = (float)(x0 - (double)xh), yl = (float)(y0 - (double)yh), zl = (float)(z0 - (double)zh);
  const float rp2 = fmaf(xh, xh, yh * yh);
  const float rc = sqrtf(fmaf(zh, zh, rp2));
  // conservative cull: all sub-particles lie within half a cell diagonal of the centre and z_of_r is
  // non-decreasing, so their redshifts lie in [z(rc-h), z(rc+h)] + dz; 1e-3 Mpc/h and 1e-5 in z cover
  // the fp32 evaluation of this test
  const float h = (float)d.dx * 0.8660254f + 1e-3f + 1e-6f * rc;
  const float zs_hi = z_of_r_f(f, rc + h) + dzf + 1e-5f, zs_lo = z_of_r_f(f, rc - h) + dzf - 1e-5f;
  const bool culled = (zs_hi < (float)d.z_lo_cull || zs_lo > (float)d.z_hi_cull);
  if (culled) {
    if (AUDIT) {
      for (int isub = 0; isub < GH_CUDA_N_SUBPART; ++isub) {
        long long pe;
        gh_point_to_shell_pixel(t, x0 + d.sub_off[isub], y0 + d.sub_off[GH_CUDA_N_SUBPART + isub],
                                z0 + d.sub_off[2 * GH_CUDA_N_SUBPART + isub], (double)dzf, &pe);
        acp->out++;
        if (pe >= 0) acp->wrong++;
      }
    }
    return 0u;
  }
  // azimuth of the cell centre; sub-particles rotate it by atan(cross/dot), |cross/dot| < 0.05 when
  // the cell is further than 24 cells from the polar axis
  const bool series = rp2 > 576.0f * (float)(d.dx * d.dx);
  const float phi_c = atan2f(yh, xh);
  const CellShells cs = cell_shells(f, zs_lo, zs_hi, dzf);
#pragma unroll 2
  for (int isub = 0; isub < GH_CUDA_N_SUBPART; ++isub) {
    const float ox = xl + d.sub_off_f[isub], oy = yl + d.sub_off_f[GH_CUDA_N_SUBPART + isub];
    const float x = xh + ox, y = yh + oy, z = zh + (zl + d.sub_off_f[2 * GH_CUDA_N_SUBPART + isub]);
    const float q = fmaf(x, x, y * y);
    const float r2 = fmaf(z, z, q);
    const float inv_r = rsqrt_ftz(r2);
    int inu, pix = -1, st;
    if (cs.ok) {
      // two comparisons of r^2 against the cell's shell-edge radii
      const bool in0 = r2 < cs.lo_a, in1 = (r2 > cs.hi_a) & (r2 < cs.lo_b), in2 = r2 > cs.hi_b;
      inu = in0 ? cs.shell[0] : (in1 ? cs.shell[1] : cs.shell[2]);
      st = (in0 | in1 | in2) ? (inu >= 0 ? GH_FAST_IN : GH_FAST_OUT) : GH_FAST_UNSURE;
    } else {
      const float nu = 1420.40575177f * rcp_ftz(1.0f + (z_of_r_f(f, r2 * inv_r) + dzf));
      st = fast_shell(f, nu, inu);
    }
    if (st == GH_FAST_IN) {
      float phi;
      if (series) {
        const float tq = fmaf(xh, oy, -yh * ox) * rcp_ftz(rp2 + fmaf(xh, ox, yh * oy));
        const float t2 = tq * tq;
        phi = fmaf(tq, fmaf(t2, fmaf(t2, 0.2f, -0.33333333f), 1.0f), phi_c);
      } else {
        phi = atan2f(y, x);
      }
      float tt = phi * 0.63661977236758134308f;
      tt += (tt < 0.f) ? 4.0f : 0.f;
      if (!fast_pixel(f, z * inv_r, tt, q, inv_r, pix)) st = GH_FAST_UNSURE;
    }
    if (!AUDIT) {
      if (st == GH_FAST_IN) atomicAdd(maps + ((size_t)d.npix * inu + pix), w);
      else if (st == GH_FAST_UNSURE) need |= 1u << isub;
    } else {
      long long pe;
      const int se = gh_point_to_shell_pixel(t, x0 + d.sub_off[isub], y0 + d.sub_off[GH_CUDA_N_SUBPART + isub],
                                             z0 + d.sub_off[2 * GH_CUDA_N_SUBPART + isub], (double)dzf, &pe);
      if (st == GH_FAST_OUT) { acp->out++; if (pe >= 0) acp->wrong++; }
      else if (st == GH_FAST_IN) { acp->in++; if (inu != se || (long long)pix != pe) acp->wrong++; }
      else acp->unsure++;
    }
  }
  return need;
}

// ---- accumulate_kernel ------------------------------------------------------------------------------------
// One thread per 2 x 2 x 2 block of cells; a warp covers 8 x 8 x 4 cells, so that its lanes see nearly the same
// shells and the same HEALPix regime, a CTA of four warps 16 x 16 x 4 cells.
//   1. Block cull: every sub-particle lies within sqrt(3) dx of the block centre and z_of_r is monotone, so their
//      redshifts lie in [z(r_C - hg) + min dz, z(r_C + hg) + max dz]; blocks whose bracket misses the shells'
//      redshift window are skipped (about 44 % of the box for the shipped frequency table).
//   2. Block expansion (gh_group_math.cuh): the two pixel coordinates and r^2 to second order about the block
//      centre, once per block; per cell the expansion point is shifted to the cell centre (exact for quadratics)
//      and the shell edges in reach become thresholds on r^2; per sub-particle 18 FMAs with constant offsets,
//      the 1.5 * 2^23 rounding trick instead of floor(), a handful of integer operations and one RED.
//   3. Cells the expansion does not cover (a few per cent, but one in most warps) are queued in shared memory and
//      taken through the per-cell fp32 path above by all threads of the CTA afterwards -- run in place they would
//      hold the whole warp for a thousand instructions each.
//   4. Sub-particles that either fast path calls unsure (a rounding decision closer to its boundary than the
//      error bound) are queued likewise and re-done in IEEE double with the exact restatement of the reference's
//      arithmetic (gh_point_to_shell_pixel).
// Accepted fast answers equal the exact path's by construction; AUDIT (gh_cuda_accumulate_audit) measures that on
// the device: nothing is deposited, every sub-particle is evaluated by both paths and the outcomes are counted,
// also with the error bounds scaled down.
// Planes: the launch covers `nplanes` consecutive planes in pairs; plane k sits at local index iz_base + k of the
// buffers passed in and is global plane zg_base + k (this rank's own slab, or planes pulled from a neighbour for
// load balance -- see enqueue_maps in gh_api.cu).  An odd last plane is a block with four cells.
#ifndef GH_ACC_MIN_BLOCKS
#define GH_ACC_MIN_BLOCKS 7
#endif
#define GH_ACC_QCAP (128 * 8 * GH_CUDA_N_SUBPART)
#ifndef GH_ACC_SUB_UNROLL
#define GH_ACC_SUB_UNROLL 2
#endif
constexpr int kAccSubUnroll = GH_ACC_SUB_UNROLL;  // #pragma unroll takes a constant expression, not a macro

// thread -> block of cells: a warp covers 4 x 4 x 2 blocks = 8 x 8 x 4 cells (compact, so that few warps straddle the
// shells' radial window, the |cos theta| = 2/3 cones or the quadrant boundaries), the four warps of a CTA 16 x 16 x 4
__device__ __forceinline__ void acc_block_of(int tid, int &cx, int &cy, int &pz0)
{
  const int lane = tid & 31, wid = tid >> 5;
  cx = 2 * (blockIdx.x * 8 + (wid & 1) * 4 + (lane & 3));
  cy = 2 * (blockIdx.y * 8 + (wid >> 1) * 4 + ((lane >> 2) & 3));
  pz0 = 2 * (blockIdx.z * 2 + (lane >> 4));
}

template <bool AUDIT>
__global__ void __launch_bounds__(128, GH_ACC_MIN_BLOCKS) accumulate_kernel(const __grid_constant__ GhDev d, const float *__restrict__ mass,
                                                         const float *__restrict__ dzrsd, float *__restrict__ maps,
                                                         float eps_scale, unsigned long long *__restrict__ counts,
                                                         int iz_base, int zg_base, int nplanes)
{
  __shared__ unsigned short xqueue[GH_ACC_QCAP];  // unsure sub-particles: tid << 7 | cell << 4 | sub-particle
  __shared__ unsigned short cqueue[128 * 8];      // cells for the per-cell path: tid << 3 | cell
  __shared__ float s_w[8][128], s_dz[8][128];
  __shared__ int s_nx, s_nc;
  const int ngx = 2 * d.nh;
  const int tid = threadIdx.x;
  if (tid == 0) { s_nx = 0; s_nc = 0; }
  int cx, cy, pz0;  // first cell of this thread's block, first plane relative to the launch
  acc_block_of(tid, cx, cy, pz0);
  const int nzv = min(2, nplanes - pz0);
  const int zg = zg_base + pz0;
  const bool active = (cx < d.n) && (cy < d.n) && (nzv > 0);
  const unsigned valid = active ? (nzv == 2 ? 0xFFu : 0x0Fu) : 0u;
  AuditCounts ac = {0, 0, 0, 0};
  float dz_min = 3.0e38f, dz_max = -3.0e38f;
  float2 ms[4];  // masses of the block's cells, two per (plane, row)
#pragma unroll
  for (int pz = 0; pz < 2; ++pz) {
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      float2 m = make_float2(0.f, 0.f), z = make_float2(0.f, 0.f);
      if (active && pz < nzv) {
        const size_t idx = ((size_t)(iz_base + pz0 + pz) * d.n + (cy + py)) * ngx + cx;
        m = __ldcs(reinterpret_cast<const float2 *>(mass + idx));
        z = __ldcs(reinterpret_cast<const float2 *>(dzrsd + idx));
        dz_min = fminf(dz_min, fminf(z.x, z.y));
        dz_max = fmaxf(dz_max, fmaxf(z.x, z.y));
      }
      const int c = pz * 4 + py * 2;
      ms[pz * 2 + py] = m;
      s_dz[c][tid] = z.x;
      s_dz[c + 1][tid] = z.y;
    }
  }
  __syncthreads();  // the queue counters
  if (active) {
    const FastCtx f = fast_ctx_of(d, eps_scale);
    // block centre (exact, double)
    const double X = d.dx * (cx + 1.0) - d.pos_obs[0], Y = d.dx * (cy + 1.0) - d.pos_obs[1], Z = d.dx * (zg + 1.0) - d.pos_obs[2];
    const float xh = (float)X, yh = (float)Y, zh = (float)Z;
    const float rc = sqrtf(fmaf(zh, zh, fmaf(xh, xh, yh * yh)));
    const float hg = (float)d.dx * 1.7320508f + 1e-3f + 1e-6f * rc;
    const float zs_hi = z_of_r_f(f, rc + hg) + dz_max + 1e-5f, zs_lo = z_of_r_f(f, rc - hg) + dz_min - 1e-5f;
    const bool culled = (zs_hi < (float)d.z_lo_cull || zs_lo > (float)d.z_hi_cull);
    if (AUDIT && culled) {
      const GhIndexTables t = tables_of(d);
      for (int c = 0; c < 8; ++c) {
        if (!((valid >> c) & 1u)) continue;
        const double x0 = d.dx * (cx + (c & 1) + 0.5) - d.pos_obs[0], y0 = d.dx * (cy + ((c >> 1) & 1) + 0.5) - d.pos_obs[1];
        const double z0 = d.dx * (zg + (c >> 2) + 0.5) - d.pos_obs[2];
        for (int isub = 0; isub < GH_CUDA_N_SUBPART; ++isub) {
          long long pe;
          gh_point_to_shell_pixel(t, x0 + d.sub_off[isub], y0 + d.sub_off[GH_CUDA_N_SUBPART + isub],
                                  z0 + d.sub_off[2 * GH_CUDA_N_SUBPART + isub], (double)s_dz[c][tid], &pe);
          ac.out++;
          if (pe >= 0) ac.wrong++;
        }
      }
    }
    if (!culled) {
      // a tenth of each cell's mass (src/pixelize.c:203: float / int is a float division in C); only blocks that
      // survive the cull pay for the eight divisions
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s_w[2 * k][tid] = __fdiv_rn(ms[k].x, (float)GH_CUDA_N_SUBPART);
        s_w[2 * k + 1][tid] = __fdiv_rn(ms[k].y, (float)GH_CUDA_N_SUBPART);
      }
      const float dz_mid = 0.5f * (dz_min + dz_max);
      const GroupShells gs = group_shells(f, zs_lo, zs_hi, dz_mid);
      GhGroupExp g;
      g.kind = GH_GRP_NONE;
      g.rc = 0.0;
      g.e = 0.f;
      if (gs.ok) gh_group_expand(X, Y, Z, hg, f.fns, eps_scale, g);
      if (g.kind == GH_GRP_NONE) {
        // the whole block goes to the per-cell path
        int at = atomicAdd(&s_nc, __popc(valid));
        for (int c = 0; c < 8; ++c)
          if ((valid >> c) & 1u) cqueue[at++] = (unsigned short)((tid << 3) | c);
      } else {
        const float rc_hi = (float)g.rc, rc_lo = (float)(g.rc - (double)rc_hi);
        const float hm = 0.5f - g.e;
        const float hdx = (float)(0.5 * d.dx);
        const int ns = f.ns, n_nu = f.n_nu, npix32 = (int)d.npix, j0 = gs.j_in;
        const bool polar = g.kind != GH_GRP_EQ;
        // equatorial: pix = k_b + (jp - jm) k_a + ip; polar: pix = ir (k_a ir + k_b) + ip_f + k_c (gh_group_math.cuh)
        int k_a, k_b, k_c;
        if (!polar) {
          k_a = 4 * ns;
          k_b = 2 * ns * (ns - 1) + ns * 4 * ns;
          k_c = (int)(2u * (unsigned)g.kbase - (unsigned)ns + 1u - 2u * (unsigned)GH_GRP_MAGIC_BITS);
        } else {
          k_a = (g.kind == GH_GRP_SOUTH) ? -2 : 2;
          k_b = g.kbase - 2;
          k_c = (int)((g.kind == GH_GRP_SOUTH ? (unsigned)npix32 : 0u) - (unsigned)GH_GRP_MAGIC_BITS);
        }
        // global address of shell j0's map (never dereferenced out of range).  Opaque to the optimiser from here on:
        // otherwise it folds the 64-bit base arithmetic back into every sub-particle's address
        unsigned long long shell_base =
            (unsigned long long)__cvta_generic_to_global(maps) + 4ull * (unsigned long long)((long long)d.npix * j0);
        int npix_neg = -npix32;
        asm volatile("" : "+l"(shell_base), "+r"(npix_neg));
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          if (!((valid >> c) & 1u)) continue;
          const float w = s_w[c][tid], dzc = s_dz[c][tid];
          GhCellExp ce;
          gh_group_recentre(g, (c & 1) ? hdx : -hdx, (c & 2) ? hdx : -hdx, (c & 4) ? hdx : -hdx, ce);
          // shell thresholds of this cell: the block's edge radii shifted by its own Delta z_RSD
          const float ddz = dzc - dz_mid;
          const float eps = fmaf(d.rz_slope_var, fabsf(ddz), f.eps_r);
          float lo_a, hi_a, lo_b, hi_b, lo_c, hi_c;
          edge_thresholds(fmaf(-ddz, gs.s0, gs.r0), eps, rc_hi, rc_lo, lo_a, hi_a);
          edge_thresholds(fmaf(-ddz, gs.s1, gs.r1), eps, rc_hi, rc_lo, lo_b, hi_b);
          edge_thresholds(fmaf(-ddz, gs.s2, gs.r2), eps, rc_hi, rc_lo, lo_c, hi_c);
          unsigned need = 0u;
          // one sub-particle; `pol` is a literal at both call sites, so each loop carries only its own regime
          auto sub = [&](int isub, bool pol) {
            const float4 o = d.sub_c[isub];
            const float S = gh_cell_S(ce, o.x, o.y, o.z, o.w);
            // shell: j0 inside the innermost edge, one less beyond each edge crossed; unsure within eps of an edge
            const bool p1 = S > hi_a, p2 = S > hi_b, p3 = S > hi_c;
            const bool sure = ((S < lo_a) || p1) && ((S < lo_b) || p2) && ((S < lo_c) || p3);
            int dj = p1 ? 1 : 0;
            if (p2) dj++;
            if (p3) dj++;
            const bool inside = (unsigned)(j0 - dj) < (unsigned)n_nu;
            const float U = gh_cell_U(g, ce, o.x, o.y), V = gh_cell_V(g, ce, o.x, o.y, o.z);
            int pix;
            const bool ok = pol ? gh_sub_polar(U, V, hm, k_a, k_b, k_c, pix) : gh_sub_eq(U, V, hm, k_a, k_b, k_c, pix);
            if (!AUDIT) {
              if (sure && inside && ok) {
                const int rel = dj * npix_neg + pix;
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(shell_base + 4ll * (long long)rel), "f"(w));
              } else if (!sure || inside) {
                need |= 1u << isub;  // sure && !inside: surely outside every shell
              }
            } else {
              const GhIndexTables t = tables_of(d);
              const double x0 = d.dx * (cx + (c & 1) + 0.5) - d.pos_obs[0], y0 = d.dx * (cy + ((c >> 1) & 1) + 0.5) - d.pos_obs[1];
              const double z0 = d.dx * (zg + (c >> 2) + 0.5) - d.pos_obs[2];
              long long pe;
              const int se = gh_point_to_shell_pixel(t, x0 + d.sub_off[isub], y0 + d.sub_off[GH_CUDA_N_SUBPART + isub],
                                                     z0 + d.sub_off[2 * GH_CUDA_N_SUBPART + isub], (double)dzc, &pe);
              if (sure && !inside) { ac.out++; if (pe >= 0) ac.wrong++; }
              else if (sure && ok) { ac.in++; if (j0 - dj != se || (long long)pix != pe) ac.wrong++; }
              else ac.unsure++;
            }
          };
          if (polar) {
#pragma unroll kAccSubUnroll
            for (int isub = 0; isub < GH_CUDA_N_SUBPART; ++isub) sub(isub, true);
          } else {
#pragma unroll kAccSubUnroll
            for (int isub = 0; isub < GH_CUDA_N_SUBPART; ++isub) sub(isub, false);
          }
          if (!AUDIT && need) {
            int at = atomicAdd(&s_nx, __popc(need));
            while (need) {
              const int isub = __ffs(need) - 1;
              need &= need - 1;
              xqueue[at++] = (unsigned short)((tid << 7) | (c << 4) | isub);
            }
          }
        }
      }
    }
  }
  // ---- per-cell path for the queued cells, all threads ----
  __syncthreads();
  const int ncells = s_nc;
  for (int j = tid; j < ncells; j += 128) {
    const unsigned e = cqueue[j];
    const int src = e >> 3, c = e & 7;
    int bx, by, bz;
    acc_block_of(src, bx, by, bz);
    const int sx = bx + (c & 1), sy = by + ((c >> 1) & 1), sz = zg_base + bz + (c >> 2);
    const double x0 = d.dx * (sx + 0.5) - d.pos_obs[0], y0 = d.dx * (sy + 0.5) - d.pos_obs[1];
    const double z0 = d.dx * (sz + 0.5) - d.pos_obs[2];
    unsigned need = generic_cell<AUDIT>(d, eps_scale, x0, y0, z0, s_w[c][src], s_dz[c][src], maps, AUDIT ? &ac : nullptr);
    if (!AUDIT && need) {
      int at = atomicAdd(&s_nx, __popc(need));
      while (need) {
        const int isub = __ffs(need) - 1;
        need &= need - 1;
        xqueue[at++] = (unsigned short)((src << 7) | (c << 4) | isub);
      }
    }
  }
  if (AUDIT) {
    atomicAdd(counts + 0, ac.out);
    atomicAdd(counts + 1, ac.in);
    atomicAdd(counts + 2, ac.unsure);
    atomicAdd(counts + 3, ac.wrong);
    return;
  }
  // ---- exact path for the unsure sub-particles of this CTA ----
  __syncthreads();
  const int total = s_nx;
  if (total == 0) return;
  const GhIndexTables t = tables_of(d);
  for (int j = tid; j < total; j += 128) {
    const unsigned e = xqueue[j];
    const int src = e >> 7, c = (e >> 4) & 7, isub = e & 15;
    int bx, by, bz;
    acc_block_of(src, bx, by, bz);
    const int sx = bx + (c & 1), sy = by + ((c >> 1) & 1), sz = zg_base + bz + (c >> 2);
    const double px = d.dx * (sx + 0.5) - d.pos_obs[0] + d.sub_off[isub];
    const double py = d.dx * (sy + 0.5) - d.pos_obs[1] + d.sub_off[GH_CUDA_N_SUBPART + isub];
    const double pz = d.dx * (sz + 0.5) - d.pos_obs[2] + d.sub_off[2 * GH_CUDA_N_SUBPART + isub];
    long long ipix;
    const int inu = gh_point_to_shell_pixel(t, px, py, pz, (double)s_dz[c][src], &ipix);
    if (ipix >= 0) atomicAdd(maps + (size_t)ipix + (size_t)d.npix * inu, s_w[c][src]);
  }
}

// Audit of the fast path against the exact path on arbitrary points: counts[0..3] = fast says out / in /
// unsure / (fast was sure but disagrees with the exact path -- must stay 0).
__global__ void __launch_bounds__(128) fastpath_audit_kernel(GhDev d, const double *__restrict__ pos,
                                                             const double *__restrict__ dz, long long n, float eps_scale,
                                                             unsigned long long *__restrict__ counts)
{
  const GhIndexTables t = tables_of(d);
  const FastCtx f = fast_ctx_of(d, eps_scale);
  unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double xd = pos[3 * i], yd = pos[3 * i + 1], zd = pos[3 * i + 2];
    const float dzf = dz ? (float)dz[i] : 0.f;
    long long pe;
    const int se = gh_point_to_shell_pixel(t, xd, yd, zd, (double)dzf, &pe);
    const float x = (float)xd, y = (float)yd, z = (float)zd;
    const float q = fmaf(x, x, y * y), r2 = fmaf(z, z, q);
    int st = GH_FAST_UNSURE, inu = -7, pix = -1;
    if (r2 > 1e-12f) {
      const float inv_r = rsqrt_ftz(r2);
      const float nu = 1420.40575177f * rcp_ftz(1.0f + (z_of_r_f(f, r2 * inv_r) + dzf));
      st = fast_shell(f, nu, inu);
      if (st == GH_FAST_IN) {
        float tt = atan2f(y, x) * 0.63661977236758134308f;
        tt += (tt < 0.f) ? 4.0f : 0.f;
        if (!fast_pixel(f, z * inv_r, tt, q, inv_r, pix)) st = GH_FAST_UNSURE;
      }
    }
    if (st == GH_FAST_OUT) { c0++; if (pe >= 0) c3++; }
    else if (st == GH_FAST_IN) { c1++; if (inu != se || (long long)pix != pe) c3++; }
    else c2++;
  }
  atomicAdd(counts + 0, c0);
  atomicAdd(counts + 1, c1);
  atomicAdd(counts + 2, c2);
  atomicAdd(counts + 3, c3);
}

// src/pixelize.c:236-261: one prefactor per shell, float * double -> float
__global__ void __launch_bounds__(256) scale_maps_kernel(float4 *__restrict__ maps, const double *__restrict__ prefac,
                                                         long long npix4, int shell0)
{
  const int sh = blockIdx.y;
  const double pf = prefac[shell0 + sh];
  float4 *m = maps + (size_t)sh * npix4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = m[i];
    v.x = (float)((double)v.x * pf);
    v.y = (float)((double)v.y * pf);
    v.z = (float)((double)v.z * pf);
    v.w = (float)((double)v.w * pf);
    m[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Sparse map reduction over peer memory (several ranks with peer mappings; GH_NO_SPARSE_REDUCE=1 falls back to
// ncclReduceScatter).
// A rank accumulates a contiguous range of z planes, so in every shell the pixels it touched lie in a band of
// latitudes -- a contiguous interval of RING indices -- and most of its [n_nu][npix] stack is zero.  NCCL's
// reduce-scatter moves (P-1)/P of the whole stack per rank regardless.  Instead: (1) every rank measures the
// touched interval [lo, hi) of each shell of its own stack, (2) the intervals are all-gathered (which is also
// the barrier that says everybody has finished accumulating), (3) the owner of a shell sums, pixel by pixel and
// in rank order, exactly the peers' intervals that cover the pixel, reading them straight from the peers'
// stacks over NVLink, applies the shell's temperature prefactor and writes its result.
__global__ void __launch_bounds__(256) shell_extent_kernel(const float4 *__restrict__ maps, long long npix4,
                                                           int *__restrict__ ext_lo, int *__restrict__ ext_hi)
{
  const int sh = blockIdx.y;
  const float4 *m = maps + (size_t)sh * npix4;
  // contiguous chunk per CTA: the first and last non-zero float4 of the chunk
  const long long per = (npix4 + gridDim.x - 1) / gridDim.x;
  const long long a = (long long)blockIdx.x * per, b = (a + per < npix4) ? a + per : npix4;
  long long lo = npix4, hi = -1;
  for (long long i = a + threadIdx.x; i < b; i += blockDim.x) {
    const float4 v = __ldg(m + i);
    if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
      lo = (i < lo) ? i : lo;
      hi = (i > hi) ? i : hi;
    }
  }
  __shared__ long long s_lo[256], s_hi[256];
  s_lo[threadIdx.x] = lo;
  s_hi[threadIdx.x] = hi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_lo[threadIdx.x] = min(s_lo[threadIdx.x], s_lo[threadIdx.x + o]);
      s_hi[threadIdx.x] = max(s_hi[threadIdx.x], s_hi[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && s_hi[0] >= 0) {  // units: float4 index; hi exclusive
    atomicMin(ext_lo + sh, (int)s_lo[0]);
    atomicMax(ext_hi + sh, (int)s_hi[0] + 1);
  }
}

struct MapPeers {
  const float4 *m[GH_MAX_RANKS];
};

// all_ext: [rank][2][n_nu_pad] (lo then hi, float4 units).  One CTA = one chunk of one owned shell.
__global__ void __launch_bounds__(256) sparse_reduce_kernel(MapPeers peers, const int *__restrict__ all_ext, int nranks,
                                                            int n_nu_pad, long long npix4, int shell0,
                                                            const double *__restrict__ prefac, float4 *__restrict__ out)
{
  const int sh = shell0 + blockIdx.y;
  const double pf = prefac[sh];
  __shared__ int s_lo[GH_MAX_RANKS], s_hi[GH_MAX_RANKS];
  if ((int)threadIdx.x < nranks) {
    s_lo[threadIdx.x] = all_ext[((size_t)threadIdx.x * 2 + 0) * n_nu_pad + sh];
    s_hi[threadIdx.x] = all_ext[((size_t)threadIdx.x * 2 + 1) * n_nu_pad + sh];
  }
  __syncthreads();
  float4 *o = out + (size_t)blockIdx.y * npix4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < nranks; ++q) {
      if (i >= s_lo[q] && i < s_hi[q]) {
        const float4 v = peers.m[q][(size_t)sh * npix4 + i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    acc.x = (float)((double)acc.x * pf);  // src/pixelize.c:236-261, as scale_maps_kernel
    acc.y = (float)((double)acc.y * pf);
    acc.z = (float)((double)acc.z * pf);
    acc.w = (float)((double)acc.w * pf);
    o[i] = acc;
  }
}

__global__ void __launch_bounds__(128) points_kernel(GhDev d, const double *__restrict__ pos, const double *__restrict__ dz,
                                                     long long n, int *__restrict__ shell, long long *__restrict__ pix)
{
  const GhIndexTables t = tables_of(d);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long ipix;
    const int inu = gh_point_to_shell_pixel(t, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dz ? dz[i] : 0.0, &ipix);
    shell[i] = inu;
    pix[i] = ipix;
  }
}

}  // namespace

int gh_launch_accumulate(gh_cuda_ctx *c, const float *mass, const float *dzrsd, int iz_base, int zg_base, int nplanes)
{
  const GhDev &d = c->d;
  if (nplanes <= 0) return 0;
  dim3 grid((d.n + 15) / 16, (d.n + 15) / 16, (nplanes + 3) / 4), block(128);
  accumulate_kernel<false><<<grid, block, 0, c->stream>>>(d, mass, dzrsd, c->maps, 1.0f, nullptr, iz_base, zg_base, nplanes);
  GH_LAUNCH_CHECK(c);
  return 0;
}

int gh_launch_accumulate_audit(gh_cuda_ctx *c, float eps_scale, unsigned long long *d_counts)
{
  const GhDev &d = c->d;
  dim3 grid((d.n + 15) / 16, (d.n + 15) / 16, (d.nz_here + 3) / 4), block(128);
  const float *m = reinterpret_cast<const float *>(c->gridA), *z = reinterpret_cast<const float *>(c->gridC);
  accumulate_kernel<true><<<grid, block, 0, c->stream>>>(d, m, z, c->maps, eps_scale, d_counts, 0, d.iz0, d.nz_here);
  GH_LAUNCH_CHECK(c);
  return 0;
}

int gh_launch_scale_maps(gh_cuda_ctx *c, float *maps, int shell0, int nshells)
{
  if (nshells <= 0) return 0;
  const long long npix4 = c->d.npix / 4;
  int bx = (int)((npix4 + 255) / 256);
  if (bx > 1024) bx = 1024;
  dim3 grid(bx, nshells);
  scale_maps_kernel<<<grid, 256, 0, c->stream>>>(reinterpret_cast<float4 *>(maps), c->d_prefac, npix4, shell0);
  GH_LAUNCH_CHECK(c);
  return 0;
}

int gh_launch_shell_extents(gh_cuda_ctx *c, int *ext_lo, int *ext_hi)
{
  const long long npix4 = c->d.npix / 4;
  int bx = (int)((npix4 + 256 * 64 - 1) / (256 * 64));  // >= 64 float4 per thread
  if (bx < 1) bx = 1;
  if (bx > 256) bx = 256;
  dim3 grid(bx, c->d.n_nu_pad);
  shell_extent_kernel<<<grid, 256, 0, c->stream>>>(reinterpret_cast<const float4 *>(c->maps), npix4, ext_lo, ext_hi);
  GH_LAUNCH_CHECK(c);
  return 0;
}

int gh_launch_sparse_reduce(gh_cuda_ctx *c, const int *all_ext, float *out, int shell0, int nshells)
{
  if (nshells <= 0) return 0;
  MapPeers mp;
  for (int q = 0; q < GH_MAX_RANKS; ++q) mp.m[q] = (q < c->d.nranks) ? reinterpret_cast<const float4 *>(c->map_peers[q]) : nullptr;
  const long long npix4 = c->d.npix / 4;
  int bx = (int)((npix4 + 255) / 256);
  if (bx > 1024) bx = 1024;
  dim3 grid(bx, nshells);
  sparse_reduce_kernel<<<grid, 256, 0, c->stream>>>(mp, all_ext, c->d.nranks, c->d.n_nu_pad, npix4, shell0, c->d_prefac,
                                                   reinterpret_cast<float4 *>(out));
  GH_LAUNCH_CHECK(c);
  return 0;
}

int gh_launch_fastpath_audit(gh_cuda_ctx *c, const double *d_pos, const double *d_dz, long long n, float eps_scale,
                              unsigned long long *d_counts)
{
  if (n <= 0) return 0;
  long long blocks = (n + 127) / 128;
  if (blocks > c->n_sm * 16) blocks = c->n_sm * 16;
  fastpath_audit_kernel<<<(unsigned)blocks, 128, 0, c->stream>>>(c->d, d_pos, d_dz, n, eps_scale, d_counts);
  GH_LAUNCH_CHECK(c);
  return 0;
}

int gh_launch_points(gh_cuda_ctx *c, const double *d_pos, const double *d_dz, long long n, int *d_shell, long long *d_pix)
{
  if (n <= 0) return 0;
  long long blocks = (n + 127) / 128;
  if (blocks > c->n_sm * 16) blocks = c->n_sm * 16;
  points_kernel<<<(unsigned)blocks, 128, 0, c->stream>>>(c->d, d_pos, d_dz, n, d_shell, d_pix);
  GH_LAUNCH_CHECK(c);
  return 0;
}
