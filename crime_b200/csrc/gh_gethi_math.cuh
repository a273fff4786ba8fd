// Per-cell arithmetic of get_HI (reference src/grid_tools.c:103-153, src/cosmo.c:52-86, src/user_defined.c:27-35)
// and the cell-centre coordinates, written with explicit rounding intrinsics only, so that the stand-alone
// get_HI kernel (gh_fields.cu) and the fused get_HI + mk_T_maps kernel (gh_pixelize.cu, built with
// -fmad=false) produce bit-identical HI masses and Delta z_RSD.
#pragma once

// Cell-centre coordinate along one axis, dx*(i+0.5) - pos_obs, in float without cancellation error:
// the integer part of (0.5 - pos_obs/dx) is subtracted from the index exactly, the fraction rides on an fma.
struct AxisF {
  int ioff;
  float dx, frac;
  __device__ __forceinline__ float at(int i) const { return fmaf(dx, (float)(i - ioff), frac); }
};
__host__ __device__ __forceinline__ AxisF make_axis(double dx, double pos_obs, int i_origin)
{
  // x = dx*(i_local + i_origin + 0.5) - pos_obs = dx*((i_local - ioff) + f),  f in [0,1)
  const double t = (double)i_origin + 0.5 - pos_obs / dx;
  const double fl = floor(t);
  AxisF a;
  a.ioff = -(int)fl;
  a.dx = (float)dx;
  a.frac = (float)(dx * (t - fl));
  return a;
}

struct Axes3 { AxisF x, y, z; };  // built once on the host, passed to the kernels by value

struct GetHIConsts {
  const float *ztab, *gdtab, *gvtab, *fractab, *biastab;
  int last;
  float idr, rmax, half_s2, dx3;
};

__device__ __forceinline__ float gh_lerp_tab(const float *__restrict__ tab, int ir, float t)
{
  const float a = __ldg(tab + ir), b = __ldg(tab + ir + 1);
  return fmaf(__fsub_rn(b, a), t, a);
}

// r2 = squared distance of the cell centre from the observer; delta, rvel = the cell's Gaussian overdensity
// and radial velocity.  Outputs the HI mass and the redshift-space shift.
__device__ __forceinline__ void gh_gethi_cell(const GetHIConsts &k, float r2, float delta, float rvel, float &mass, float &dz)
{
  const float r = __fmul_rn(r2, rsqrtf(fmaxf(r2, 1e-30f)));
  // z_of_r / dgrowth_of_r / vgrowth_of_r: same bin, same weight (src/cosmo.c:52-86); r=0 and the clamp beyond
  // the table fall out of the arithmetic (tables start at (0,1,1))
  const float s = __fmul_rn(fminf(r, k.rmax), k.idr);
  const int ir = min((int)s, k.last - 1);
  const float t = __fsub_rn(s, (float)ir);
  // bias_HI(z(r)) and fraction_HI(z(r)) come from the host's tabulation of the user hooks on the same radial grid
  // (src/user_defined.c:27-35); linear interpolation of them in r differs from evaluating them at the interpolated
  // redshift by O(f'' dz^2 / 8) ~ 1e-8 relative for any smooth hook
  const float gd = gh_lerp_tab(k.gdtab, ir, t);
  const float gv = gh_lerp_tab(k.gvtab, ir, t);
  const float bias = gh_lerp_tab(k.biastab, ir, t);
  const float frac = gh_lerp_tab(k.fractab, ir, t);
  const float gfd = __fmul_rn(gd, bias);                                                                 // D(r) b_HI(z)
  const float dens_ln = exp2f(__fmul_rn(1.4426950408889634f, __fmul_rn(gfd, fmaf(-k.half_s2, gfd, delta))));  // lognormal
  mass = __fmul_rn(__fmul_rn(k.dx3, frac), dens_ln);                                                     // dx^3 x_HI(z) rho_LN
  dz = __fmul_rn(rvel, gv);                                                                              // Delta z_RSD
}

__device__ __forceinline__ GetHIConsts make_gethi_consts(const GhDev &d, float sigma2_gauss)
{
  GetHIConsts k;
  k.ztab = d.z_r2z_f; k.gdtab = d.gd_f; k.gvtab = d.gv_f; k.fractab = d.frac_f; k.biastab = d.bias_f;
  k.last = d.nz_tab - 1;
  k.idr = (float)d.glob_idr; k.rmax = (float)d.r_tab_max;
  k.half_s2 = __fmul_rn(0.5f, sigma2_gauss);
  k.dx3 = (float)(d.dx * d.dx * d.dx);
  return k;
}
