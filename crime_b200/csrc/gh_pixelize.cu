// mk_T_maps on sm_100a (reference src/pixelize.c:150-262): every cell is split into 10 sub-particles at
// fixed offsets, each is sent to (frequency shell, HEALPix RING pixel) and deposits a tenth of the cell's
// HI mass.  Shell and pixel indices must equal the reference's bit for bit, so this translation unit is
// built with -fmad=false and does the geometry in IEEE double (gh_index_math.cuh).
//
// The stage is bound by double-precision instruction and L2-atomic throughput, not by HBM (8 B/cell of
// grid traffic against ~10 sub-particles x ~150 fp64 instructions): cells whose whole extent lies
// outside the shells' redshift window are culled by a conservative radial bound first (about half of
// the box for the shipped frequency table).
#include "gh_internal.cuh"
#include "gh_index_math.cuh"

namespace {

__device__ __forceinline__ GhIndexTables tables_of(const GhDev &d)
{
  GhIndexTables t;
  t.z_r2z = d.z_r2z; t.r_r2z = d.r_r2z; t.nz_tab = d.nz_tab; t.glob_idr = d.glob_idr;
  t.nu0 = d.nu0; t.nuf = d.nuf; t.n_nu = d.n_nu; t.irregular = d.irregular;
  t.nu_min = d.nu_min; t.inv_dnu = d.inv_dnu; t.nside = d.nside;
  return t;
}

__global__ void __launch_bounds__(128) accumulate_kernel(GhDev d, const float *__restrict__ mass,
                                                         const float *__restrict__ dzrsd, float *__restrict__ maps)
{
  const int ngx = 2 * d.nh;
  const int iy = blockIdx.y, iz = blockIdx.z;
  const int ix = blockIdx.x * blockDim.x + threadIdx.x;
  if (ix >= d.n) return;
  const GhIndexTables t = tables_of(d);
  const double x0 = d.dx * (ix + 0.5) - d.pos_obs[0];
  const double y0 = d.dx * (iy + 0.5) - d.pos_obs[1];
  const double z0 = d.dx * (iz + d.iz0 + 0.5) - d.pos_obs[2];
  const size_t idx = ((size_t)iz * d.n + iy) * ngx + ix;
  const double dz = (double)dzrsd[idx];
  // conservative cull: all sub-particles lie within half a cell diagonal of the centre and z_of_r is
  // non-decreasing, so their redshifts lie in [z(rc-h), z(rc+h)] + dz
  {
    const double rc = sqrt(x0 * x0 + y0 * y0 + z0 * z0);
    const double h = d.dx * 0.8660254037844387 + 1e-9 * (rc + d.dx);
    const double zs_hi = gh_z_of_r(t, rc + h) + dz, zs_lo = gh_z_of_r(t, rc - h) + dz;
    if (zs_hi < d.z_lo_cull || zs_lo > d.z_hi_cull) return;
  }
  const double mass_sub = (double)mass[idx] / GH_CUDA_N_SUBPART;  // src/pixelize.c:203
  const float w = (float)mass_sub;
#pragma unroll 1
  for (int isub = 0; isub < GH_CUDA_N_SUBPART; ++isub) {
    const double x = x0 + d.sub_off[isub];
    const double y = y0 + d.sub_off[GH_CUDA_N_SUBPART + isub];
    const double z = z0 + d.sub_off[2 * GH_CUDA_N_SUBPART + isub];
    long long ipix;
    const int inu = gh_point_to_shell_pixel(t, x, y, z, dz, &ipix);
    if (ipix >= 0) atomicAdd(maps + (size_t)ipix + (size_t)d.npix * inu, w);
  }
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

int gh_launch_accumulate(gh_cuda_ctx *c)
{
  const GhDev &d = c->d;
  dim3 grid((d.n + 127) / 128, d.n, d.nz_here);
  accumulate_kernel<<<grid, 128, 0, c->stream>>>(d, reinterpret_cast<const float *>(c->gridA),
                                                 reinterpret_cast<const float *>(c->gridC), c->maps);
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

int gh_launch_points(gh_cuda_ctx *c, const double *d_pos, const double *d_dz, long long n, int *d_shell, long long *d_pix)
{
  if (n <= 0) return 0;
  long long blocks = (n + 127) / 128;
  if (blocks > c->n_sm * 16) blocks = c->n_sm * 16;
  points_kernel<<<(unsigned)blocks, 128, 0, c->stream>>>(c->d, d_pos, d_dz, n, d_shell, d_pix);
  GH_LAUNCH_CHECK(c);
  return 0;
}
