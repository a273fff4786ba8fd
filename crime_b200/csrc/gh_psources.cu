// Point sources on the same grids: get_point_sources (reference src/grid_tools.c:24-101) Poisson-samples
// a number of sources per cell from the Gaussian density field, mk_psources_maps (src/pixelize.c:58-148) scatters
// every source inside its cell, draws its luminosity (draw_luminosity, src/psources.c:133-154) and adds its
// brightness temperature (temp_of_l, src/psources.c:159-167) to the same pixel of every frequency shell.
//
// The reference draws from one MT19937 per OpenMP thread, so its catalogue depends on the thread and rank count; here
// every cell draws from Philox4x32-10 keyed on (seed, GLOBAL cell index), identical for any number of GPUs.  Parity
// with the reference is therefore statistical (source counts against the Poisson means, shell temperatures), plus
// exact for the deterministic pieces (the means themselves, temp_of_l), see tests/test_gpu_psources.py.
//   * Poisson: multiplication method below lambda = 12, Hoermann's PTRS transformed rejection above (any lambda).
//   * Luminosity: the reference rejection-samples P(log10 L | z) under a flat envelope; here the cumulative
//     distribution is tabulated by the host per redshift bin ([nz][nl + 1]) and inverted -- the same distribution up
//     to the bin width of the table -- so the user-definable luminosity function stays a host function.
//   * spec_ed (user-definable SED) is tabulated by the host on a uniform grid in log10(nu).
// Positions, redshifts and pixels of the sources are IEEE double (gh_index_math.cuh, the exact path of mk_T_maps).
#include "gh_internal.cuh"
#include "gh_index_math.cuh"
#include "gh_philox.cuh"

namespace {

struct PsDev {
  int nz;
  double z_max, inv_dz;
  const double *nz_arr, *bias_arr;
  int nl;
  double logl_min, dlogl;
  const double *lcdf;
  int nsed;
  double lognu_min, inv_dlognu;
  const double *sed;
  double hhub;
};

// per-cell uniform stream: counter = (cell lo, cell hi, block, stream), key = (seed, 'PSrc')
struct CellRng {
  uint32_t c0, c1, blk, stream, seed;
  uint32_t w[4];
  int have;
  __device__ CellRng(unsigned long long cell, uint32_t stream_, uint32_t seed_)
      : c0((uint32_t)cell), c1((uint32_t)(cell >> 32)), blk(0), stream(stream_), seed(seed_), have(0) {}
  __device__ double uniform()  // [0, 1) with 32-bit resolution, as gsl_rng_uniform of MT19937
  {
    if (!have) {
      philox4x32_10(c0, c1, seed, 0x50537263u, w[0], w[1], w[2], w[3], blk++, stream);
      have = 4;
    }
    return (double)w[--have] * 2.3283064365386963e-10;
  }
};

__device__ int poisson(CellRng &g, double lam)
{
  if (!(lam > 0)) return 0;
  if (lam < 12.0) {
    const double L = exp(-lam);
    double p = 1.0;
    int k = 0;
    do { k++; p *= g.uniform(); } while (p > L);
    return k - 1;
  }
  // PTRS (W. Hoermann, Insurance: Mathematics and Economics 12 (1993) 39)
  const double slam = sqrt(lam), loglam = log(lam), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2);
  for (int it = 0; it < 1000; ++it) {
    const double U = g.uniform() - 0.5, V = g.uniform();
    const double us = 0.5 - fabs(U);
    const double kf = floor((2 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return (int)kf;
    if (kf < 0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1)) return (int)kf;
  }
  return (int)lam;
}

__device__ __forceinline__ double interp_r(const GhDev &d, const double *tab, double r, double at_zero)
{
  // src/cosmo.c:52-86: linear in r on the uniform radial grid
  if (r <= 0) return at_zero;
  if (r >= d.r_r2z[d.nz_tab - 1]) return tab[d.nz_tab - 1];
  const int ir = (int)(r * d.glob_idr);
  return tab[ir] + (tab[ir + 1] - tab[ir]) * (r - d.r_r2z[ir]) * d.glob_idr;
}

// src/psources.c:74-91
__device__ __forceinline__ double n_of_z(const PsDev &ps, double z)
{
  const int iz = (int)(z * ps.inv_dz);
  if (iz >= ps.nz || iz < 0) return -1;
  if (iz == ps.nz - 1) return ps.nz_arr[ps.nz - 1];
  const double zi = iz * (1.0 / ps.inv_dz);
  return ps.nz_arr[iz] + (ps.nz_arr[iz + 1] - ps.nz_arr[iz]) * (z - zi) * ps.inv_dz;
}

// Poisson mean of one cell (src/grid_tools.c:62-77); <= 0: no sources
__device__ __forceinline__ double cell_lambda(const GhDev &d, const PsDev &ps, int ix, int iy, int zg, float delta, double sigma2)
{
  const double x0 = (ix + 0.5) * d.dx - d.pos_obs[0];  // the reference subtracts pos_obs[1] here (a typo; the components are equal there)
  const double y0 = (iy + 0.5) * d.dx - d.pos_obs[1], z0 = (zg + 0.5) * d.dx - d.pos_obs[2];
  const double r = sqrt(x0 * x0 + y0 * y0 + z0 * z0);
  const double redshift = interp_r(d, d.z_r2z, r, 0.0);
  const double ndens = n_of_z(ps, redshift);
  if (!(ndens > 0)) return 0.0;
  const int izb = min(max((int)(redshift * ps.inv_dz), 0), ps.nz - 1);
  const double gfb = interp_r(d, d.gd, r, 1.0) * ps.bias_arr[izb];
  return ndens * d.dx * d.dx * d.dx * exp(gfb * ((double)delta - 0.5 * gfb * sigma2));
}

__global__ void __launch_bounds__(256) psources_poisson_kernel(GhDev d, PsDev ps, const float *__restrict__ dens, const double *__restrict__ sigma2p,
                                                               int *__restrict__ nsrc, float *__restrict__ lambda_out,
                                                               unsigned long long *__restrict__ total)
{
  const int ngx = 2 * d.nh;
  const long long ncell = (long long)d.nz_here * d.n * d.n;
  const double sigma2 = *sigma2p;
  unsigned long long mine = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(i % d.n), iy = (int)((i / d.n) % d.n), iz = (int)(i / ((long long)d.n * d.n));
    const size_t idx = ((size_t)iz * d.n + iy) * ngx + ix;
    const double lam = cell_lambda(d, ps, ix, iy, iz + d.iz0, dens[idx], sigma2);
    CellRng g(((unsigned long long)(iz + d.iz0) * d.n + iy) * d.n + ix, 0u, d.seed);
    const int np = poisson(g, lam);
    nsrc[i] = np;
    if (lambda_out) lambda_out[i] = (float)lam;
    mine += (unsigned long long)np;
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, mine);
}

// log10 L from the tabulated cumulative distribution of redshift bin iz (src/psources.c:133-154)
__device__ __forceinline__ double draw_logl(const PsDev &ps, int iz, double u)
{
  const double *cdf = ps.lcdf + (size_t)iz * (ps.nl + 1);
  int lo = 0, hi = ps.nl;  // cdf[lo] <= u < cdf[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid; else hi = mid;
  }
  const double w = cdf[hi] - cdf[lo];
  const double f = w > 0 ? (u - cdf[lo]) / w : 0.5;
  return ps.logl_min + (lo + f) * ps.dlogl;
}

__device__ __forceinline__ double sed_lookup(const PsDev &ps, double lognu)
{
  double s = (lognu - ps.lognu_min) * ps.inv_dlognu;
  s = fmin(fmax(s, 0.0), (double)(ps.nsed - 1) - 1e-9);
  const int i = (int)s;
  return ps.sed[i] + (ps.sed[i + 1] - ps.sed[i]) * (s - i);
}

#define FLUX2TEMP 3.2548291E-2 /* src/psources.c:156 */
#define LUM2FLUX 8.35774E7     /* src/psources.c:157 */

// one warp per cell: lanes take the cell's sources round-robin, every source loops over the shells
__global__ void __launch_bounds__(256) psources_maps_kernel(GhDev d, PsDev ps, const int *__restrict__ nsrc, const float *__restrict__ dzrsd,
                                                            const double *__restrict__ lognu_shell, const double *__restrict__ inv_nu2_shell,
                                                            float *__restrict__ maps)
{
  const int ngx = 2 * d.nh;
  const long long ncell = (long long)d.nz_here * d.n * d.n;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const double dOmega = 4 * 3.14159265358979323846 / (double)d.npix;
  for (long long i = warp0; i < ncell; i += nwarps) {
    const int np = nsrc[i];
    if (np <= 0) continue;
    const int ix = (int)(i % d.n), iy = (int)((i / d.n) % d.n), iz = (int)(i / ((long long)d.n * d.n));
    const double dz_rsd = (double)dzrsd[((size_t)iz * d.n + iy) * ngx + ix];
    const double x0 = d.dx * ix - d.pos_obs[0], y0 = d.dx * iy - d.pos_obs[1], z0 = d.dx * (iz + d.iz0) - d.pos_obs[2];
    const unsigned long long cell = ((unsigned long long)(iz + d.iz0) * d.n + iy) * d.n + ix;
    for (int ip = lane; ip < np; ip += 32) {
      CellRng g(cell, 1u + (uint32_t)ip, d.seed);  // one stream per source
      const double px = x0 + d.dx * g.uniform(), py = y0 + d.dx * g.uniform(), pz = z0 + d.dx * g.uniform();
      const double r = sqrt(px * px + py * py + pz * pz);
      const double red_cosmo = interp_r(d, d.z_r2z, r, 0.0), red_true = red_cosmo + dz_rsd;
      if (red_cosmo >= ps.z_max || !(r > 0)) continue;  // draw_luminosity returns 0: no temperature
      const long long ipix = gh_vec2pix_ring(d.nside, px, py, pz, r);
      const int izb = min(max((int)(red_cosmo * ps.inv_dz), 0), ps.nz - 1);
      const double l0 = pow(10.0, draw_logl(ps, izb, g.uniform()));
      // temp_of_l (src/psources.c:159-167): everything that does not depend on the shell
      const double amp = FLUX2TEMP * LUM2FLUX * 4 * 3.14159265358979323846 * l0 * ps.hhub * ps.hhub / (r * r * (1 + red_true) * dOmega);
      const double lg1z = log10(1 + red_true);
      for (int inu = 0; inu < d.n_nu; ++inu) {
        const double t = amp * sed_lookup(ps, lognu_shell[inu] + lg1z) * inv_nu2_shell[inu];
        if (t > 0) atomicAdd(maps + (size_t)d.npix * inu + ipix, (float)t);
      }
    }
  }
}

}  // namespace

struct gh_ps_state {
  PsDev ps;
  double *d_tab;      // all the small tables
  int *d_nsrc;        // [nz_here][n][n]
  float *d_lambda;    // the same, Poisson means (kept for the tests)
  float *d_maps;      // [n_nu_pad][npix]
  float *d_maps_recv; // reduce-scatter output on several ranks
  double *d_shell;    // log10 nu, 1 / nu^2 per shell
  unsigned long long *d_total;
  bool sampled;
};

static gh_ps_state *g_states[64];
static gh_cuda_ctx *g_owner[64];

static gh_ps_state *state_of(gh_cuda_ctx *c, bool create)
{
  for (int i = 0; i < 64; ++i)
    if (g_owner[i] == c) return g_states[i];
  if (!create) return nullptr;
  for (int i = 0; i < 64; ++i)
    if (!g_owner[i]) {
      g_owner[i] = c;
      g_states[i] = new gh_ps_state();
      memset(g_states[i], 0, sizeof(gh_ps_state));
      return g_states[i];
    }
  return nullptr;
}

void gh_psources_release(gh_cuda_ctx *c)
{
  for (int i = 0; i < 64; ++i)
    if (g_owner[i] == c) {
      gh_ps_state *s = g_states[i];
      cudaFree(s->d_tab); cudaFree(s->d_nsrc); cudaFree(s->d_lambda); cudaFree(s->d_maps); cudaFree(s->d_maps_recv);
      cudaFree(s->d_shell); cudaFree(s->d_total);
      delete s;
      g_owner[i] = nullptr;
      g_states[i] = nullptr;
    }
}

#define PS_REQ(cond, ...)                             \
  do {                                                \
    if (!(cond)) { gh_set_error(__VA_ARGS__); return 1; } \
  } while (0)

extern "C" int gh_cuda_get_point_sources(gh_cuda_ctx *c, const gh_cuda_psources_params *p, long long *np_total_out)
{
  PS_REQ(c && p, "gh_cuda_get_point_sources: null context or parameters");
  PS_REQ(cudaSetDevice(c->device) == cudaSuccess, "cudaSetDevice(%d) failed", c->device);
  PS_REQ(p->nz >= 2 && p->nl >= 2 && p->nsed >= 2 && p->nz_arr && p->bias_arr && p->lcdf && p->sed_arr && p->z_max > 0 &&
             p->logl_max > p->logl_min && p->lognu_max > p->lognu_min,
         "gh_cuda_get_point_sources: incomplete point-source tables");
  PS_REQ(c->sigma_ready || c->sigma_overridden, "gh_cuda_get_point_sources: the variance of the Gaussian field is not known yet "
                                                "(call gh_cuda_create_d_and_vr_fields first, before gh_cuda_get_HI)");
  gh_ps_state *s = state_of(c, true);
  PS_REQ(s, "gh_cuda_get_point_sources: too many contexts");
  const GhDev &d = c->d;
  const size_t ncell = (size_t)d.nz_here * d.n * d.n;
  const size_t ntab = (size_t)2 * p->nz + (size_t)p->nz * (p->nl + 1) + p->nsed;
  if (!s->d_tab) {
    GH_CUDA_OK(cudaMalloc(&s->d_tab, ntab * sizeof(double)));
    GH_CUDA_OK(cudaMalloc(&s->d_nsrc, ncell * sizeof(int)));
    GH_CUDA_OK(cudaMalloc(&s->d_lambda, ncell * sizeof(float)));
    GH_CUDA_OK(cudaMalloc(&s->d_total, sizeof(unsigned long long)));
    GH_CUDA_OK(cudaMalloc(&s->d_shell, 2 * sizeof(double) * d.n_nu_pad));
  }
  double *h = (double *)malloc(ntab * sizeof(double));
  PS_REQ(h, "out of host memory");
  size_t o = 0;
  memcpy(h + o, p->nz_arr, sizeof(double) * p->nz); const size_t o_nz = o; o += p->nz;
  memcpy(h + o, p->bias_arr, sizeof(double) * p->nz); const size_t o_b = o; o += p->nz;
  memcpy(h + o, p->lcdf, sizeof(double) * p->nz * (p->nl + 1)); const size_t o_c = o; o += (size_t)p->nz * (p->nl + 1);
  memcpy(h + o, p->sed_arr, sizeof(double) * p->nsed); const size_t o_s = o;
  cudaError_t e = cudaMemcpy(s->d_tab, h, ntab * sizeof(double), cudaMemcpyHostToDevice);
  free(h);
  PS_REQ(e == cudaSuccess, "gh_cuda_get_point_sources: %s", cudaGetErrorString(e));
  PsDev &ps = s->ps;
  ps.nz = p->nz; ps.z_max = p->z_max; ps.inv_dz = p->nz / p->z_max;
  ps.nz_arr = s->d_tab + o_nz; ps.bias_arr = s->d_tab + o_b;
  ps.nl = p->nl; ps.logl_min = p->logl_min; ps.dlogl = (p->logl_max - p->logl_min) / p->nl; ps.lcdf = s->d_tab + o_c;
  ps.nsed = p->nsed; ps.lognu_min = p->lognu_min; ps.inv_dlognu = (p->nsed - 1) / (p->lognu_max - p->lognu_min); ps.sed = s->d_tab + o_s;
  ps.hhub = p->hhub;
  GH_CUDA_OK(cudaMemsetAsync(s->d_total, 0, sizeof(unsigned long long), c->stream));
  long long blocks = ((long long)ncell + 255) / 256;
  if (blocks > (long long)c->n_sm * 32) blocks = (long long)c->n_sm * 32;
  // the Gaussian density is still in grid A (get_HI has not run yet); its variance sits on the device
  psources_poisson_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(d, ps, reinterpret_cast<const float *>(c->gridA),
                                                                  c->d_partials + (c->sigma_overridden ? 6 : 5), s->d_nsrc, s->d_lambda, s->d_total);
  GH_LAUNCH_CHECK(c);
  unsigned long long tot = 0;
  GH_CUDA_OK(cudaMemcpyAsync(&tot, s->d_total, sizeof(tot), cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  if (d.nranks > 1) {
    GH_CUDA_OK(cudaMemcpyAsync(s->d_total, &tot, sizeof(tot), cudaMemcpyHostToDevice, c->stream));
    GH_NCCL_OK(ncclAllReduce(s->d_total, s->d_total, 1, ncclUint64, ncclSum, c->comm, c->stream));
    GH_CUDA_OK(cudaMemcpyAsync(&tot, s->d_total, sizeof(tot), cudaMemcpyDeviceToHost, c->stream));
    GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  }
  if (np_total_out) *np_total_out = (long long)tot;
  s->sampled = true;
  return 0;
}

extern "C" int gh_cuda_mk_psources_maps(gh_cuda_ctx *c, float *maps_ps_host)
{
  PS_REQ(c, "null gh_cuda context");
  PS_REQ(cudaSetDevice(c->device) == cudaSuccess, "cudaSetDevice(%d) failed", c->device);
  gh_ps_state *s = state_of(c, false);
  PS_REQ(s && s->sampled, "gh_cuda_mk_psources_maps: call gh_cuda_get_point_sources first");
  const GhDev &d = c->d;
  const size_t stack = (size_t)d.n_nu_pad * d.npix;
  int n_here = 0, s0 = 0;
  if (gh_cuda_shells(c, &n_here, &s0)) return 1;
  if (!s->d_maps) {
    GH_CUDA_OK(cudaMalloc(&s->d_maps, stack * sizeof(float)));
    if (d.nranks > 1) GH_CUDA_OK(cudaMalloc(&s->d_maps_recv, stack / d.nranks * sizeof(float)));
  }
  // shell centres (src/pixelize.c:63-70)
  double hs[2 * 4096];
  for (int i = 0; i < d.n_nu; ++i) {
    const double nu = c->h_nu_centre[i];
    hs[i] = log10(nu);
    hs[d.n_nu_pad + i] = 1.0 / (nu * nu);
  }
  GH_CUDA_OK(cudaMemcpyAsync(s->d_shell, hs, 2 * sizeof(double) * d.n_nu_pad, cudaMemcpyHostToDevice, c->stream));
  GH_CUDA_OK(cudaMemsetAsync(s->d_maps, 0, stack * sizeof(float), c->stream));
  const long long ncell = (long long)d.nz_here * d.n * d.n;
  long long blocks = (ncell * 32 + 255) / 256;
  if (blocks > (long long)c->n_sm * 16) blocks = (long long)c->n_sm * 16;
  // Delta z_RSD is in grid C once get_HI has run (src/pixelize.c:104 reads grid_rvel after get_HI)
  psources_maps_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(d, s->ps, s->d_nsrc, reinterpret_cast<const float *>(c->gridC), s->d_shell,
                                                               s->d_shell + d.n_nu_pad, s->d_maps);
  GH_LAUNCH_CHECK(c);
  const float *result = s->d_maps;
  if (d.nranks > 1) {
    GH_NCCL_OK(ncclReduceScatter(s->d_maps, s->d_maps_recv, stack / d.nranks, ncclFloat, ncclSum, c->comm, c->stream));
    result = s->d_maps_recv;
  }
  if (maps_ps_host && n_here > 0)
    GH_CUDA_OK(cudaMemcpyAsync(maps_ps_host, result, (size_t)n_here * d.npix * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gh_cuda_download_point_sources(gh_cuda_ctx *c, int *nsources_out, float *lambda_out)
{
  PS_REQ(c, "null gh_cuda context");
  PS_REQ(cudaSetDevice(c->device) == cudaSuccess, "cudaSetDevice(%d) failed", c->device);
  gh_ps_state *s = state_of(c, false);
  PS_REQ(s && s->sampled, "gh_cuda_download_point_sources: call gh_cuda_get_point_sources first");
  const size_t ncell = (size_t)c->d.nz_here * c->d.n * c->d.n;
  if (nsources_out) GH_CUDA_OK(cudaMemcpyAsync(nsources_out, s->d_nsrc, ncell * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (lambda_out) GH_CUDA_OK(cudaMemcpyAsync(lambda_out, s->d_lambda, ncell * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}
