// Real-space field kernels: radial velocity, Gaussian variance, lognormal / HI-mass / RSD transform.
// One HBM round trip each.  Float arithmetic throughout (the reference computes in double from float
// loads and stores floats; parity is rel <= 1e-5 on the stored floats), double only for the variance sums.
#include "gh_internal.cuh"
#include "gh_gethi_math.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// radial_velocity_from_potential (reference src/fourier.c:307-373): v = +grad(phi) by central
// differences, periodic in x and y, neighbour planes in z come from the adjacent slabs
// (src/fourier.c:415-428), projected on the line of sight from the observer.
// Two cells per thread (rows are 8-byte aligned), x neighbours through warp shuffles, y/z neighbours as
// float2 loads that hit L1/L2 (three planes of phi stay in the 126 MB L2 while a plane is swept), so
// DRAM sees ~4 B read + 4 B written per cell.  One CTA = one row segment of 2*blockDim cells.
__global__ void __launch_bounds__(256) radial_velocity_kernel(GhDev d, Axes3 axes, const float *__restrict__ vpot,
                                                              const float *__restrict__ plane_lo,
                                                              const float *__restrict__ plane_hi,
                                                              float *__restrict__ rvel)
{
  const int ngx = 2 * d.nh;
  const int iy = blockIdx.y, iz = blockIdx.z;
  const float hidx = d.half_inv_dx;
  const AxisF ax = axes.x;
  const float y = axes.y.at(iy), z = axes.z.at(iz);
  const int iy_hi = (iy == d.n - 1) ? 0 : iy + 1, iy_lo = (iy == 0) ? d.n - 1 : iy - 1;
  const size_t plane = (size_t)ngx * d.n;
  const float *p0 = vpot + (size_t)iz * plane;
  const float *pz_lo = ((iz == 0) ? plane_lo : p0 - plane) + (size_t)iy * ngx;
  const float *pz_hi = ((iz == d.nz_here - 1) ? plane_hi : p0 + plane) + (size_t)iy * ngx;
  const float *row = p0 + (size_t)iy * ngx, *row_hi = p0 + (size_t)iy_hi * ngx, *row_lo = p0 + (size_t)iy_lo * ngx;
  float *out = rvel + (size_t)iz * plane + (size_t)iy * ngx;
  const int lane = threadIdx.x & 31;
  const float yz2 = fmaf(y, y, z * z);
  const int nper = d.n / 2;
  // every lane of a warp runs the same number of iterations (the shuffles need the whole warp)
  for (int base = blockIdx.x * blockDim.x; base < nper; base += gridDim.x * blockDim.x) {
    const int ip = base + threadIdx.x;
    const bool act = ip < nper;
    const int ix = act ? 2 * ip : 0;
    const float2 c = __ldg(reinterpret_cast<const float2 *>(row + ix));
    const float2 yh = __ldg(reinterpret_cast<const float2 *>(row_hi + ix)), yl = __ldg(reinterpret_cast<const float2 *>(row_lo + ix));
    const float2 zh = __ldg(reinterpret_cast<const float2 *>(pz_hi + ix)), zl = __ldg(reinterpret_cast<const float2 *>(pz_lo + ix));
    // x neighbours: the left cell of my pair needs phi[ix-1] (previous lane's .y), the right one phi[ix+2]
    float left = __shfl_up_sync(0xffffffffu, c.y, 1), right = __shfl_down_sync(0xffffffffu, c.x, 1);
    if (lane == 0) left = __ldg(row + ((ix == 0) ? d.n - 1 : ix - 1));
    if (lane == 31 || ip >= nper - 1) right = __ldg(row + ((ix + 2 >= d.n) ? 0 : ix + 2));
    if (!act) continue;
    const float x0 = ax.at(ix), x1 = ax.at(ix + 1);
    const float vx0 = hidx * (c.y - left), vx1 = hidx * (right - c.x);
    const float vy0 = hidx * (yh.x - yl.x), vy1 = hidx * (yh.y - yl.y);
    const float vz0 = hidx * (zh.x - zl.x), vz1 = hidx * (zh.y - zl.y);
    const float ir0 = rsqrtf(fmaf(x0, x0, yz2)), ir1 = rsqrtf(fmaf(x1, x1, yz2));
    float2 r;
    r.x = fmaf(vx0, x0, fmaf(vy0, y, vz0 * z)) * ir0;
    r.y = fmaf(vx1, x1, fmaf(vy1, y, vz1 * z)) * ir1;
    *reinterpret_cast<float2 *>(out + ix) = r;
  }
}

// ------------------------------------------------------------------------------------------------
// compute_sigma_dens (src/fourier.c:24-76): sum d and d*d (float product, double accumulation, exactly
// the reference's promotion) over the real cells of the slab (padding skipped).  Two-stage and
// deterministic: per-CTA partials, then one CTA folds them.
__global__ void __launch_bounds__(256) sigma_partial_kernel(GhDev d, const float *__restrict__ dens,
                                                            double *__restrict__ partials)
{
  const int ngx = 2 * d.nh;
  const long long nrows = (long long)d.nz_here * d.n;
  double s1 = 0.0, s2 = 0.0;
  const int half = d.n / 2;  // rows are 8-byte aligned: read float2
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const float2 *r = reinterpret_cast<const float2 *>(dens + row * ngx);
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
      const float2 v = __ldg(r + i);
      s1 += (double)v.x + (double)v.y;
      s2 += (double)__fmul_rn(v.x, v.x) + (double)__fmul_rn(v.y, v.y);
    }
  }
  __shared__ double sh1[256], sh2[256];
  sh1[threadIdx.x] = s1;
  sh2[threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh1[threadIdx.x] += sh1[threadIdx.x + o];
      sh2[threadIdx.x] += sh2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partials[GH_PARTIALS_BASE + 2 * blockIdx.x] = sh1[0];
    partials[GH_PARTIALS_BASE + 1 + 2 * blockIdx.x] = sh2[0];
  }
}

__global__ void __launch_bounds__(256) sigma_final_kernel(double *__restrict__ partials, int nblocks)
{
  __shared__ double sh1[256], sh2[256];
  double s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    s1 += partials[GH_PARTIALS_BASE + 2 * i];
    s2 += partials[GH_PARTIALS_BASE + 1 + 2 * i];
  }
  sh1[threadIdx.x] = s1;
  sh2[threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh1[threadIdx.x] += sh1[threadIdx.x + o];
      sh2[threadIdx.x] += sh2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partials[0] = sh1[0];
    partials[1] = sh2[0];
  }
}

// mean and variance from the (all-reduced) sums: partials[4] = <d>, partials[5] = <d^2> - <d>^2 (src/fourier.c:59-74)
// The four numbers the host wants (sum, sum of squares, mean, variance) are stored straight into mapped pinned
// host memory: a cudaMemcpyAsync here would queue behind the previous realisation's 200 MB map download on the
// device->host copy engine and stall the compute stream for ~1 ms (measured).  A caller-supplied sigma2_gauss
// (gh_cuda_set_sigma2_gauss) lives in its own slot, partials[6], which nothing here writes.
__global__ void sigma_finish_kernel(double *__restrict__ partials, double inv_ng_tot, double *__restrict__ host_stats)
{
  const double sum = partials[0], sumsq = partials[1];
  const double mean = sum * inv_ng_tot, var = sumsq * inv_ng_tot - mean * mean;
  partials[4] = mean;
  partials[5] = var;
  host_stats[0] = sum; host_stats[1] = sumsq; host_stats[2] = mean; host_stats[3] = var;
  __threadfence_system();
}

// ------------------------------------------------------------------------------------------------
// get_HI (src/grid_tools.c:103-153) with z_of_r / dgrowth_of_r / vgrowth_of_r (src/cosmo.c:52-86) and
// bias_HI / fraction_HI (src/user_defined.c:27-35), in place: dens <- HI mass, rvel <- Delta z_RSD.
// 8 B read + 8 B written per cell; the three 5001-entry float tables are read through L1.
__global__ void __launch_bounds__(256) get_HI_kernel(GhDev d, Axes3 axes, float *__restrict__ dens, float *__restrict__ rvel,
                                                     const double *__restrict__ sigma2_ptr)
{
  const GetHIConsts k = make_gethi_consts(d, (float)*sigma2_ptr);  // the variance stays on the device
  const int ngx = 2 * d.nh;
  const int iy = blockIdx.y, iz = blockIdx.z;
  const AxisF ax = axes.x;
  const float y = axes.y.at(iy), z = axes.z.at(iz);
  const float yz2 = fmaf(y, y, __fmul_rn(z, z));
  const size_t base = ((size_t)iz * d.n + iy) * ngx;
  // two cells per thread: rows are 8-byte aligned
  for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < d.n / 2; ip += gridDim.x * blockDim.x) {
    const float2 dv = *reinterpret_cast<const float2 *>(dens + base + 2 * ip);
    const float2 vv = *reinterpret_cast<const float2 *>(rvel + base + 2 * ip);
    const float x0 = ax.at(2 * ip), x1 = ax.at(2 * ip + 1);
    float2 m, dzr;
    gh_gethi_cell(k, fmaf(x0, x0, yz2), dv.x, vv.x, m.x, dzr.x);
    gh_gethi_cell(k, fmaf(x1, x1, yz2), dv.y, vv.y, m.y, dzr.y);
    *reinterpret_cast<float2 *>(dens + base + 2 * ip) = m;
    *reinterpret_cast<float2 *>(rvel + base + 2 * ip) = dzr;
  }
}

// ------------------------------------------------------------------------------------------------
// radial velocity + get_HI in one pass (used by gh_cuda_run / gh_cuda_run_async, where nobody looks at the
// radial velocity itself): the velocity never goes to memory, so the pair costs 8 B read (phi, delta) + 8 B
// written (HI mass, Delta z_RSD) per cell instead of 24.  Same thread mapping and the very same expressions as
// the two kernels above, so the results are bit-identical to running them one after the other
// (tests/test_gpu_parity.py).
__global__ void __launch_bounds__(256) velocity_get_HI_kernel(GhDev d, Axes3 axes, const float *__restrict__ vpot,
                                                              const float *__restrict__ plane_lo,
                                                              const float *__restrict__ plane_hi, float *__restrict__ dens,
                                                              float *__restrict__ dz_out, const double *__restrict__ sigma2_ptr)
{
  const GetHIConsts k = make_gethi_consts(d, (float)*sigma2_ptr);
  const int ngx = 2 * d.nh;
  const int iy = blockIdx.y, iz = blockIdx.z;
  const float hidx = d.half_inv_dx;
  const AxisF ax = axes.x;
  const float y = axes.y.at(iy), z = axes.z.at(iz);
  const int iy_hi = (iy == d.n - 1) ? 0 : iy + 1, iy_lo = (iy == 0) ? d.n - 1 : iy - 1;
  const size_t plane = (size_t)ngx * d.n;
  const float *p0 = vpot + (size_t)iz * plane;
  const float *pz_lo = ((iz == 0) ? plane_lo : p0 - plane) + (size_t)iy * ngx;
  const float *pz_hi = ((iz == d.nz_here - 1) ? plane_hi : p0 + plane) + (size_t)iy * ngx;
  const float *row = p0 + (size_t)iy * ngx, *row_hi = p0 + (size_t)iy_hi * ngx, *row_lo = p0 + (size_t)iy_lo * ngx;
  const size_t base_off = (size_t)iz * plane + (size_t)iy * ngx;
  const int lane = threadIdx.x & 31;
  const float yz2 = fmaf(y, y, z * z);
  const float yz2_h = fmaf(y, y, __fmul_rn(z, z));  // get_HI_kernel's spelling of the same sum
  const int nper = d.n / 2;
  for (int base = blockIdx.x * blockDim.x; base < nper; base += gridDim.x * blockDim.x) {
    const int ip = base + threadIdx.x;
    const bool act = ip < nper;
    const int ix = act ? 2 * ip : 0;
    const float2 c = __ldg(reinterpret_cast<const float2 *>(row + ix));
    const float2 yh = __ldg(reinterpret_cast<const float2 *>(row_hi + ix)), yl = __ldg(reinterpret_cast<const float2 *>(row_lo + ix));
    const float2 zh = __ldg(reinterpret_cast<const float2 *>(pz_hi + ix)), zl = __ldg(reinterpret_cast<const float2 *>(pz_lo + ix));
    const float2 dv = *reinterpret_cast<const float2 *>(dens + base_off + ix);
    float left = __shfl_up_sync(0xffffffffu, c.y, 1), right = __shfl_down_sync(0xffffffffu, c.x, 1);
    if (lane == 0) left = __ldg(row + ((ix == 0) ? d.n - 1 : ix - 1));
    if (lane == 31 || ip >= nper - 1) right = __ldg(row + ((ix + 2 >= d.n) ? 0 : ix + 2));
    if (!act) continue;
    const float x0 = ax.at(ix), x1 = ax.at(ix + 1);
    const float vx0 = hidx * (c.y - left), vx1 = hidx * (right - c.x);
    const float vy0 = hidx * (yh.x - yl.x), vy1 = hidx * (yh.y - yl.y);
    const float vz0 = hidx * (zh.x - zl.x), vz1 = hidx * (zh.y - zl.y);
    const float ir0 = rsqrtf(fmaf(x0, x0, yz2)), ir1 = rsqrtf(fmaf(x1, x1, yz2));
    const float rv0 = fmaf(vx0, x0, fmaf(vy0, y, vz0 * z)) * ir0;
    const float rv1 = fmaf(vx1, x1, fmaf(vy1, y, vz1 * z)) * ir1;
    float2 m, dzr;
    gh_gethi_cell(k, fmaf(x0, x0, yz2_h), dv.x, rv0, m.x, dzr.x);
    gh_gethi_cell(k, fmaf(x1, x1, yz2_h), dv.y, rv1, m.y, dzr.y);
    *reinterpret_cast<float2 *>(dens + base_off + ix) = m;
    *reinterpret_cast<float2 *>(dz_out + base_off + ix) = dzr;
  }
}

// ------------------------------------------------------------------------------------------------
// Position-weighted checksum of the real cells of a plane range: sum over cells of bits(value) * (2*g + 1) mod 2^64
// with g the cell's global index (zg*n + iy)*n + ix.  Integer addition commutes, so the result does not depend on
// the order of the atomics; two slab decompositions of a bit-identical field give identical per-plane-range sums
// (tests: 2048^3 on eight GPUs against one GPU without moving 100 GB of grids through the host).
__global__ void __launch_bounds__(256) checksum_kernel(const uint32_t *__restrict__ grid, int n, int ngx, int nplanes,
                                                       long long zg0, unsigned long long *__restrict__ out)
{
  const long long nrows = (long long)nplanes * n;
  unsigned long long acc = 0ULL;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const unsigned long long g0 = (unsigned long long)((zg0 + row / n) * n + row % n) * (unsigned long long)n;
    const uint32_t *r = grid + row * ngx;
    for (int ix = threadIdx.x; ix < n; ix += blockDim.x) acc += (unsigned long long)__ldg(r + ix) * (2ULL * (g0 + ix) + 1ULL);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

}  // namespace

int gh_launch_checksum(gh_cuda_ctx *c, const float *grid, int z0_local, int nplanes, unsigned long long *d_out)
{
  const GhDev &d = c->d;
  const int ngx = 2 * d.nh;
  long long blocks = (long long)nplanes * d.n;
  if (blocks > c->n_sm * 16) blocks = c->n_sm * 16;
  checksum_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(reinterpret_cast<const uint32_t *>(grid) + (size_t)z0_local * d.n * ngx, d.n, ngx,
                                                           nplanes, (long long)d.iz0 + z0_local, d_out);
  GH_LAUNCH_CHECK(c);
  return 0;
}

// neighbour planes of the velocity potential for this slab's first and last plane
static int halo_planes(gh_cuda_ctx *c, const float **lo_out, const float **hi_out)
{
  const GhDev &d = c->d;
  const float *vpot = reinterpret_cast<const float *>(c->gridB);
  const size_t plane = (size_t)2 * d.nh * d.n;
  const float *lo, *hi;
  if (d.nranks > 1) {
    // halo exchange (src/fourier.c:415-424): last plane -> right neighbour, first plane -> left neighbour
    const int right = (d.rank + 1) % d.nranks, left = (d.rank + d.nranks - 1) % d.nranks;
    GH_NCCL_OK(ncclGroupStart());
    GH_NCCL_OK(ncclSend(vpot + (size_t)(d.nz_here - 1) * plane, plane, ncclFloat, right, c->comm, c->stream));
    GH_NCCL_OK(ncclRecv(c->halo_lo, plane, ncclFloat, left, c->comm, c->stream));
    GH_NCCL_OK(ncclSend(vpot, plane, ncclFloat, left, c->comm, c->stream));
    GH_NCCL_OK(ncclRecv(c->halo_hi, plane, ncclFloat, right, c->comm, c->stream));
    GH_NCCL_OK(ncclGroupEnd());
    lo = c->halo_lo;
    hi = c->halo_hi;
  } else {
    lo = vpot + (size_t)(d.n - 1) * plane;  // src/fourier.c:425-427
    hi = vpot;
  }
  *lo_out = lo;
  *hi_out = hi;
  return 0;
}

int gh_launch_halo_exchange(gh_cuda_ctx *c)
{
  const float *lo, *hi;
  return halo_planes(c, &lo, &hi);
}

// call after gh_launch_halo_exchange on several ranks (the exchange is not repeated here)
int gh_launch_velocity_get_HI(gh_cuda_ctx *c)
{
  const GhDev &d = c->d;
  const float *vpot = reinterpret_cast<const float *>(c->gridB);
  const size_t plane = (size_t)2 * d.nh * d.n;
  const float *lo = (d.nranks > 1) ? c->halo_lo : vpot + (size_t)(d.n - 1) * plane;
  const float *hi = (d.nranks > 1) ? c->halo_hi : vpot;
  dim3 grid((d.n / 2 + 255) / 256, d.n, d.nz_here);
  const Axes3 axes = {make_axis(d.dx, d.pos_obs[0], 0), make_axis(d.dx, d.pos_obs[1], 0), make_axis(d.dx, d.pos_obs[2], d.iz0)};
  velocity_get_HI_kernel<<<grid, 256, 0, c->stream>>>(d, axes, vpot, lo, hi, reinterpret_cast<float *>(c->gridA),
                                                      reinterpret_cast<float *>(c->gridC), c->d_partials + (c->sigma_overridden ? 6 : 5));
  GH_LAUNCH_CHECK(c);
  return 0;
}

int gh_launch_radial_velocity(gh_cuda_ctx *c)
{
  const GhDev &d = c->d;
  const float *vpot = reinterpret_cast<const float *>(c->gridB);
  const float *lo, *hi;
  if (halo_planes(c, &lo, &hi)) return 1;
  dim3 grid((d.n / 2 + 255) / 256, d.n, d.nz_here);
  const Axes3 axes = {make_axis(d.dx, d.pos_obs[0], 0), make_axis(d.dx, d.pos_obs[1], 0), make_axis(d.dx, d.pos_obs[2], d.iz0)};
  radial_velocity_kernel<<<grid, 256, 0, c->stream>>>(d, axes, vpot, lo, hi, reinterpret_cast<float *>(c->gridC));
  GH_LAUNCH_CHECK(c);
  return 0;
}

int gh_launch_sigma(gh_cuda_ctx *c)
{
  const GhDev &d = c->d;
  if (c->fft_stats_blocks > 0) {
    // the density FFT's last pass already produced the per-CTA partial sums
    sigma_final_kernel<<<1, 256, 0, c->stream>>>(c->d_partials, c->fft_stats_blocks);
    GH_LAUNCH_CHECK(c);
    return 0;
  }
  int blocks = c->n_sm * 8;
  const long long nrows = (long long)d.nz_here * d.n;
  if (blocks > nrows) blocks = (int)nrows;
  sigma_partial_kernel<<<blocks, 256, 0, c->stream>>>(d, reinterpret_cast<const float *>(c->gridA), c->d_partials);
  GH_LAUNCH_CHECK(c);
  sigma_final_kernel<<<1, 256, 0, c->stream>>>(c->d_partials, blocks);
  GH_LAUNCH_CHECK(c);
  return 0;
}

int gh_launch_sigma_finish(gh_cuda_ctx *c)
{
  const double ng_tot = (double)c->d.n * ((double)c->d.n * (double)c->d.n);
  sigma_finish_kernel<<<1, 1, 0, c->stream>>>(c->d_partials, 1.0 / ng_tot, c->h_stats_dev);
  GH_LAUNCH_CHECK(c);
  return 0;
}

int gh_launch_get_HI(gh_cuda_ctx *c)
{
  const GhDev &d = c->d;
  dim3 grid((d.n / 2 + 255) / 256, d.n, d.nz_here);
  const Axes3 axes = {make_axis(d.dx, d.pos_obs[0], 0), make_axis(d.dx, d.pos_obs[1], 0), make_axis(d.dx, d.pos_obs[2], d.iz0)};
  get_HI_kernel<<<grid, 256, 0, c->stream>>>(d, axes, reinterpret_cast<float *>(c->gridA), reinterpret_cast<float *>(c->gridC),
                                             c->d_partials + (c->sigma_overridden ? 6 : 5));
  GH_LAUNCH_CHECK(c);
  return 0;
}
