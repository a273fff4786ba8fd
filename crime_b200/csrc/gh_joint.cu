// JoinT's ingestion of GetHI output, kept on the device: merge_maps (reference src/main_jt.c:98-211)
// sums the component map stacks shell by shell in a fixed order and hands each sum to he_udgrade
// (src/healpix_extra.c:318-385) to change its resolution.  The cosmological signal is the stack gh_cuda_mk_T_maps
// has just left in device memory, so it enters the sum from there instead of going through FITS files and back.
//
// he_udgrade on RING maps walks the NEST hierarchy: output pixel -> ring2nest -> its ratio = (nside_in / nside_out)^2
// children -> nest2ring -> input pixels, summed in double in child order, times 1/ratio, rounded to float once.  The
// kernels below do exactly that, one thread per output pixel, so the result is bit-identical to the reference's.
// RING <-> NEST is the standard HEALPix bijection (chealpix nest2ring / ring2nest): face number, (ix, iy) bit
// interleave, jrll / jpll offsets.
#include "gh_internal.cuh"

namespace {

__device__ __constant__ int kJrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
__device__ __constant__ int kJpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};

// every second bit of v, packed (v < 2^32)
__device__ __forceinline__ unsigned compress_bits(unsigned long long v)
{
  v &= 0x5555555555555555ULL;
  v = (v | (v >> 1)) & 0x3333333333333333ULL;
  v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
  v = (v | (v >> 4)) & 0x00FF00FF00FF00FFULL;
  v = (v | (v >> 8)) & 0x0000FFFF0000FFFFULL;
  v = (v | (v >> 16)) & 0x00000000FFFFFFFFULL;
  return (unsigned)v;
}
__device__ __forceinline__ unsigned long long spread_bits(unsigned v)
{
  unsigned long long x = v;
  x = (x | (x << 16)) & 0x0000FFFF0000FFFFULL;
  x = (x | (x << 8)) & 0x00FF00FF00FF00FFULL;
  x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0FULL;
  x = (x | (x << 2)) & 0x3333333333333333ULL;
  x = (x | (x << 1)) & 0x5555555555555555ULL;
  return x;
}

__device__ __forceinline__ long long isqrt_ll(long long v)
{
  long long r = (long long)sqrt((double)v + 0.5);
  while (r * r > v) --r;
  while ((r + 1) * (r + 1) <= v) ++r;
  return r;
}

__device__ long long nest2ring(long long nside, long long ipnest)
{
  const long long npface = nside * nside, npix = 12 * npface, nl4 = 4 * nside, ncap = 2 * nside * (nside - 1);
  const int face = (int)(ipnest / npface);
  const unsigned long long ipf = (unsigned long long)(ipnest - face * npface);
  const long long ix = compress_bits(ipf), iy = compress_bits(ipf >> 1);
  const long long jr = kJrll[face] * nside - ix - iy - 1;
  long long nr = nside, n_before = ncap + nl4 * (jr - nside), kshift = (jr - nside) & 1;
  if (jr < nside) { nr = jr; n_before = 2 * nr * (nr - 1); kshift = 0; }
  else if (jr > 3 * nside) { nr = nl4 - jr; n_before = npix - 2 * (nr + 1) * nr; kshift = 0; }
  long long jp = (kJpll[face] * nr + ix - iy + 1 + kshift) / 2;
  if (jp > nl4) jp -= nl4;
  if (jp < 1) jp += nl4;
  return n_before + jp - 1;
}

__device__ long long ring2nest(long long nside, long long ipring)
{
  const long long npix = 12 * nside * nside, nl2 = 2 * nside, nl4 = 4 * nside, ncap = 2 * nside * (nside - 1);
  long long irn, iphi, nr, kshift;
  int face;
  if (ipring < ncap) {  // north polar cap
    irn = (1 + isqrt_ll(1 + 2 * ipring)) / 2;
    iphi = ipring + 1 - 2 * irn * (irn - 1);
    kshift = 0;
    nr = irn;
    face = (int)((iphi - 1) / nr);
  } else if (ipring < npix - ncap) {  // equatorial belt
    const long long ip = ipring - ncap;
    irn = ip / nl4 + nside;
    iphi = ip % nl4 + 1;
    kshift = (irn + nside) & 1;
    nr = nside;
    const long long ire = irn - nside + 1, irm = nl2 + 2 - ire;
    const long long ifm = (iphi - ire / 2 + nside - 1) / nside, ifp = (iphi - irm / 2 + nside - 1) / nside;
    if (ifp == ifm) face = (ifp == 4) ? 4 : (int)ifp + 4;
    else if (ifp < ifm) face = (int)ifp;
    else face = (int)ifm + 8;
  } else {  // south polar cap
    const long long ip = npix - ipring;
    const long long irs = (1 + isqrt_ll(2 * ip - 1)) / 2;
    iphi = 4 * irs + 1 - (ip - 2 * irs * (irs - 1));
    kshift = 0;
    nr = irs;
    irn = nl4 - irs;
    face = (int)((iphi - 1) / nr) + 8;
  }
  const long long irt = irn - kJrll[face] * nside + 1;
  long long ipt = 2 * iphi - kJpll[face] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  const long long ix = (ipt - irt) / 2, iy = (-(ipt + irt)) / 2;
  return face * nside * nside + (long long)(spread_bits((unsigned)ix) + 2 * spread_bits((unsigned)iy));
}

// one thread per output pixel (src/healpix_extra.c:318-385)
__global__ void __launch_bounds__(256) udgrade_kernel(const float *__restrict__ in, long long nside_in, float *__restrict__ out,
                                                      long long nside_out, int nest)
{
  const long long npix_out = 12 * nside_out * nside_out;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < npix_out; o += (long long)gridDim.x * blockDim.x) {
    if (nside_in == nside_out) {
      out[o] = in[o];
    } else if (nside_in > nside_out) {
      const long long ratio = (nside_in / nside_out) * (nside_in / nside_out);
      const double inv = 1. / ((double)ratio);
      const long long base = ratio * (nest ? o : ring2nest(nside_out, o));
      double tot = 0;
      for (long long j = 0; j < ratio; ++j) tot += (double)in[nest ? base + j : nest2ring(nside_in, base + j)];
      out[o] = (float)(tot * inv);
    } else {
      const long long ratio = (nside_out / nside_in) * (nside_out / nside_in);
      const long long parent = (nest ? o : ring2nest(nside_out, o)) / ratio;
      out[o] = in[nest ? parent : nest2ring(nside_in, parent)];
    }
  }
}

// map_result += map_to_sum (src/main_jt.c:84-96), after the optional float *= double of the leakage term (:183)
__global__ void __launch_bounds__(256) add_maps_kernel(float *__restrict__ acc, const float *__restrict__ src, double scale, int scaled,
                                                       long long n)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = src[i];
    if (scaled) v = (float)((double)v * scale);
    acc[i] += v;
  }
}

__global__ void __launch_bounds__(128) nestring_kernel(long long nside, const long long *__restrict__ in, long long *__restrict__ out,
                                                       long long n, int to_ring)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = to_ring ? nest2ring(nside, in[i]) : ring2nest(nside, in[i]);
}

bool pow2(long v) { return v > 0 && (v & (v - 1)) == 0; }

int blocks_for(gh_cuda_ctx *c, long long n, int nt)
{
  long long b = (n + nt - 1) / nt;
  const long long cap = (long long)c->n_sm * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

#define GH_CTX_J(c)                                                  \
  do {                                                               \
    if (!(c)) { gh_set_error("null gh_cuda context"); return 1; }    \
    if (cudaSetDevice((c)->device) != cudaSuccess) { gh_set_error("cudaSetDevice(%d) failed", (c)->device); return 1; } \
  } while (0)
#define GH_REQ_J(cond, ...)                          \
  do {                                               \
    if (!(cond)) { gh_set_error(__VA_ARGS__); return 1; } \
  } while (0)

extern "C" int gh_cuda_udgrade(gh_cuda_ctx *c, const float *maps_in, long nside_in, float *maps_out, long nside_out, int nest, int n_maps)
{
  GH_CTX_J(c);
  GH_REQ_J(maps_in && maps_out && n_maps >= 0, "gh_cuda_udgrade: null map or negative count");
  GH_REQ_J(pow2(nside_in) && pow2(nside_out) && nside_in <= 8192 && nside_out <= 8192, "gh_cuda_udgrade: nside must be a power of two <= 8192 (%ld -> %ld)",
           nside_in, nside_out);
  const size_t npi = 12 * (size_t)nside_in * nside_in, npo = 12 * (size_t)nside_out * nside_out;
  float *d_in = nullptr, *d_out = nullptr;
  GH_CUDA_OK(cudaMalloc(&d_in, npi * sizeof(float)));
  if (cudaMalloc(&d_out, npo * sizeof(float)) != cudaSuccess) { cudaFree(d_in); gh_set_error("gh_cuda_udgrade: out of device memory"); return 1; }
  int rc = 0;
  for (int m = 0; m < n_maps && !rc; ++m) {
    rc = cudaMemcpyAsync(d_in, maps_in + (size_t)m * npi, npi * sizeof(float), cudaMemcpyHostToDevice, c->stream) != cudaSuccess;
    udgrade_kernel<<<blocks_for(c, (long long)npo, 256), 256, 0, c->stream>>>(d_in, nside_in, d_out, nside_out, nest);
    c->launches++;
    rc = rc || cudaMemcpyAsync(maps_out + (size_t)m * npo, d_out, npo * sizeof(float), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess;
    rc = rc || cudaStreamSynchronize(c->stream) != cudaSuccess;
  }
  cudaFree(d_in);
  cudaFree(d_out);
  if (rc) gh_set_error("gh_cuda_udgrade: %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}

extern "C" int gh_cuda_jt_merge_maps(gh_cuda_ctx *c, int n_comp, const float *const *comp_host, const double *scale, long nside_out,
                                     float *out_host)
{
  GH_CTX_J(c);
  const GhDev &d = c->d;
  GH_REQ_J(n_comp >= 1 && comp_host && out_host, "gh_cuda_jt_merge_maps: no components or null output");
  GH_REQ_J(pow2(nside_out) && nside_out <= 8192 && pow2((long)d.nside), "gh_cuda_jt_merge_maps: nside must be a power of two <= 8192");
  int n_here = 0, s0 = 0;
  if (gh_cuda_shells(c, &n_here, &s0)) return 1;
  const size_t npi = (size_t)d.npix, npo = 12 * (size_t)nside_out * nside_out;
  const float *own = c->out_buf[c->out_cur];  // this rank's shells as gh_cuda_mk_T_maps left them
  float *d_acc = nullptr, *d_comp = nullptr, *d_out = nullptr;
  GH_CUDA_OK(cudaMalloc(&d_acc, npi * sizeof(float)));
  if (cudaMalloc(&d_comp, npi * sizeof(float)) != cudaSuccess || cudaMalloc(&d_out, npo * sizeof(float)) != cudaSuccess) {
    cudaFree(d_acc); cudaFree(d_comp);
    gh_set_error("gh_cuda_jt_merge_maps: out of device memory");
    return 1;
  }
  int rc = 0;
  const int nb = blocks_for(c, (long long)npi, 256);
  for (int s = 0; s < n_here && !rc; ++s) {
    // map_in = calloc; then the components in the caller's (= the reference's) order
    rc = cudaMemsetAsync(d_acc, 0, npi * sizeof(float), c->stream) != cudaSuccess;
    for (int k = 0; k < n_comp && !rc; ++k) {
      const float *src;
      if (comp_host[k]) {
        rc = cudaMemcpyAsync(d_comp, comp_host[k] + (size_t)s * npi, npi * sizeof(float), cudaMemcpyHostToDevice, c->stream) != cudaSuccess;
        src = d_comp;
      } else {
        src = own + (size_t)s * npi;  // NULL entry: the device-resident cosmological signal
      }
      const bool scaled = scale && scale[k] != 1.0;
      add_maps_kernel<<<nb, 256, 0, c->stream>>>(d_acc, src, scale ? scale[k] : 1.0, scaled ? 1 : 0, (long long)npi);
      c->launches++;
      // d_comp is reused by the next component: pageable host memory makes the copy synchronous, pinned memory needs this
      rc = rc || cudaStreamSynchronize(c->stream) != cudaSuccess;
    }
    udgrade_kernel<<<blocks_for(c, (long long)npo, 256), 256, 0, c->stream>>>(d_acc, (long long)d.nside, d_out, nside_out, 0);
    c->launches++;
    rc = rc || cudaMemcpyAsync(out_host + (size_t)s * npo, d_out, npo * sizeof(float), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess;
    rc = rc || cudaStreamSynchronize(c->stream) != cudaSuccess;
  }
  cudaFree(d_acc); cudaFree(d_comp); cudaFree(d_out);
  if (rc) gh_set_error("gh_cuda_jt_merge_maps: %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}

extern "C" int gh_cuda_nest_ring(gh_cuda_ctx *c, long nside, const long long *pix_in, long long *pix_out, long long n, int to_ring)
{
  GH_CTX_J(c);
  GH_REQ_J(pix_in && pix_out && n >= 0 && pow2(nside) && nside <= 8192, "gh_cuda_nest_ring: bad arguments");
  if (n == 0) return 0;
  long long *d_in = nullptr, *d_out = nullptr;
  GH_CUDA_OK(cudaMalloc(&d_in, n * sizeof(long long)));
  if (cudaMalloc(&d_out, n * sizeof(long long)) != cudaSuccess) { cudaFree(d_in); gh_set_error("gh_cuda_nest_ring: out of device memory"); return 1; }
  int rc = cudaMemcpyAsync(d_in, pix_in, n * sizeof(long long), cudaMemcpyHostToDevice, c->stream) != cudaSuccess;
  nestring_kernel<<<blocks_for(c, n, 128), 128, 0, c->stream>>>(nside, d_in, d_out, n, to_ring);
  c->launches++;
  rc = rc || cudaMemcpyAsync(pix_out, d_out, n * sizeof(long long), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess;
  rc = rc || cudaStreamSynchronize(c->stream) != cudaSuccess;
  cudaFree(d_in); cudaFree(d_out);
  if (rc) gh_set_error("gh_cuda_nest_ring: %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}
