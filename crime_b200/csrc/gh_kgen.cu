// k-space Gaussian realisation of delta(k) and the velocity potential, hand-written for sm_100a.
//
// Replaces create_density_and_velpot_fourier (reference src/fourier.c:234-305), pk_linear0
// (src/cosmo.c:153-170) and rng_delta_gauss (src/common.c:154-164).  The reference draws from one
// MT19937 per OpenMP thread, so its realisation depends on the thread and rank count; here every mode
// draws from Philox4x32-10 keyed on (seed, GLOBAL mode index g = kx + nh*(ky + n*kz)) -- the reference's own
// single-process index (src/fourier.c:278) -- so the field is identical for any number of GPUs.
//   counter = (g>>1 lo, g>>1 hi, 0, 0), key = (seed, 'GetH'); the even mode of a pair takes words 0,1 and the
//   odd one words 2,3:  u1 = (w_a >> 8) * 2^-24 -> phase = 2 pi u1;  u2 = w_b * 2^-32 (all 32 bits, as
//   gsl_rng_uniform, so the Rayleigh tail reaches sqrt(ln 2^32) = 4.7 sigma) -> |delta| = sqrt(-sigma2 ln(1-u2)).
//   One thread owns one pair, so one Philox block serves two modes and the pair leaves as one 16-byte store per field.
// P(k): the reference indexes its table with (int)((0.5 log10 k^2 - logkmin) * idlogk) in double.  Here the
// bin comes from a float log2; only when that lands within 2e-3 of a bin boundary (where the reference's
// interpolant is discontinuous, so the bin matters) is the double expression evaluated.
// Layout written: [kz][ky_local][kx], ky_local in this rank's ky slab -- ready for a local z transform.
// Write-only, 16 B/mode (two complex-float fields), coalesced 16-byte stores (a pair of adjacent modes per thread).
#include "gh_internal.cuh"
#include "gh_philox.cuh"

namespace {

__device__ __forceinline__ int signed_idx(int i, int n) { return (2 * i <= n) ? i : i - n; }

// src/cosmo.c:153-170 with the bin chosen in float (double only next to a bin boundary)
__device__ __forceinline__ float pk_lookup(const GhDev &d, const double *s_logk, const double *s_pk, const float *s_logk_f,
                                           const float *s_pk_f, float k2f, double k2)
{
  // lgk = 0.5 log10(k2)
  const float lgk = 0.15051499783199059761f * __log2f(k2f);  // 0.5*log10(2)*log2
  const float xf = (lgk - (float)d.logkmin) * (float)d.idlogk;
  int ik = (int)floorf(xf);
  const float fr = xf - (float)ik;
  if (fr < 2e-3f || fr > 1.0f - 2e-3f || ik < 1 || ik >= d.numk - 2) {
    // rare: decide the bin (and the two extrapolation branches) exactly as the reference does
    const double lg = 0.5 * log10(k2);
    const int ikd = (int)((lg - d.logkmin) * d.idlogk);
    if (ikd < 0) return (float)(s_pk[0] * pow(10.0, d.n_scal * (lg - d.logkmin)));
    if (ikd >= d.numk) return (float)(s_pk[d.numk - 1] * pow(10.0, -3.0 * (lg - d.logkmax)));
    const double hi = (ikd + 1 < d.numk) ? s_pk[ikd + 1] : s_pk[ikd];
    return (float)(s_pk[ikd] + (lg - s_logk[ikd]) * (hi - s_pk[ikd]) * d.idlogk);
  }
  const float a = s_pk_f[ik], b = s_pk_f[ik + 1];
  return fmaf((lgk - s_logk_f[ik]) * (float)d.idlogk, b - a, a);
}

__device__ __forceinline__ void one_mode(const GhDev &d, const double *s_logk, const double *s_pk, const float *s_logk_f,
                                         const float *s_pk_f, int kx, int ky, int kz, uint32_t wa, uint32_t wb,
                                         float2 &dk_out, float2 &vk_out)
{
  const int ix = signed_idx(kx, d.n), iy = signed_idx(ky, d.n), iz = signed_idx(kz, d.n);
  const int m2 = ix * ix + iy * iy + iz * iz;  // <= 3 (n/2)^2 < 2^24: exact in float
  if (m2 == 0) {  // src/fourier.c:287-290
    dk_out = make_float2(0.f, 0.f);
    vk_out = make_float2(0.f, 0.f);
    return;
  }
  const double k2 = d.dk * d.dk * (double)m2;
  const float k2f = (float)k2;
  float s2f = pk_lookup(d, s_logk, s_pk, s_logk_f, s_pk_f, k2f, k2) * (float)d.idk3;
  if (d.do_smoothing) s2f *= __expf(-(float)d.r2_smooth * k2f);   // src/fourier.c:294-295
  // -ln(1 - u2), u2 = wb / 2^32 (src/common.c:163), to ~2e-7 relative over the whole range:
  //   u2 <  1/2: 2 atanh(t), t = u2 / (2 - u2) <= 1/3, odd series to t^13 (next term < 1.2e-8 relative);
  //   u2 >= 1/2: -ln(v) with v = (2^32 - wb) / 2^32 exact to 24 bits where the tail needs it, |ln v| >= 0.69
  float nl;
  if (wb < 0x80000000u) {
    const float u2 = (float)wb * 2.3283064365386963e-10f;  // 2^-32
    const float t = __fdividef(u2, 2.0f - u2), t2 = t * t;
    float p = fmaf(t2, 0.15384615f, 0.18181818f);          // 2/13, 2/11, 2/9, 2/7, 2/5, 2/3, 2
    p = fmaf(t2, p, 0.22222222f);
    p = fmaf(t2, p, 0.28571429f);
    p = fmaf(t2, p, 0.4f);
    p = fmaf(t2, p, 0.66666667f);
    p = fmaf(t2, p, 2.0f);
    nl = t * p;
  } else {
    const float v = (float)(0u - wb) * 2.3283064365386963e-10f;  // 2^32 - wb <= 2^31: exact in float up to rounding to 24 bits
    nl = -0.69314718055994530942f * __log2f(v);
  }
  float mod;
  {
    const float arg = s2f * nl;                                   // src/common.c:163: sqrt(-sigma2 ln(1 - u2))
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mod) : "f"(arg));
  }
  // phase = 2 pi u1 (src/common.c:161) through the SFU: the argument is folded into [-pi, pi), where sin/cos.approx
  // are good to 2^-21 absolute (on a unit phasor: 5e-7 of the modulus)
  const float u1c = (float)(wa >> 8) * 5.9604644775390625e-8f - 0.5f;  // u1 - 1/2, exact
  const float ang = 6.283185307179586f * u1c;
  const float cs = -__cosf(ang), sn = -__sinf(ang);                    // cos(x + pi) = -cos x
  dk_out = make_float2(mod * cs, mod * sn);
  const float vf = __fdividef((float)d.vfactor, k2f);              // f0*H0/k^2, src/fourier.c:298
  vk_out = make_float2(dk_out.x * vf, dk_out.y * vf);
}

__global__ void __launch_bounds__(256) kgen_kernel(GhDev d, float2 *__restrict__ dens_k, float2 *__restrict__ vpot_k)
{
  extern __shared__ double s_tab[];
  double *s_logk = s_tab, *s_pk = s_tab + d.numk;
  float *s_logk_f = reinterpret_cast<float *>(s_tab + 2 * d.numk), *s_pk_f = s_logk_f + d.numk;
  for (int i = threadIdx.x; i < d.numk; i += blockDim.x) {
    const double a = d.logkarr[i], b = d.pkarr[i];
    s_logk[i] = a; s_pk[i] = b;
    s_logk_f[i] = (float)a; s_pk_f[i] = (float)b;
  }
  __syncthreads();
  const int kz = blockIdx.y;
  const unsigned plane_len = (unsigned)d.nky_here * (unsigned)d.nh;                       // modes of this plane here
  const unsigned long long gb = (unsigned long long)d.nh * ((unsigned long long)d.ky0 + (unsigned long long)d.n * kz);
  const unsigned long long q0 = gb >> 1;
  const unsigned npairs = (unsigned)(((gb + plane_len + 1) >> 1) - q0);
  const float inv_nh = 1.0f / (float)d.nh;
  float2 *dplane = dens_k + (size_t)kz * plane_len, *vplane = vpot_k + (size_t)kz * plane_len;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < npairs; t += gridDim.x * blockDim.x) {
    const unsigned long long q = q0 + t;
    uint32_t w0, w1, w2, w3;
    philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), d.seed, 0x47657448u, w0, w1, w2, w3);
    float2 a[2], v[2];
    bool in[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const long long off = (long long)(2 * q + s) - (long long)gb;
      in[s] = off >= 0 && off < (long long)plane_len;
      if (!in[s]) continue;
      const unsigned loc = (unsigned)off;
      int kyl = (int)((float)loc * inv_nh);
      int kx = (int)loc - kyl * d.nh;
      if (kx < 0) { kyl--; kx += d.nh; }
      else if (kx >= d.nh) { kyl++; kx -= d.nh; }
      one_mode(d, s_logk, s_pk, s_logk_f, s_pk_f, kx, d.ky0 + kyl, kz, s ? w2 : w0, s ? w3 : w1, a[s], v[s]);
    }
    const long long off0 = (long long)(2 * q) - (long long)gb;
    if (in[0] && in[1] && !(gb & 1)) {
      // the usual case (n and the ky slab are even, so every plane starts on an even global index): the pair is
      // 16-byte aligned in both fields
      *reinterpret_cast<float4 *>(dplane + off0) = make_float4(a[0].x, a[0].y, a[1].x, a[1].y);
      *reinterpret_cast<float4 *>(vplane + off0) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    } else {
      if (in[0]) { dplane[off0] = a[0]; vplane[off0] = v[0]; }
      if (in[1]) { dplane[off0 + 1] = a[1]; vplane[off0 + 1] = v[1]; }
    }
  }
}

}  // namespace

int gh_launch_kgen(gh_cuda_ctx *c)
{
  const GhDev &d = c->d;
  const size_t smem = (2 * sizeof(double) + 2 * sizeof(float)) * (size_t)d.numk;
  if (smem > 200 * 1024) {
    gh_set_error("P(k) table with %d rows does not fit in shared memory", d.numk);
    return 1;
  }
  GH_CUDA_OK(cudaFuncSetAttribute(kgen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned pairs = ((unsigned)d.nky_here * (unsigned)d.nh + 3) / 2;
  unsigned bx = (pairs + 255) / 256;
  if (bx > 8) bx = 8;  // few, fat CTAs per plane amortise the 13 KB table load; the n planes supply the parallelism
  dim3 grid(bx, d.n);
  kgen_kernel<<<grid, 256, smem, c->stream>>>(d, c->gridA, c->gridB);
  GH_LAUNCH_CHECK(c);
  return 0;
}
