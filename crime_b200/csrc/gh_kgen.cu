// k-space Gaussian realisation of delta(k) and the velocity potential, hand-written for sm_100a.
//
// Replaces create_density_and_velpot_fourier (reference src/fourier.c:234-305), pk_linear0
// (src/cosmo.c:153-170) and rng_delta_gauss (src/common.c:154-164).  The reference draws from one
// MT19937 per OpenMP thread, so its realisation depends on the thread and rank count; here every mode
// draws from Philox4x32-10 keyed on (seed, GLOBAL mode index kx + nh*(ky + n*kz)) -- the reference's own
// single-process index (src/fourier.c:278) -- so the field is identical for any number of GPUs.
//   counter = (index lo, index hi, 0, 0), key = (seed, 'GetH');
//   u1 = (out[0] >> 8) * 2^-24  -> phase = 2 pi u1;   u2 = (out[1] >> 8) * 2^-24 -> |delta| = sqrt(-sigma2 ln(1-u2))
// P(k): bin index and interpolation in double from shared memory (a float log10 would flip bins at
// table nodes, where the reference's interpolant is discontinuous); Rayleigh/phase maths in float.
// Layout written: [kz][ky_local][kx], ky_local in this rank's ky slab -- ready for a local z transform.
// Write-only, 16 B/mode (two complex-float fields): each thread produces two adjacent modes and stores
// one float4 per field.
#include "gh_internal.cuh"

namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t &o0,
                                              uint32_t &o1)
{
  uint32_t c2 = 0u, c3 = 0u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o0 = c0; o1 = c1;
}

__device__ __forceinline__ double signed_k(int i, int n, double dk) { return (2 * i <= n) ? i * dk : -(n - i) * dk; }

// src/cosmo.c:153-170
__device__ __forceinline__ double pk_linear0(const GhDev &d, const double *s_logk, const double *s_pk, double lgk)
{
  const int ik = (int)((lgk - d.logkmin) * d.idlogk);
  if (ik < 0) return s_pk[0] * pow(10.0, d.n_scal * (lgk - d.logkmin));
  if (ik < d.numk) {
    const double hi = (ik + 1 < d.numk) ? s_pk[ik + 1] : s_pk[ik];
    return s_pk[ik] + (lgk - s_logk[ik]) * (hi - s_pk[ik]) * d.idlogk;
  }
  return s_pk[d.numk - 1] * pow(10.0, -3.0 * (lgk - d.logkmax));
}

__device__ __forceinline__ void one_mode(const GhDev &d, const double *s_logk, const double *s_pk, long long local,
                                         float2 &dk_out, float2 &vk_out)
{
  // local index -> (kz, ky_local, kx) of the [kz][ky_local][kx] slab
  const int kx = (int)(local % d.nh);
  const long long t = local / d.nh;
  const int kyl = (int)(t % d.nky_here), kz = (int)(t / d.nky_here);
  const int ky = d.ky0 + kyl;
  const double fx = signed_k(kx, d.n, d.dk), fy = signed_k(ky, d.n, d.dk), fz = signed_k(kz, d.n, d.dk);
  const double k2 = fx * fx + fy * fy + fz * fz;
  if (k2 <= 0.0) {  // src/fourier.c:287-290
    dk_out = make_float2(0.f, 0.f);
    vk_out = make_float2(0.f, 0.f);
    return;
  }
  const unsigned long long gidx = (unsigned long long)kx + (unsigned long long)d.nh * ((unsigned long long)ky + (unsigned long long)d.n * kz);
  uint32_t r0, r1;
  philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), d.seed, 0x47657448u, r0, r1);
  const float u1 = (float)(r0 >> 8) * 5.9604644775390625e-8f;  // 2^-24
  const float u2 = (float)(r1 >> 8) * 5.9604644775390625e-8f;
  const double lgk = 0.5 * log10(k2);
  double sigma2 = pk_linear0(d, s_logk, s_pk, lgk) * d.idk3;
  float s2f = (float)sigma2;
  const float k2f = (float)k2;
  if (d.do_smoothing) s2f *= expf(-(float)d.r2_smooth * k2f);  // src/fourier.c:294-295
  const float mod = sqrtf(-s2f * log1pf(-u2));                 // src/common.c:163
  float sn, cs;
  sincospif(2.0f * u1, &sn, &cs);                              // phase = 2 pi u1, src/common.c:161
  dk_out = make_float2(mod * cs, mod * sn);
  const float vf = (float)d.vfactor / k2f;                     // f0*H0/k^2, src/fourier.c:298
  vk_out = make_float2(dk_out.x * vf, dk_out.y * vf);
}

__global__ void __launch_bounds__(256) kgen_kernel(GhDev d, float2 *__restrict__ dens_k, float2 *__restrict__ vpot_k,
                                                   long long npairs)
{
  extern __shared__ double s_tab[];
  double *s_logk = s_tab, *s_pk = s_tab + d.numk;
  for (int i = threadIdx.x; i < d.numk; i += blockDim.x) {
    s_logk[i] = d.logkarr[i];
    s_pk[i] = d.pkarr[i];
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x; pair < npairs; pair += stride) {
    float2 a0, v0, a1, v1;
    one_mode(d, s_logk, s_pk, 2 * pair, a0, v0);
    one_mode(d, s_logk, s_pk, 2 * pair + 1, a1, v1);
    reinterpret_cast<float4 *>(dens_k)[pair] = make_float4(a0.x, a0.y, a1.x, a1.y);
    reinterpret_cast<float4 *>(vpot_k)[pair] = make_float4(v0.x, v0.y, v1.x, v1.y);
  }
}

}  // namespace

int gh_launch_kgen(gh_cuda_ctx *c)
{
  const GhDev &d = c->d;
  const long long nmodes = (long long)d.n * d.nky_here * d.nh;  // even: n is even
  const long long npairs = nmodes / 2;
  const size_t smem = 2 * sizeof(double) * (size_t)d.numk;
  if (smem > 200 * 1024) {
    gh_set_error("P(k) table with %d rows does not fit in shared memory", d.numk);
    return 1;
  }
  GH_CUDA_OK(cudaFuncSetAttribute(kgen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = (npairs + 255) / 256;
  const long long cap = (long long)c->n_sm * 8;  // grid-stride over a whole number of waves
  if (blocks > cap) blocks = cap;
  kgen_kernel<<<(unsigned)blocks, 256, smem, c->stream>>>(d, c->gridA, c->gridB, npairs);
  GH_LAUNCH_CHECK(c);
  return 0;
}
