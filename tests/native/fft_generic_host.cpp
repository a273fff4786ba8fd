// Host build of crime_b200/csrc/gh_fft_generic.cuh (test infrastructure, not part of the product): runs the CTA phase
// functions of the general-length FFT kernels block by block, phase by phase, thread by thread, with the launch geometry
// the launcher computes, so that tests/test_fft_generic_cpu.py can compare them with numpy without a GPU.
// The shared-memory tile and the field are surrounded by NaN guard zones: an out-of-range read poisons the result, an
// out-of-range write is reported.  `reverse` runs the threads of every phase in the opposite order: phases must not
// depend on the order (that is what the __syncthreads() between them relies on).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../crime_b200/csrc/gh_fft_generic.cuh"

namespace {

const size_t GUARD = 64;  // float2 elements either side

struct Guarded {
  std::vector<float2> v;
  size_t n;
  explicit Guarded(size_t n_) : v(n_ + 2 * GUARD), n(n_) { fill(); }
  void fill()
  {
    for (auto &e : v) e = make_float2(NAN, NAN);
  }
  float2 *data() { return v.data() + GUARD; }
  bool intact() const
  {
    for (size_t i = 0; i < GUARD; ++i)
      if (!isnan(v[i].x) || !isnan(v[GUARD + n + i].x)) return false;
    return true;
  }
};

std::vector<float2> twiddles(int n)
{
  std::vector<float2> tw(n);
  for (int j = 0; j < n; ++j) {
    const double a = 2.0 * M_PI * (double)j / (double)n;  // as gh_cuda_create fills ctx->twiddle
    tw[j] = make_float2((float)cos(a), (float)sin(a));
  }
  return tw;
}

template <class F> void for_threads(int nthreads, int reverse, F f)
{
  if (!reverse)
    for (int t = 0; t < nthreads; ++t) f(t);
  else
    for (int t = nthreads - 1; t >= 0; --t) f(t);
}

int g_fast = 0;  // gfft_host_set_fast

int run_strided(float2 *data, const float2 *tw, const GfftPlan &plan, int W, const GfftGeom &g, long long blocks, size_t smem,
                int nthreads, int reverse)
{
  Guarded sm(smem / sizeof(float2));
  for (long long b = 0; b < blocks; ++b) {
    sm.fill();
    for (int phase = 0; phase < plan.nfact + 2; ++phase)
      for_threads(nthreads, reverse, [&](int t) { gfft_strided_cta_phase(phase, sm.data(), data, tw, plan, W, g, g_fast, b, t, nthreads); });
    if (!sm.intact()) return 1;
  }
  return 0;
}

int run_rows(float2 *data, const float2 *tw, const GfftPlan &plan, int W, int pitch, long long nrows, int nh, float norm,
             long long blocks, size_t smem, int nthreads, int reverse)
{
  Guarded sm(smem / sizeof(float2));
  for (long long b = 0; b < blocks; ++b) {
    sm.fill();
    for (int phase = 0; phase < plan.nfact + 2; ++phase)
      for_threads(nthreads, reverse,
                  [&](int t) { gfft_rows_cta_phase(phase, sm.data(), data, tw, plan, W, pitch, nrows, nh, norm, g_fast, b, t, nthreads); });
    if (!sm.intact()) return 1;
  }
  return 0;
}

}  // namespace

// 1: one thread per butterfly for the radices 2, 3, 4, 5, 7 (gfft_pass_small); 0: one thread per output element everywhere
extern "C" void gfft_host_set_fast(int fast) { g_fast = fast; }

// The radix plan of a length (for the tests to inspect): returns nfact, fills fact[16]
extern "C" int gfft_host_plan(int n, int *fact)
{
  GfftPlan p;
  if (!gfft_make_plan(n, &p)) return -1;
  for (int i = 0; i < p.nfact; ++i) fact[i] = p.fact[i];
  return p.nfact;
}

// Launch geometry for (n, nz): W, WR, pitch, smem_s, smem_r, blocks_z, blocks_y, blocks_x
extern "C" int gfft_host_launch(int n, int nz, long long *out)
{
  GfftLaunch L;
  if (!gfft_make_launch(n, nz, &L)) return 1;
  out[0] = L.W; out[1] = L.WR; out[2] = L.pitch; out[3] = (long long)L.smem_s; out[4] = (long long)L.smem_r;
  out[5] = L.blocks_z; out[6] = L.blocks_y; out[7] = L.blocks_x;
  return 0;
}

// One whole field, in place: field = [n][n][n/2+1] complex-float (the padded real layout on return), exactly the three
// launches of fft_field_generic.  Returns 0, or 1 launch geometry refused, 2 shared-memory guard hit, 3 field guard hit.
extern "C" int gfft_host_field(float *field, int n, double norm, int nthreads, int reverse)
{
  GfftLaunch L;
  if (!gfft_make_launch(n, n, &L)) return 1;
  const int nh = n / 2 + 1;
  const size_t total = (size_t)n * n * nh;
  Guarded buf(total);
  memcpy(buf.data(), field, total * sizeof(float2));
  const std::vector<float2> tw = twiddles(n);
  if (run_strided(buf.data(), tw.data(), L.pn, L.W, L.gz, L.blocks_z, L.smem_s, nthreads, reverse)) return 2;
  if (run_strided(buf.data(), tw.data(), L.pn, L.W, L.gy, L.blocks_y, L.smem_s, nthreads, reverse)) return 2;
  if (run_rows(buf.data(), tw.data(), L.ph, L.WR, L.pitch, L.nrows, nh, (float)norm, L.blocks_x, L.smem_r, nthreads, reverse)) return 2;
  if (!buf.intact()) return 3;
  memcpy(field, buf.data(), total * sizeof(float2));
  return 0;
}

// One strided pass over data = [n][lines] complex-float (element pos of line l at pos * lines + l), with the tile width
// and shared-memory size the launcher would use for this n: the large lengths, where a whole cube is out of reach here.
extern "C" int gfft_host_strided_lines(float *data, int n, int lines, int nthreads, int reverse)
{
  GfftLaunch L;
  if (!gfft_make_launch(n, 2, &L)) return 1;
  GfftGeom g;
  g.lines_per_group = lines;
  g.tiles_per_group = (lines + L.W - 1) / L.W;
  g.group_stride = 0;
  g.stride = lines;
  const size_t total = (size_t)n * lines;
  Guarded buf(total);
  memcpy(buf.data(), data, total * sizeof(float2));
  const std::vector<float2> tw = twiddles(n);
  if (run_strided(buf.data(), tw.data(), L.pn, L.W, g, g.tiles_per_group, L.smem_s, nthreads, reverse)) return 2;
  if (!buf.intact()) return 3;
  memcpy(data, buf.data(), total * sizeof(float2));
  return 0;
}

// The x pass over nrows rows of n/2+1 modes each, in place
extern "C" int gfft_host_rows(float *data, int n, long long nrows, double norm, int nthreads, int reverse)
{
  GfftLaunch L;
  if (!gfft_make_launch(n, 2, &L)) return 1;
  const int nh = n / 2 + 1;
  const size_t total = (size_t)nrows * nh;
  Guarded buf(total);
  memcpy(buf.data(), data, total * sizeof(float2));
  const std::vector<float2> tw = twiddles(n);
  const long long blocks = (nrows + L.WR - 1) / L.WR;
  if (run_rows(buf.data(), tw.data(), L.ph, L.WR, L.pitch, nrows, nh, (float)norm, blocks, L.smem_r, nthreads, reverse)) return 2;
  if (!buf.intact()) return 3;
  memcpy(data, buf.data(), total * sizeof(float2));
  return 0;
}
