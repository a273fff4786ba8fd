// Line transforms of any length for the 3-D c2r FFT: the arithmetic shared by the generic kernels of gh_fft.cu and the
// host-side test harness (tests/native/fft_generic_host.cpp), which runs the very same code on the CPU against numpy.
//
// The reference hands its grids to FFTW (src/fourier.c:78-99), which takes any n_grid; the tuned sm_100a kernels of
// gh_fft.cu are instantiated for powers of two only.  Every other even n_grid goes through the functions below: a
// mixed-radix Stockham autosort transform (exponent sign +, unnormalised, like FFTW's backward transform) over a tile of
// W adjacent lines held in shared memory, ping-ponging between two buffers so that a pass never reads what it writes
// and the result comes out in natural order without a digit reversal.  A thread owns one output element of a pass and
// gathers its r inputs, so any radix works with the same code (direct r-point DFT from the length-n twiddle table):
// O(n sum r_i) per line -- fast for n = 2^a 3^b 5^c 7^d, correct for any factorisation.
//
// Every function takes (tid, nthreads) and touches only elements it owns: a CTA calls them with threadIdx.x / blockDim.x
// and a __syncthreads() in between; the host harness loops tid serially per phase, which is equivalent because no phase
// reads an element that the same phase writes.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define GFFT_HD __host__ __device__ __forceinline__
#else
#define GFFT_HD static inline
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

#define GFFT_MAX_FACT 16

struct GfftPlan {
  int n;                    // transform length
  int nfact;                // number of passes
  int fact[GFFT_MAX_FACT];  // radices in pass order, product = n
};

// Radix plan: 4s first (fewest passes), then 2, 3, 5, 7, ... any remaining prime.  Returns 0 when n has more than
// GFFT_MAX_FACT factors (cannot happen below 2^20) or n < 1.
static inline int gfft_make_plan(int n, GfftPlan *p)
{
  if (n < 1) return 0;
  p->n = n;
  p->nfact = 0;
  int m = n;
  while (m % 4 == 0) { if (p->nfact == GFFT_MAX_FACT) return 0; p->fact[p->nfact++] = 4; m /= 4; }
  for (int f = 2; m > 1; f += (f == 2) ? 1 : 2) {
    if ((long long)f * f > m) f = m;  // m is prime
    while (m % f == 0) { if (p->nfact == GFFT_MAX_FACT) return 0; p->fact[p->nfact++] = f; m /= f; }
  }
  return 1;
}

// One Stockham pass of radix r over a [n][pitch] tile (line w of position pos at pos * pitch + w, w < W):
// the tile holds s interleaved transforms of length n_cur (s = product of the radices already applied, s * n_cur = n).
// Output position o = q + s (r p + k), q < s, k < r, p < n_cur / r:
//   y[o] = exp(2 pi i p k / n_cur) * sum_j x[q + s (p + j n_cur / r)] exp(2 pi i j k / r)
// tw[j * tws] = exp(2 pi i j / n).
GFFT_HD void gfft_pass(const float2 *x, float2 *y, const float2 *tw, int tws, int n, int W, int pitch, int r, int n_cur, int s,
                       int tid, int nthreads)
{
  const int m = n_cur / r;
  const int step_r = (n / r) * tws;  // exp(2 pi i / r) = tw[step_r]
  const int ntw = n * tws;
  for (int item = tid; item < n * W; item += nthreads) {
    const int w = item % W, o = item / W;
    const int q = o % s, t = o / s;
    const int k = t % r, p = t / r;
    const float2 *xi = x + (q + s * p) * pitch + w;
    const int in_stride = s * m * pitch;
    const int ek = k * step_r;  // < ntw
    float ax = xi[0].x, ay = xi[0].y;
    int e = 0;
    for (int j = 1; j < r; ++j) {
      e += ek;
      if (e >= ntw) e -= ntw;
      const float2 a = xi[j * in_stride];
      const float2 c = tw[e];
      ax += a.x * c.x - a.y * c.y;
      ay += a.x * c.y + a.y * c.x;
    }
    const float2 c = tw[p * k * s * tws];  // p k < n_cur, so the index stays below n * tws
    y[o * pitch + w] = make_float2(ax * c.x - ay * c.y, ax * c.y + ay * c.x);
  }
}

// The same pass with one thread per butterfly instead of one per output element, for the radices worth a register array
// (2, 3, 4, 5, 7): the R inputs are read once, the R outputs leave together.  Bit-for-bit the arithmetic of gfft_pass for
// R = 3, 5, 7 (same products, same summation order); R = 2 and 4 use the exact +-1 / +-i butterflies instead of table
// look-ups of those values.  fast = 0 in the launch geometry keeps every pass on gfft_pass.
template <int R>
GFFT_HD void gfft_pass_small(const float2 *x, float2 *y, const float2 *tw, int tws, int n, int W, int pitch, int n_cur, int s, int tid,
                             int nthreads)
{
  const int m = n_cur / R;
  const int step_r = (n / R) * tws;
  const int in_stride = s * m * pitch, out_stride = s * pitch;
  for (int item = tid; item < (n / R) * W; item += nthreads) {
    const int w = item % W, b = item / W;  // butterfly b = q + s p
    const int q = b % s, p = b / s;
    const float2 *xi = x + b * pitch + w;
    float2 a[R], o[R];
#pragma unroll
    for (int j = 0; j < R; ++j) a[j] = xi[j * in_stride];
    if (R == 2) {
      o[0] = make_float2(a[0].x + a[1].x, a[0].y + a[1].y);
      o[1] = make_float2(a[0].x - a[1].x, a[0].y - a[1].y);
    } else if (R == 4) {
      const float2 e = make_float2(a[0].x + a[2].x, a[0].y + a[2].y), f = make_float2(a[0].x - a[2].x, a[0].y - a[2].y);
      const float2 g = make_float2(a[1].x + a[3].x, a[1].y + a[3].y), d = make_float2(a[1].x - a[3].x, a[1].y - a[3].y);
      o[0] = make_float2(e.x + g.x, e.y + g.y);
      o[1] = make_float2(f.x - d.y, f.y + d.x);  // + i d
      o[2] = make_float2(e.x - g.x, e.y - g.y);
      o[3] = make_float2(f.x + d.y, f.y - d.x);  // - i d
    } else {
      float2 wr[R];  // exp(2 pi i j / R)
#pragma unroll
      for (int j = 1; j < R; ++j) wr[j] = tw[j * step_r];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        float ax = a[0].x, ay = a[0].y;
#pragma unroll
        for (int j = 1; j < R; ++j) {
          const float2 c = wr[(j * k) % R == 0 ? 1 : (j * k) % R];  // (j k) % R == 0 only for k == 0, handled below
          if (k == 0) {
            ax += a[j].x;
            ay += a[j].y;
          } else {
            ax += a[j].x * c.x - a[j].y * c.y;
            ay += a[j].x * c.y + a[j].y * c.x;
          }
        }
        o[k] = make_float2(ax, ay);
      }
    }
    float2 *yo = y + (q + s * R * p) * pitch + w;
    yo[0] = o[0];
    const int et = p * s * tws;  // exp(2 pi i p k / n_cur) = tw[k et], k p < n_cur
#pragma unroll
    for (int k = 1; k < R; ++k) {
      const float2 c = tw[k * et];
      yo[k * out_stride] = make_float2(o[k].x * c.x - o[k].y * c.y, o[k].x * c.y + o[k].y * c.x);
    }
  }
}

GFFT_HD void gfft_pass_any(int fast, const float2 *x, float2 *y, const float2 *tw, int tws, int n, int W, int pitch, int r, int n_cur,
                           int s, int tid, int nthreads)
{
  if (fast) {
    switch (r) {
      case 2: gfft_pass_small<2>(x, y, tw, tws, n, W, pitch, n_cur, s, tid, nthreads); return;
      case 3: gfft_pass_small<3>(x, y, tw, tws, n, W, pitch, n_cur, s, tid, nthreads); return;
      case 4: gfft_pass_small<4>(x, y, tw, tws, n, W, pitch, n_cur, s, tid, nthreads); return;
      case 5: gfft_pass_small<5>(x, y, tw, tws, n, W, pitch, n_cur, s, tid, nthreads); return;
      case 7: gfft_pass_small<7>(x, y, tw, tws, n, W, pitch, n_cur, s, tid, nthreads); return;
      default: break;
    }
  }
  gfft_pass(x, y, tw, tws, n, W, pitch, r, n_cur, s, tid, nthreads);
}

// Strided axis: line w of the tile is src[w], element pos of a line sits pos * stride further on.
GFFT_HD void gfft_load_strided(float2 *x, const float2 *src, long long stride, int n, int W, int nvalid, int tid, int nthreads)
{
  for (int item = tid; item < n * W; item += nthreads) {
    const int w = item % W, pos = item / W;
    x[pos * W + w] = (w < nvalid) ? src[(long long)pos * stride + w] : make_float2(0.f, 0.f);
  }
}

GFFT_HD void gfft_store_strided(const float2 *x, float2 *dst, long long stride, int n, int W, int nvalid, int tid, int nthreads)
{
  for (int item = tid; item < n * W; item += nthreads) {
    const int w = item % W, pos = item / W;
    if (w < nvalid) dst[(long long)pos * stride + w] = x[pos * W + w];
  }
}

// x axis, half-complex -> real through a half-length complex transform (n even, H = n / 2).  Row w of the tile is
// rows + w * row_stride, H + 1 modes X[0..H].  With A = X[k], B = conj X[H - k], E = A + B, T = w^k (A - B), w = exp(2 pi i / n):
//   Z[k] = E + i T, k < H;  the length-H transform z of Z carries the real row as z[m] = x[2m] + i x[2m + 1].
// Im X[0] and Im X[H] are not part of a half-complex spectrum and are never read (as FFTW's c2r, src/fourier.c:78-99).
GFFT_HD void gfft_rows_stage(float2 *x, const float2 *rows, long long row_stride, const float2 *tw, int H, int W, int pitch,
                             int nvalid, int tid, int nthreads)
{
  for (int item = tid; item < H * W; item += nthreads) {
    const int k = item % H, w = item / H;
    float2 z = make_float2(0.f, 0.f);
    if (w < nvalid) {
      const float2 *r = rows + (long long)w * row_stride;
      const float2 a = r[k];
      float2 b = r[H - k];
      if (k == 0) {
        z = make_float2(a.x + b.x, a.x - b.x);
      } else {
        b.y = -b.y;
        const float ex = a.x + b.x, ey = a.y + b.y, dx = a.x - b.x, dy = a.y - b.y;
        const float2 c = tw[k];
        const float tx = c.x * dx - c.y * dy, ty = c.x * dy + c.y * dx;
        z = make_float2(ex - ty, ey + tx);
      }
    }
    x[k * pitch + w] = z;
  }
}

// z[m] * norm -> the first H complex slots of each row (= the n real cells; the padding element stays as it is)
GFFT_HD void gfft_rows_gather(const float2 *x, float2 *rows, long long row_stride, float norm, int H, int W, int pitch, int nvalid,
                              int tid, int nthreads)
{
  for (int item = tid; item < H * W; item += nthreads) {
    const int m = item % H, w = item / H;
    if (w < nvalid) {
      const float2 z = x[m * pitch + w];
      rows[(long long)w * row_stride + m] = make_float2(z.x * norm, z.y * norm);
    }
  }
}

// ---- whole CTAs, phase by phase -------------------------------------------------------------------------------
// A CTA's work is a sequence of phases separated by barriers: phase 0 stages the tile, phases 1..nfact are the Stockham
// passes (buffer (phase - 1) & 1 -> buffer phase & 1), phase nfact + 1 writes the result back in place.  The kernels of
// gh_fft.cu are `for (phase...) { gfft_*_cta_phase(...); __syncthreads(); }`; the host harness runs the same functions
// block by block, phase by phase, thread by thread.
struct GfftGeom {
  int lines_per_group, tiles_per_group;
  long long group_stride;  // between the groups (z planes of the y pass)
  long long stride;        // between consecutive elements of a line
};

// n_cur and s of pass f (0-based): s = product of the radices before it
GFFT_HD void gfft_pass_state(const GfftPlan &plan, int f, int &n_cur, int &s)
{
  n_cur = plan.n;
  s = 1;
  for (int i = 0; i < f; ++i) {
    n_cur /= plan.fact[i];
    s *= plan.fact[i];
  }
}

// strided axis (z: one flat group of n * nh columns; y: one group per z plane), W lines per CTA, pitch W
GFFT_HD void gfft_strided_cta_phase(int phase, float2 *sm, float2 *data, const float2 *tw, const GfftPlan &plan, int W,
                                    const GfftGeom &g, int fast, long long block, int tid, int nthreads)
{
  const int n = plan.n;
  float2 *const b0 = sm, *const b1 = sm + (size_t)n * W;  // buffer i holds the tile after pass i (i & 1)
  const long long grp = block / g.tiles_per_group;
  const int l0 = (int)(block - grp * g.tiles_per_group) * W;
  float2 *base = data + (grp * g.group_stride + l0);
  const int nvalid = (g.lines_per_group - l0 < W) ? g.lines_per_group - l0 : W;
  if (phase == 0) {
    gfft_load_strided(b0, base, g.stride, n, W, nvalid, tid, nthreads);
  } else if (phase <= plan.nfact) {
    int n_cur, s;
    gfft_pass_state(plan, phase - 1, n_cur, s);
    gfft_pass_any(fast, (phase & 1) ? b0 : b1, (phase & 1) ? b1 : b0, tw, 1, n, W, W, plan.fact[phase - 1], n_cur, s, tid, nthreads);
  } else {
    gfft_store_strided((plan.nfact & 1) ? b1 : b0, base, g.stride, n, W, nvalid, tid, nthreads);
  }
}

// x axis: plan is the half-length transform (H = n_grid / 2), tw the length-n_grid table (exp(2 pi i j / H) = tw[2 j]),
// rows of nh = H + 1 modes, W rows per CTA at pitch `pitch` >= W
GFFT_HD void gfft_rows_cta_phase(int phase, float2 *sm, float2 *data, const float2 *tw, const GfftPlan &plan, int W, int pitch,
                                 long long nrows, int nh, float norm, int fast, long long block, int tid, int nthreads)
{
  const int H = plan.n;
  float2 *const b0 = sm, *const b1 = sm + (size_t)H * pitch;
  const long long row0 = block * W;
  const int nvalid = (nrows - row0 < W) ? (int)(nrows - row0) : W;
  float2 *rows = data + row0 * nh;
  if (phase == 0) {
    gfft_rows_stage(b0, rows, nh, tw, H, W, pitch, nvalid, tid, nthreads);
  } else if (phase <= plan.nfact) {
    int n_cur, s;
    gfft_pass_state(plan, phase - 1, n_cur, s);
    gfft_pass_any(fast, (phase & 1) ? b0 : b1, (phase & 1) ? b1 : b0, tw, 2, H, W, pitch, plan.fact[phase - 1], n_cur, s, tid, nthreads);
  } else {
    gfft_rows_gather((plan.nfact & 1) ? b1 : b0, rows, nh, norm, H, W, pitch, nvalid, tid, nthreads);
  }
}

// Lines per tile of a length-len transform: two buffers of len * W modes; up to 96 KB (two CTAs per SM) while that leaves
// rows of >= 32 bytes, else up to 200 KB.  0: a line does not fit twice.
static inline int gfft_tile_width(int len)
{
  const int w96 = (96 * 1024) / (16 * len);
  if (w96 >= 4) return w96 > 16 ? 16 : w96;
  const int w200 = (200 * 1024) / (16 * len);
  return w200 > 4 ? 4 : w200;
}

// Launch geometry of one field on one rank, shared by the launcher and the host harness
struct GfftLaunch {
  GfftPlan pn, ph;        // length n_grid and n_grid / 2
  int W, WR, pitch;       // strided tile width; rows per tile of the x pass and their pitch (odd: staging and gathering walk
                          // along a line with the row fixed)
  size_t smem_s, smem_r;  // dynamic shared memory of the two kernels
  GfftGeom gz, gy;
  long long blocks_z, blocks_y, blocks_x, nrows;
};

static inline int gfft_make_launch(int n, int nz, GfftLaunch *L)
{
  if (n < 2 || (n & 1)) return 0;
  const int nh = n / 2 + 1, H = n / 2;
  if (!gfft_make_plan(n, &L->pn) || !gfft_make_plan(H, &L->ph)) return 0;
  L->W = gfft_tile_width(n);
  L->WR = gfft_tile_width(H);
  while (L->WR > 1 && (size_t)2 * H * (L->WR | 1) * sizeof(float2) > 200 * 1024) --L->WR;
  L->pitch = L->WR | 1;
  L->smem_s = (size_t)2 * n * L->W * sizeof(float2);
  L->smem_r = (size_t)2 * H * L->pitch * sizeof(float2);
  if (L->W < 1 || L->WR < 1 || L->smem_s > 200 * 1024 || L->smem_r > 200 * 1024) return 0;
  // (1) z axis: all n * nh columns as one flat group
  L->gz.lines_per_group = n * nh;
  L->gz.tiles_per_group = (L->gz.lines_per_group + L->W - 1) / L->W;
  L->gz.group_stride = 0;
  L->gz.stride = (long long)n * nh;
  L->blocks_z = L->gz.tiles_per_group;
  // (2) y axis, per z plane
  L->gy.lines_per_group = nh;
  L->gy.tiles_per_group = (nh + L->W - 1) / L->W;
  L->gy.group_stride = (long long)n * nh;
  L->gy.stride = nh;
  L->blocks_y = (long long)L->gy.tiles_per_group * nz;
  // (3) x axis
  L->nrows = (long long)nz * n;
  L->blocks_x = (L->nrows + L->WR - 1) / L->WR;
  return 1;
}
