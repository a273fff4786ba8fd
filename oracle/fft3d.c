/* TEST INFRASTRUCTURE ONLY (oracle).  See fft3d.h.
 * Mixed-radix Stockham autosort FFT, vectorised over a block of FFT_B independent lines. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "fft3d.h"

#define FFT_B 16
#define MAX_FACT 64

typedef struct {
  int n, nfact, fact[MAX_FACT];
  double *twr, *twi; /* exp(+2 pi i j / n), j<n */
} fft_plan;

static void plan_make(fft_plan *p, int n)
{
  p->n = n;
  p->nfact = 0;
  int m = n;
  while (m % 4 == 0) { p->fact[p->nfact++] = 4; m /= 4; }
  while (m % 2 == 0) { p->fact[p->nfact++] = 2; m /= 2; }
  for (int f = 3; m > 1; f += 2)
    while (m % f == 0) { p->fact[p->nfact++] = f; m /= f; }
  p->twr = malloc(sizeof(double) * n);
  p->twi = malloc(sizeof(double) * n);
  for (int j = 0; j < n; j++) {
    double a = 2.0 * M_PI * (double)j / (double)n;
    p->twr[j] = cos(a);
    p->twi[j] = sin(a);
  }
}

static void plan_free(fft_plan *p) { free(p->twr); free(p->twi); }

/* One Stockham pass of radix R over FFT_B interleaved lines.
 * Work item i in [0,n/R): k = i mod p, outputs go to (i-k)*R + k + r*p. */
__attribute__((target_clones("avx2,fma", "default")))
static void stockham_pass(const fft_plan *pl, int R, int p, const double *restrict xr, const double *restrict xi,
                          double *restrict yr, double *restrict yi)
{
  const int n = pl->n, T = n / R;
  const int tstep = n / (p * R); /* twiddle index stride: exp(2 pi i r k/(pR)) = tw[r k tstep] */
  double ur[8][FFT_B], ui[8][FFT_B];
  for (int i = 0; i < T; i++) {
    int k = i % p, j = (i - k) * R + k;
    if (R <= 5) {
      for (int r = 0; r < R; r++) {
        const double *sr = xr + (size_t)(i + r * T) * FFT_B, *si = xi + (size_t)(i + r * T) * FFT_B;
        int ti = (int)(((long)r * k * tstep) % n);
        double wr = pl->twr[ti], wi = pl->twi[ti];
        for (int b = 0; b < FFT_B; b++) {
          ur[r][b] = sr[b] * wr - si[b] * wi;
          ui[r][b] = sr[b] * wi + si[b] * wr;
        }
      }
      if (R == 2) {
        double *o0r = yr + (size_t)j * FFT_B, *o0i = yi + (size_t)j * FFT_B;
        double *o1r = yr + (size_t)(j + p) * FFT_B, *o1i = yi + (size_t)(j + p) * FFT_B;
        for (int b = 0; b < FFT_B; b++) {
          o0r[b] = ur[0][b] + ur[1][b]; o0i[b] = ui[0][b] + ui[1][b];
          o1r[b] = ur[0][b] - ur[1][b]; o1i[b] = ui[0][b] - ui[1][b];
        }
      } else if (R == 4) {
        double *o0r = yr + (size_t)j * FFT_B, *o0i = yi + (size_t)j * FFT_B;
        double *o1r = o0r + (size_t)p * FFT_B, *o1i = o0i + (size_t)p * FFT_B;
        double *o2r = o1r + (size_t)p * FFT_B, *o2i = o1i + (size_t)p * FFT_B;
        double *o3r = o2r + (size_t)p * FFT_B, *o3i = o2i + (size_t)p * FFT_B;
        for (int b = 0; b < FFT_B; b++) {
          double ar = ur[0][b] + ur[2][b], ai = ui[0][b] + ui[2][b];
          double br = ur[0][b] - ur[2][b], bi = ui[0][b] - ui[2][b];
          double cr = ur[1][b] + ur[3][b], ci = ui[1][b] + ui[3][b];
          double dr = ur[1][b] - ur[3][b], di = ui[1][b] - ui[3][b];
          /* sign +: W4 = +i */
          o0r[b] = ar + cr; o0i[b] = ai + ci;
          o1r[b] = br - di; o1i[b] = bi + dr;
          o2r[b] = ar - cr; o2i[b] = ai - ci;
          o3r[b] = br + di; o3i[b] = bi - dr;
        }
      } else { /* R = 3 or 5: small dense DFT */
        for (int q = 0; q < R; q++) {
          double *orr = yr + (size_t)(j + q * p) * FFT_B, *oi = yi + (size_t)(j + q * p) * FFT_B;
          for (int b = 0; b < FFT_B; b++) { orr[b] = 0; oi[b] = 0; }
          for (int r = 0; r < R; r++) {
            int ti = (int)(((long)q * r * (n / R)) % n);
            double wr = pl->twr[ti], wi = pl->twi[ti];
            for (int b = 0; b < FFT_B; b++) {
              orr[b] += ur[r][b] * wr - ui[r][b] * wi;
              oi[b] += ur[r][b] * wi + ui[r][b] * wr;
            }
          }
        }
      }
    } else { /* generic prime radix: O(R^2) */
      for (int q = 0; q < R; q++) {
        double *orr = yr + (size_t)(j + q * p) * FFT_B, *oi = yi + (size_t)(j + q * p) * FFT_B;
        for (int b = 0; b < FFT_B; b++) { orr[b] = 0; oi[b] = 0; }
        for (int r = 0; r < R; r++) {
          const double *sr = xr + (size_t)(i + r * T) * FFT_B, *si = xi + (size_t)(i + r * T) * FFT_B;
          long ti = ((long)r * k * tstep + (long)q * r * (n / R)) % n;
          double wr = pl->twr[ti], wi = pl->twi[ti];
          for (int b = 0; b < FFT_B; b++) {
            orr[b] += sr[b] * wr - si[b] * wi;
            oi[b] += sr[b] * wi + si[b] * wr;
          }
        }
      }
    }
  }
}

/* FFT of FFT_B lines held as re[n][FFT_B], im[n][FFT_B]; result left in (re,im); (wr,wi) is scratch */
static void fft_block(const fft_plan *pl, double *re, double *im, double *wr, double *wi)
{
  double *ar = re, *ai = im, *br = wr, *bi = wi;
  int p = 1;
  for (int s = 0; s < pl->nfact; s++) {
    int R = pl->fact[s];
    stockham_pass(pl, R, p, ar, ai, br, bi);
    p *= R;
    double *t;
    t = ar; ar = br; br = t;
    t = ai; ai = bi; bi = t;
  }
  if (ar != re) {
    memcpy(re, ar, sizeof(double) * pl->n * FFT_B);
    memcpy(im, ai, sizeof(double) * pl->n * FFT_B);
  }
}

void oracle_fft1d(int n, double _Complex *x)
{
  fft_plan pl;
  plan_make(&pl, n);
  double *buf = calloc((size_t)4 * n * FFT_B, sizeof(double));
  double *re = buf, *im = buf + (size_t)n * FFT_B, *wr = im + (size_t)n * FFT_B, *wi = wr + (size_t)n * FFT_B;
  for (int j = 0; j < n; j++) { re[(size_t)j * FFT_B] = creal(x[j]); im[(size_t)j * FFT_B] = cimag(x[j]); }
  fft_block(&pl, re, im, wr, wi);
  for (int j = 0; j < n; j++) x[j] = re[(size_t)j * FFT_B] + I * im[(size_t)j * FFT_B];
  free(buf);
  plan_free(&pl);
}

/* complex FFT along an axis of length n and element stride `stride` for `nlines` lines whose first
 * elements are base + line*1 (lines are adjacent in memory: the fast index) */
static void fft_strided_lines(const fft_plan *pl, float _Complex *base, size_t stride, int nlines)
{
  const int n = pl->n;
  double *buf = malloc(sizeof(double) * 4 * (size_t)n * FFT_B);
  double *re = buf, *im = buf + (size_t)n * FFT_B, *wr = im + (size_t)n * FFT_B, *wi = wr + (size_t)n * FFT_B;
  for (int l0 = 0; l0 < nlines; l0 += FFT_B) {
    int nb = nlines - l0 < FFT_B ? nlines - l0 : FFT_B;
    for (int j = 0; j < n; j++) {
      const float _Complex *src = base + (size_t)j * stride + l0;
      for (int b = 0; b < nb; b++) { re[(size_t)j * FFT_B + b] = crealf(src[b]); im[(size_t)j * FFT_B + b] = cimagf(src[b]); }
      for (int b = nb; b < FFT_B; b++) { re[(size_t)j * FFT_B + b] = 0; im[(size_t)j * FFT_B + b] = 0; }
    }
    fft_block(pl, re, im, wr, wi);
    for (int j = 0; j < n; j++) {
      float _Complex *dst = base + (size_t)j * stride + l0;
      for (int b = 0; b < nb; b++) dst[b] = (float)re[(size_t)j * FFT_B + b] + I * (float)im[(size_t)j * FFT_B + b];
    }
  }
  free(buf);
}

/* half-complex -> real along contiguous rows: `nrows` rows of nh=n/2+1 complex, row stride nh complex,
 * output n reals at the start of each row (in place).  Im(DC), Im(Nyquist) are never read. */
static void c2r_rows(const fft_plan *pl, float _Complex *base, int nrows)
{
  const int n = pl->n, nh = n / 2 + 1;
  double *buf = malloc(sizeof(double) * 4 * (size_t)n * FFT_B);
  double *re = buf, *im = buf + (size_t)n * FFT_B, *wr = im + (size_t)n * FFT_B, *wi = wr + (size_t)n * FFT_B;
  for (int r0 = 0; r0 < nrows; r0 += FFT_B) {
    int nb = nrows - r0 < FFT_B ? nrows - r0 : FFT_B;
    for (int b = 0; b < FFT_B; b++) {
      if (b >= nb) { for (int j = 0; j < n; j++) { re[(size_t)j * FFT_B + b] = 0; im[(size_t)j * FFT_B + b] = 0; } continue; }
      const float _Complex *row = base + (size_t)(r0 + b) * nh;
      re[b] = crealf(row[0]); im[b] = 0;
      for (int j = 1; j < nh; j++) {
        double a = crealf(row[j]), c = cimagf(row[j]);
        if (2 * j == n) { re[(size_t)j * FFT_B + b] = a; im[(size_t)j * FFT_B + b] = 0; }
        else {
          re[(size_t)j * FFT_B + b] = a; im[(size_t)j * FFT_B + b] = c;
          re[(size_t)(n - j) * FFT_B + b] = a; im[(size_t)(n - j) * FFT_B + b] = -c;
        }
      }
    }
    fft_block(pl, re, im, wr, wi);
    for (int b = 0; b < nb; b++) {
      float *out = (float *)(base + (size_t)(r0 + b) * nh);
      for (int j = 0; j < n; j++) out[j] = (float)re[(size_t)j * FFT_B + b];
    }
  }
  free(buf);
}

void oracle_fft_axis0(int n, int ny, int nh, float _Complex *data)
{
  fft_plan pl;
  plan_make(&pl, n);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ny; y++) fft_strided_lines(&pl, data + (size_t)y * nh, (size_t)ny * nh, nh);
  plan_free(&pl);
}

void oracle_fft_axis1_c2r_axis2(int n, int nz, float _Complex *data)
{
  const int nh = n / 2 + 1;
  fft_plan pl;
  plan_make(&pl, n);
#pragma omp parallel for schedule(static)
  for (int z = 0; z < nz; z++) {
    float _Complex *plane = data + (size_t)z * n * nh;
    fft_strided_lines(&pl, plane, (size_t)nh, nh);
    c2r_rows(&pl, plane, n);
  }
  plan_free(&pl);
}

void oracle_c2r_3d_inplace(int n, float _Complex *data)
{
  oracle_fft_axis0(n, n, n / 2 + 1, data);
  oracle_fft_axis1_c2r_axis2(n, n, data);
}
