/* TEST INFRASTRUCTURE ONLY (oracle).  3-D complex-to-real inverse DFT with FFTW's conventions
 * (fftwf_plan_dft_c2r_3d as called from fourier.c:78-99): unnormalised, exponent sign +, input the
 * non-redundant half spectrum [n][n][n/2+1] complex-float, output in place as padded reals
 * [n][n][2(n/2+1)] (last two floats of each row are padding).  FFTW (3.x, unpinned) is an absent
 * third-party dependency; this restates the published row-column algorithm: complex transforms over
 * the two slow axes, then a half-complex-to-real transform over the fast axis, which reads neither
 * Im(DC) nor Im(Nyquist) -- so the reference's non-Hermitian kx=0 / kx=n/2 planes are projected
 * exactly as FFTW's rdft2 does.  Arithmetic is double internally, rounded to float after each axis. */
#ifndef ORACLE_FFT3D_H
#define ORACLE_FFT3D_H
#include <complex.h>
/* in-place; data holds n*n*(n/2+1) float complex on entry, n*n*2(n/2+1) floats on exit */
void oracle_c2r_3d_inplace(int n, float _Complex *data);
/* slab variant used by the multi-rank restatement: complex transform over axis 0 only for a block
 * of `ny` rows (layout [n][ny][nh]) */
void oracle_fft_axis0(int n, int ny, int nh, float _Complex *data);
/* complex transform over axis 1 then c2r over axis 2 for `nz` planes (layout [nz][n][nh]) */
void oracle_fft_axis1_c2r_axis2(int n, int nz, float _Complex *data);
/* 1-D helpers (double), exponent sign +, unnormalised; any n (radix 2,3,4,5 + generic) */
void oracle_fft1d(int n, double _Complex *x);
#endif
