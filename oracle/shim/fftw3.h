/* TEST INFRASTRUCTURE ONLY -- stand-in for <fftw3.h> (FFTW 3.x, unpinned; absent from this image).
 * Only the single-precision c2r 3-D entry points GetHI calls (fourier.c:78-232) are provided; they
 * run oracle/fft3d.c (mixed-radix Stockham, double internally).  common.h includes <complex.h>
 * before this header, so the complex types are the C99 ones, as with real FFTW. */
#ifndef SHIM_FFTW3_H
#define SHIM_FFTW3_H
#include <stddef.h>
#include <complex.h>
typedef float _Complex fftwf_complex;
typedef double _Complex fftw_complex;
typedef struct shim_fftwf_plan_s *fftwf_plan;
#define FFTW_ESTIMATE (1U << 6)
fftwf_complex *fftwf_alloc_complex(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_c2r_3d(int n0, int n1, int n2, fftwf_complex *in, float *out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
int fftwf_init_threads(void);
void fftwf_plan_with_nthreads(int n);
void fftwf_cleanup_threads(void);
/* hooks for the oracle harness (inject / dump the k-space input and real-space output) */
typedef void (*shim_fft_hook)(int call_index, int n, fftwf_complex *kspace_or_null, float *real_or_null);
void shim_fftw_set_hooks(shim_fft_hook before, shim_fft_hook after);
void shim_fftw_reset_call_index(void);
#endif
