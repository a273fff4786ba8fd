/* TEST INFRASTRUCTURE ONLY -- FFTW3 single-precision c2r 3-D entry points over oracle/fft3d.c, with
 * optional before/after hooks so the oracle harness can dump or inject the k-space field the
 * reference hands to FFTW (fourier.c:391-392) without touching any reference source. */
#include <stdlib.h>
#include <fftw3.h>
#include "../fft3d.h"

struct shim_fftwf_plan_s { int n; fftwf_complex *in; float *out; };

static shim_fft_hook hook_before, hook_after;
static int call_index;

void shim_fftw_set_hooks(shim_fft_hook before, shim_fft_hook after) { hook_before = before; hook_after = after; }
void shim_fftw_reset_call_index(void) { call_index = 0; }

fftwf_complex *fftwf_alloc_complex(size_t n)
{
  void *p = NULL;
  if (posix_memalign(&p, 64, n * sizeof(fftwf_complex))) return NULL;
  return (fftwf_complex *)p;
}
void fftwf_free(void *p) { free(p); }

fftwf_plan fftwf_plan_dft_c2r_3d(int n0, int n1, int n2, fftwf_complex *in, float *out, unsigned flags)
{
  (void)flags;
  if (n0 != n1 || n1 != n2 || (void *)in != (void *)out) return NULL; /* GetHI only plans cubic in-place */
  fftwf_plan p = malloc(sizeof(*p));
  p->n = n0; p->in = in; p->out = out;
  return p;
}
void fftwf_execute(const fftwf_plan p)
{
  if (hook_before) hook_before(call_index, p->n, p->in, NULL);
  oracle_c2r_3d_inplace(p->n, p->in);
  if (hook_after) hook_after(call_index, p->n, NULL, p->out);
  call_index++;
}
void fftwf_destroy_plan(fftwf_plan p) { free(p); }
int fftwf_init_threads(void) { return 1; }
void fftwf_plan_with_nthreads(int n) { (void)n; }
void fftwf_cleanup_threads(void) {}
