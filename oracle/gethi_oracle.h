/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked, imported or executed by the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
 *
 * Plain-C CPU restatement of the GetHI hot path of damonge/CRIME, one function per reference stage,
 * each citing the reference file:line it follows.  It is pinned against the unmodified reference
 * compiled in oracle/_ref/ (tests/test_oracle_vs_ref.py, run wherever /root/reference or a prebuilt
 * oracle/_ref exists) and against the golden vectors in tests/golden/ that were generated from it.
 * Third-party arithmetic (GSL mt19937, FFTW c2r, chealpix vec2pix_ring) is restated from the published
 * algorithms in mt19937.c / fft3d.c / healpix_ring.c; the reference pins no version and has no tests of
 * its own, so for those three boundaries PARITY IS UNPINNED beyond our own known-answer tests.
 *
 * Arithmetic contract (reference Makefile:6,43: gcc -O3, no -ffast-math, no -march): IEEE double,
 * no FMA contraction, float storage (-D_SPREC).  Compile with -ffp-contract=off.
 */
#ifndef GETHI_ORACLE_H
#define GETHI_ORACLE_H
#include <stdint.h>
#include <complex.h>
#include "../include/gh_cuda.h" /* gh_cuda_params: the POD mirror of ParamGetHI shared with the product ABI */

/* slab description: ParamGetHI.nz_here / iz0_here (fourier.c:141-148,169-171) */
typedef struct { int nz_here, iz0_here; } oracle_slab;

/* --- cosmo.c / user_defined.c table look-ups --- */
double oracle_pk_linear0(const gh_cuda_params *p, double lgk);    /* cosmo.c:153-170 */
double oracle_r_of_z(const gh_cuda_params *p, double z);          /* cosmo.c:40-50 */
double oracle_z_of_r(const gh_cuda_params *p, double r);          /* cosmo.c:52-62 */
double oracle_dgrowth_of_r(const gh_cuda_params *p, double r);    /* cosmo.c:64-74 */
double oracle_vgrowth_of_r(const gh_cuda_params *p, double r);    /* cosmo.c:76-86 */
void oracle_set_user_defined(double a, double p, double b0, double b1, double q); /* default: the shipped 0.008, 0.6, 0.904, 0.135, 1.696 */
double oracle_fraction_HI(double z);                              /* user_defined.c:27-30 */
double oracle_bias_HI(double z);                                  /* user_defined.c:32-35 */

/* --- fourier.c --- */
/* the reference's own stream: per-OpenMP-thread MT19937 seeded seed+ithr, planes statically chunked
 * (fourier.c:234-305, common.c:154-164); n_threads fixes the realisation */
void oracle_kgen_mt19937(const gh_cuda_params *p, oracle_slab s, int n_threads, float _Complex *dens_k,
                         float _Complex *vpot_k);
/* the product's stream: Philox4x32-10 keyed on (seed, global mode index); same field maths */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void oracle_kgen_philox(const gh_cuda_params *p, int ky0, int nky, float _Complex *dens_k, float _Complex *vpot_k,
                        int transposed_layout);
/* fourier.c:397-412 */
void oracle_normalize(const gh_cuda_params *p, oracle_slab s, float *dens, float *vpot);
/* fourier.c:307-373; slice_left/right are the halo planes of fourier.c:415-428 */
void oracle_radial_velocity(const gh_cuda_params *p, oracle_slab s, const float *vpot, const float *slice_left,
                            const float *slice_right, float *rvel);
/* fourier.c:24-76: this slab's contribution; returns sum(d)/ng_tot and sum(d*d)/ng_tot */
void oracle_sigma_partial(const gh_cuda_params *p, oracle_slab s, const float *dens, double *mean_part,
                          double *sigma2_part);
/* fourier.c:375-438 for a single slab holding the whole box, starting from given k-space fields */
void oracle_fields_from_k(const gh_cuda_params *p, float _Complex *dens_k_inout, float _Complex *vpot_k_inout,
                          float *rvel_out, double *sigma2_gauss, double *mean_gauss);

/* --- grid_tools.c:103-153 --- */
void oracle_get_HI(const gh_cuda_params *p, oracle_slab s, double sigma2_gauss, float *dens, float *rvel);

/* --- pixelize.c --- */
void oracle_subparticle_offsets(const gh_cuda_params *p, double *xyz30);                 /* pixelize.c:157-164 */
int oracle_get_inu(const gh_cuda_params *p, double nu, int inu_start);                   /* pixelize.c:28-55 */
int oracle_shell_of_nu(const gh_cuda_params *p, double nu, int inu_prev);                /* pixelize.c:213-217 */
/* pixelize.c:186-231; maps is [n_nu][npix], accumulated in float exactly like `flouble` maps_HI */
void oracle_accumulate_maps(const gh_cuda_params *p, oracle_slab s, const float *mass, const float *dz_rsd,
                            float *maps);
/* pixelize.c:236-261 */
void oracle_normalize_maps(const gh_cuda_params *p, float *maps);
void oracle_shell_prefactors(const gh_cuda_params *p, double *prefac); /* [n_nu] */
/* (shell, pixel) of arbitrary observer-centred points, the inner body of pixelize.c:206-223 */
void oracle_points_to_shell_pixel(const gh_cuda_params *p, const double *pos, const double *dz_rsd, long long n,
                                  int *shell_out, long long *pix_out);
#endif
