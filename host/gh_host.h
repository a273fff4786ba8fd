/*
 * gh_host.h -- host side of the B200-native GetHI: parameter file, cosmology tables, FITS output and the
 * five hot-path entry points with the reference's names, implemented over the C-ABI of libgh_cuda.so.
 *
 * It mirrors the interface a GetHI user sees (reference src/common_gh.h:133-251, src/main_gh.c:24-80):
 *     ./GetHI <param_file>          same keys / quirks as src/io_gh.c:188-296
 *     <prefix>_%03d.fits            same BINTABLE layout as src/healpix_extra.c:132-164
 *     <prefix>_nuTable.dat          same text format as src/io_gh.c:84-107
 * Host code is plain C99; all grid work happens on the GPU.  There is no CPU fallback.
 */
#ifndef GH_HOST_H
#define GH_HOST_H

#include <stdio.h>
#include "../include/gh_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GH_NZ 5001     /* reference NZ, src/common_gh.h:34 */
#define GH_DZ 0.001    /* reference DZ, src/common_gh.h:33 */
#define GH_NU_21 1420.40575177
#define GH_NZ_PSOURCES 256 /* reference NZ_PSOURCES, src/common_gh.h:35 */

/* The run state.  Field names follow ParamGetHI (src/common_gh.h:138-209) where they carry the same
 * meaning; device buffers live behind `cuda`, so there are no host grid pointers to free. */
typedef struct ParamGetHI {
  char fnamePk[256];
  double OmegaM, OmegaL, OmegaB, hhub, weos, n_scal, sig8;
  double fgrowth_0, hubble_0, z_max, z_min, r_max, r_min, r2_smooth;
  int do_smoothing;
  int numk;
  double logkmax, logkmin, idlogk;
  double *logkarr, *pkarr;
  double z_arr_z2r[GH_NZ], r_arr_z2r[GH_NZ], z_arr_r2z[GH_NZ], r_arr_r2z[GH_NZ];
  double growth_d_arr[GH_NZ], growth_v_arr[GH_NZ];
  /* fraction_HI / bias_HI (user_defined.c) sampled at z_arr_r2z: how the user hooks reach the device */
  double frac_HI_arr[GH_NZ], bias_HI_arr[GH_NZ];
  double glob_idr;
  unsigned int seed_rng;
  int irregular_nutable; /* the reference's -D_IRREGULAR_NUTABLE compile-time personality, here a run-time flag */
  char fnameNuTable[256];
  double *nu0_arr, *nuf_arr;
  long n_side;
  double nu_max, nu_min;
  int n_nu;
  int n_grid;
  double l_box;
  int nz_here, iz0_here;
  char prefixOut[256];
  double pos_obs[3];
  int do_psources;
  /* point sources (src/common_gh.h:111-116 + the tabulated user functions that cross the C-ABI, host/psources.c) */
  double glob_dz, glob_inv_dz;
  double nz_psources_arr[GH_NZ_PSOURCES], max_Lpdf_arr[GH_NZ_PSOURCES], ps_bias_arr[GH_NZ_PSOURCES];
  double *ps_lcdf, *ps_sed;
  double ps_lognu_min, ps_lognu_max;
  float *maps_PS; /* this rank's shells of the point-source maps, page-locked */
  double sigma2_gauss, mean_gauss;
  /* this rank's shells of the finished map stack, [n_shells_here][12 n_side^2], page-locked */
  float *maps_HI;
  int n_shells_here, shell0_here;
  int maps_streaming; /* mk_T_maps_begin was called: write_maps waits shell by shell for the download */
  /* process layout: one process per GPU */
  int rank, nranks, device;
  gh_cuda_ctx *cuda;
} ParamGetHI;

/* process bring-up: src/common_gh.c:31 (mpi_init).  Rank / size / NCCL id come from the launcher (host/main.c):
 * GH_NGPUS=P forks P ranks and pipes the id; or GH_RANK, GH_NRANKS (or RANK / WORLD_SIZE / LOCAL_RANK) with the id
 * in the file GH_UNIQUE_ID_FILE */
extern int NodeThis, NNodes;
void gh_mpi_init(int rank, int nranks, int device, const void *unique_id);
void print_info(const char *fmt, ...);
void report_error(int level, const char *fmt, ...);
void timer(int i);
void gh_phase(const char *name); /* GH_HOST_TIMING=1: time since the previous call, labelled, on stderr */

/* src/io_gh.c */
ParamGetHI *read_run_params(const char *fname);
ParamGetHI *read_run_params_ex(const char *fname, int with_device);  /* with_device=0: parse + cosmology only */
void write_maps(ParamGetHI *par);
void param_gethi_free(ParamGetHI *par);
/* src/cosmo.c */
void cosmo_set(ParamGetHI *par);
double pk_linear0(const ParamGetHI *par, double lgk);
double r_of_z(const ParamGetHI *par, double z);
double z_of_r(const ParamGetHI *par, double r);
double dgrowth_of_r(const ParamGetHI *par, double r);
double vgrowth_of_r(const ParamGetHI *par, double r);
/* src/user_defined.c */
double fraction_HI(double z);
double bias_HI(double z);
/* src/fourier.c, src/grid_tools.c, src/pixelize.c: the hot path, on the GPU */
void init_fftw(ParamGetHI *par);
void create_d_and_vr_fields(ParamGetHI *par);
void get_HI(ParamGetHI *par);
void mk_T_maps(ParamGetHI *par);
void mk_T_maps_begin(ParamGetHI *par); /* non-blocking form; write_maps then overlaps the download */
void end_fftw(ParamGetHI *par);

/* src/psources.c, src/grid_tools.c:24-101, src/pixelize.c:58-148 (do_psources = 1) */
void setup_psources(ParamGetHI *par);
void get_point_sources(ParamGetHI *par);
void mk_psources_maps(ParamGetHI *par);
double n_of_z_psources(const ParamGetHI *par, double z);
double bias_psources(double z);
double temp_of_l(const ParamGetHI *par, double L0, double nu_obs, double z, double r, double dOmega);
void gh_fill_psources_params(const ParamGetHI *par, gh_cuda_psources_params *out);

/* helpers */
void gh_fill_cuda_params(const ParamGetHI *par, gh_cuda_params *out);
int gh_write_healpix_map(const float *map, long nside, const char *fname); /* src/healpix_extra.c:132-164 */
int gh_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif
