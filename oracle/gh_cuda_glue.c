/* gh_cuda_glue.c -- link with -lgh_cuda instead of fourier.o / pixelize.o and FFTW */
#include "common_gh.h"
#include "gh_cuda.h"

static gh_cuda_ctx *ctx;
static double frac_tab[NZ], bias_tab[NZ];                            /* user_defined.c on the radial grid */

static void fill(const ParamGetHI *par, gh_cuda_params *p)          /* ParamGetHI -> POD mirror */
{
  memset(p, 0, sizeof(*p));
  p->n_grid = par->n_grid;   p->l_box = par->l_box;
  memcpy(p->pos_obs, par->pos_obs, sizeof(p->pos_obs));
  p->seed_rng = par->seed_rng; p->do_smoothing = par->do_smoothing; p->r2_smooth = par->r2_smooth;
  p->fgrowth_0 = par->fgrowth_0; p->hubble_0 = par->hubble_0;
  p->numk = par->numk; p->logkmin = par->logkmin; p->logkmax = par->logkmax; p->idlogk = par->idlogk;
  p->n_scal = par->n_scal; p->logkarr = par->logkarr; p->pkarr = par->pkarr;
  p->nz_tab = NZ; p->glob_idr = par->glob_idr; p->dz_tab = DZ;
  p->z_arr_r2z = par->z_arr_r2z; p->r_arr_r2z = par->r_arr_r2z;
  p->growth_d_arr = par->growth_d_arr; p->growth_v_arr = par->growth_v_arr;
  p->z_arr_z2r = par->z_arr_z2r; p->r_arr_z2r = par->r_arr_z2r;
  p->n_side = par->n_side; p->n_nu = par->n_nu; p->nu_min = par->nu_min; p->nu_max = par->nu_max;
#ifdef _IRREGULAR_NUTABLE
  p->irregular_nutable = 1; p->nu0_arr = par->nu0_arr; p->nuf_arr = par->nuf_arr;
#endif
  p->OmegaB = par->OmegaB; p->hhub = par->hhub;
  for (int i = 0; i < NZ; i++) {                                     /* edits to user_defined.c reach the device */
    frac_tab[i] = fraction_HI(par->z_arr_r2z[i]);
    bias_tab[i] = bias_HI(par->z_arr_r2z[i]);
  }
  p->frac_HI_arr = frac_tab; p->bias_HI_arr = bias_tab;
}

void init_fftw(ParamGetHI *par)                       /* src/fourier.c:101 */
{
  gh_cuda_params p; fill(par, &p);
  unsigned char id[GH_CUDA_UNIQUE_ID_BYTES];
  if (NodeThis == 0) gh_cuda_get_unique_id(id);
#ifdef _HAVE_MPI
  MPI_Bcast(id, sizeof(id), MPI_BYTE, 0, MPI_COMM_WORLD);            /* one MPI rank per GPU */
#endif
  if (gh_cuda_create(&p, NodeThis, NNodes, id, NodeThis /* local device */, &ctx))
    report_error(1, "%s\n", gh_cuda_last_error());
  gh_cuda_slab(ctx, &par->nz_here, &par->iz0_here);
  /* grid_dens / grid_vpot / grid_rvel stay NULL: the grids live on the device */
}

void create_d_and_vr_fields(ParamGetHI *par)          /* src/fourier.c:375 */
{
  double mean;
  if (gh_cuda_create_d_and_vr_fields(ctx, &par->sigma2_gauss, &mean)) report_error(1, "%s\n", gh_cuda_last_error());
  print_info(" <d>=%.3lE, <d^2>=%.3lE\n", mean, sqrt(par->sigma2_gauss));
}

void get_HI(ParamGetHI *par)    { if (gh_cuda_get_HI(ctx)) report_error(1, "%s\n", gh_cuda_last_error()); }

void mk_T_maps(ParamGetHI *par)                       /* src/pixelize.c:150 */
{
  int n_here, s0;  gh_cuda_shells(ctx, &n_here, &s0);
  /* this rank's shells land at their place in the reference's shell-major stack */
  if (gh_cuda_mk_T_maps(ctx, par->maps_HI + (size_t)s0 * 12 * par->n_side * par->n_side))
    report_error(1, "%s\n", gh_cuda_last_error());
  /* write_maps then loops over [s0, s0+n_here) on every rank instead of all shells on rank 0 */
}

void end_fftw(void)             { gh_cuda_destroy(ctx); ctx = NULL; }
