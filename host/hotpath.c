/* The five hot-path entry points with the reference's names (src/common_gh.h:237-251), implemented by
 * handing the run state to libgh_cuda.so.  A non-zero status from the C-ABI becomes the reference's
 * fatal-error behaviour: message on stderr, exit(1) (src/common_gh.c:104-120). */
#include <stdlib.h>
#include <string.h>
#include "gh_host.h"

int gh_runtime_device(void);
const void *gh_runtime_unique_id(void);

void gh_fill_cuda_params(const ParamGetHI *par, gh_cuda_params *p)
{
  memset(p, 0, sizeof(*p));
  p->n_grid = par->n_grid; p->l_box = par->l_box;
  for (int i = 0; i < 3; i++) p->pos_obs[i] = par->pos_obs[i];
  p->seed_rng = par->seed_rng; p->do_smoothing = par->do_smoothing; p->r2_smooth = par->r2_smooth;
  p->fgrowth_0 = par->fgrowth_0; p->hubble_0 = par->hubble_0;
  p->numk = par->numk; p->logkmin = par->logkmin; p->logkmax = par->logkmax; p->idlogk = par->idlogk; p->n_scal = par->n_scal;
  p->logkarr = par->logkarr; p->pkarr = par->pkarr;
  p->nz_tab = GH_NZ; p->glob_idr = par->glob_idr;
  p->z_arr_r2z = par->z_arr_r2z; p->r_arr_r2z = par->r_arr_r2z;
  p->growth_d_arr = par->growth_d_arr; p->growth_v_arr = par->growth_v_arr;
  p->z_arr_z2r = par->z_arr_z2r; p->r_arr_z2r = par->r_arr_z2r; p->dz_tab = GH_DZ;
  p->n_side = par->n_side; p->n_nu = par->n_nu; p->irregular_nutable = par->irregular_nutable;
  p->nu0_arr = par->nu0_arr; p->nuf_arr = par->nuf_arr; p->nu_min = par->nu_min; p->nu_max = par->nu_max;
  p->OmegaB = par->OmegaB; p->hhub = par->hhub;
  p->frac_HI_arr = par->frac_HI_arr; p->bias_HI_arr = par->bias_HI_arr;
}

static void check(int rc, const char *what)
{
  if (rc) report_error(1, "%s: %s\n", what, gh_cuda_last_error());
}

void init_fftw(ParamGetHI *par)
{
  gh_cuda_params p;
  gh_fill_cuda_params(par, &p);
  par->rank = NodeThis; par->nranks = NNodes; par->device = gh_runtime_device();
  check(gh_cuda_create(&p, par->rank, par->nranks, gh_runtime_unique_id(), par->device, &par->cuda), "init_fftw");
  check(gh_cuda_slab(par->cuda, &par->nz_here, &par->iz0_here), "init_fftw");
  check(gh_cuda_shells(par->cuda, &par->n_shells_here, &par->shell0_here), "init_fftw");
  /* allocate_maps (src/io_gh.c:60-67): this rank's shells, page-locked for the device->host copy */
  const size_t bytes = (size_t)(par->n_shells_here > 0 ? par->n_shells_here : 1) * 12 * par->n_side * par->n_side * sizeof(float);
  void *m = NULL;
  check(gh_cuda_host_alloc(&m, bytes), "allocate_maps");
  par->maps_HI = (float *)m;
}

void create_d_and_vr_fields(ParamGetHI *par)
{
  print_info("*** Creating Gaussian density field \n");
  if (NodeThis == 0) timer(0);
  check(gh_cuda_create_d_and_vr_fields(par->cuda, &par->sigma2_gauss, &par->mean_gauss), "create_d_and_vr_fields");
  double ms[GH_T_NSLOTS];
  check(gh_cuda_stage_times(par->cuda, ms), "create_d_and_vr_fields");
  print_info("Creating Fourier-space density and velocity potential \n>    Relative time ellapsed %.1lf ms\n", ms[GH_T_KGEN]);
  print_info("Transforming and normalizing density and velocity potential\n>    Relative time ellapsed %.1lf ms\n", ms[GH_T_FFT]);
  print_info("Calculating radial velocity \n>    Relative time ellapsed %.1lf ms\n", ms[GH_T_VEL]);
  print_info(" <d>=%.3lE, <d^2>=%.3lE\n", par->mean_gauss, par->sigma2_gauss > 0 ? __builtin_sqrt(par->sigma2_gauss) : 0.0);
  print_info("\n");
}

void get_HI(ParamGetHI *par)
{
  print_info("*** Gettin' HI\n");
  check(gh_cuda_get_HI(par->cuda), "get_HI");
  check(gh_cuda_synchronize(par->cuda), "get_HI");
  double ms[GH_T_NSLOTS];
  check(gh_cuda_stage_times(par->cuda, ms), "get_HI");
  print_info(">    Relative time ellapsed %.1lf ms\n\n", ms[GH_T_GETHI]);
}

void mk_T_maps(ParamGetHI *par)
{
  print_info("*** Making maps\n Collecting masses\n Normalizing to temperature\n");
  check(gh_cuda_mk_T_maps(par->cuda, par->maps_HI), "mk_T_maps");
  double ms[GH_T_NSLOTS];
  check(gh_cuda_stage_times(par->cuda, ms), "mk_T_maps");
  print_info(">    Relative time ellapsed %.1lf ms\n\n", ms[GH_T_MAPS] + ms[GH_T_REDUCE] + ms[GH_T_D2H]);
}

/* The same stage without waiting: accumulation, reduction, scaling and the download are queued; write_maps
 * starts on shell s as soon as it has landed (gh_cuda_wait_shells) and prints the stage's timing at its end. */
void mk_T_maps_begin(ParamGetHI *par)
{
  print_info("*** Making maps\n Collecting masses\n Normalizing to temperature\n");
  check(gh_cuda_mk_T_maps_begin(par->cuda, par->maps_HI), "mk_T_maps");
  par->maps_streaming = 1;
}

void end_fftw(ParamGetHI *par)
{
  if (!par || !par->cuda) return;
  if (par->maps_HI) gh_cuda_host_free(par->maps_HI);
  par->maps_HI = NULL;
  gh_cuda_destroy(par->cuda);
  par->cuda = NULL;
}
