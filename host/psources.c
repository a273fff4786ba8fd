/* Point sources, host side (reference src/psources.c, src/grid_tools.c:24-101, src/pixelize.c:58-148): the three
 * user-definable functions (luminosity function, SED, bias), setup_psources' redshift tables, and the tabulations
 * the device needs of the user functions; get_point_sources / mk_psources_maps hand the work to libgh_cuda.so. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "gh_host.h"

#define NL_PSOURCES 256 /* src/psources.c:24 */
#define LLOGMIN -5.     /* only luminosities between 10^17 and 10^28 W/Hz are considered, src/psources.c:25-26 */
#define LLOGMAX 6.
#define GH_PS_NL 1024   /* luminosity bins of the cumulative distribution handed to the device */
#define GH_PS_NSED 4096 /* SED samples, uniform in log10(nu) */

/* ---- user-definable functions (src/psources.c:28-71; edit them here as one would there) ---- */
#define LX_TOY 2.11
#define ALPHA_TOY -0.633
#define RHO_TOY 3.2E-4
/* luminosity function dn(z, L)/dlog10(L), L in units of 10^22 W/Hz (src/psources.c:38-52) */
static double l_z_function(double L, double z)
{
  const double rho_l = 2.5 * M_LN10 * RHO_TOY * pow(L / LX_TOY, ALPHA_TOY + 1) * exp(-L / LX_TOY); /* RHO_TOY is per magnitude */
  const double fz = (z < 1.5) ? pow(1 + z, 3.1) : 17.124;
  return fz * rho_l;
}
#define SED_NORM 0.114461
/* SED, normalised to 1 at the frequency of the luminosity function; nu in MHz (src/psources.c:55-64) */
static double spec_ed(double nu)
{
  const double nu_GHz = nu * 0.001;
  return SED_NORM * (pow(nu_GHz, -0.1) + 10 * pow(nu_GHz, -0.75));
}
double bias_psources(double z) { (void)z; return 1; } /* src/psources.c:66-69 */
/* ---- end of user-definable functions ---- */

/* src/psources.c:74-91 */
double n_of_z_psources(const ParamGetHI *par, double z)
{
  const int iz = (int)(z * par->glob_inv_dz);
  if (iz >= GH_NZ_PSOURCES || iz < 0) return -1;
  if (iz == GH_NZ_PSOURCES - 1) return par->nz_psources_arr[GH_NZ_PSOURCES - 1];
  const double zi = iz * par->glob_dz;
  return par->nz_psources_arr[iz] + (par->nz_psources_arr[iz + 1] - par->nz_psources_arr[iz]) * (z - zi) * par->glob_inv_dz;
}

/* src/psources.c:98-131, plus what crosses the C-ABI: per redshift bin the cumulative distribution of log10 L (the
 * reference samples it by rejection under 1.1 x its maximum; the device inverts the table), bias_psources and
 * spec_ed on a frequency grid that covers every (1 + z) nu_obs of the run */
void setup_psources(ParamGetHI *par)
{
  const double dlogL = (LLOGMAX - LLOGMIN) / NL_PSOURCES;
  par->glob_dz = par->z_max / GH_NZ_PSOURCES;
  par->glob_inv_dz = 1. / par->glob_dz;
  if (!par->ps_lcdf) par->ps_lcdf = (double *)malloc(sizeof(double) * GH_NZ_PSOURCES * (GH_PS_NL + 1));
  if (!par->ps_sed) par->ps_sed = (double *)malloc(sizeof(double) * GH_PS_NSED);
  if (!par->ps_lcdf || !par->ps_sed) report_error(1, "setup_psources: out of memory\n");
  for (int ii = 0; ii < GH_NZ_PSOURCES; ii++) {
    const double z = ii * par->glob_dz;
    double max_dist = -1;
    par->nz_psources_arr[ii] = 0;
    for (int jj = 0; jj < NL_PSOURCES; jj++) {
      const double logL = LLOGMIN + (jj + 0.5) * dlogL;
      const double lfunc = l_z_function(pow(10., logL), z);
      if (lfunc >= max_dist) max_dist = lfunc;
      par->nz_psources_arr[ii] += dlogL * lfunc;
    }
    par->max_Lpdf_arr[ii] = par->nz_psources_arr[ii] > 0 ? 1.1 * max_dist / par->nz_psources_arr[ii] : 0;
    par->ps_bias_arr[ii] = bias_psources(z);
    /* cumulative distribution on the finer grid, midpoint rule per bin, normalised to end at 1 */
    double *cdf = par->ps_lcdf + (size_t)ii * (GH_PS_NL + 1);
    const double dl = (LLOGMAX - LLOGMIN) / GH_PS_NL;
    cdf[0] = 0;
    for (int jj = 0; jj < GH_PS_NL; jj++) cdf[jj + 1] = cdf[jj] + dl * l_z_function(pow(10., LLOGMIN + (jj + 0.5) * dl), z);
    const double tot = cdf[GH_PS_NL];
    for (int jj = 0; jj <= GH_PS_NL; jj++) cdf[jj] = tot > 0 ? cdf[jj] / tot : (double)jj / GH_PS_NL;
  }
  /* test hook: GH_PSOURCES_THIN=f thins the catalogue (n(z) -> f n(z)); the full density is ~1e10 sources per box */
  const char *thin = getenv("GH_PSOURCES_THIN");
  if (thin && atof(thin) > 0) {
    for (int ii = 0; ii < GH_NZ_PSOURCES; ii++) { par->nz_psources_arr[ii] *= atof(thin); par->max_Lpdf_arr[ii] /= atof(thin); }
  }
  /* rest-frame frequencies reached: nu_obs in the shells, z_true in [~-0.1, z_max + 0.1] */
  double nu_lo, nu_hi;
  if (par->irregular_nutable) { nu_lo = par->nu0_arr[0]; nu_hi = par->nuf_arr[par->n_nu - 1]; }
  else { nu_lo = par->nu_min; nu_hi = par->nu_max; }
  par->ps_lognu_min = log10(0.5 * nu_lo);
  par->ps_lognu_max = log10(nu_hi * (2.2 + par->z_max));
  for (int i = 0; i < GH_PS_NSED; i++)
    par->ps_sed[i] = spec_ed(pow(10., par->ps_lognu_min + (par->ps_lognu_max - par->ps_lognu_min) * i / (GH_PS_NSED - 1)));
}

void gh_fill_psources_params(const ParamGetHI *par, gh_cuda_psources_params *p)
{
  memset(p, 0, sizeof(*p));
  p->nz = GH_NZ_PSOURCES; p->z_max = par->z_max; p->nz_arr = par->nz_psources_arr; p->bias_arr = par->ps_bias_arr;
  p->nl = GH_PS_NL; p->logl_min = LLOGMIN; p->logl_max = LLOGMAX; p->lcdf = par->ps_lcdf;
  p->nsed = GH_PS_NSED; p->lognu_min = par->ps_lognu_min; p->lognu_max = par->ps_lognu_max; p->sed_arr = par->ps_sed;
  p->hhub = par->hhub;
}

/* src/psources.c:159-167 (the device evaluates the same expression; here for the host-side tests) */
double temp_of_l(const ParamGetHI *par, double L0, double nu_obs, double z, double r, double dOmega)
{
  const double s_nu = 8.35774E7 * 4 * M_PI * L0 * spec_ed((1 + z) * nu_obs) * par->hhub * par->hhub / (r * r * (1 + z));
  return 3.2548291E-2 * s_nu / (dOmega * nu_obs * nu_obs);
}

static void ps_check(int rc, const char *what)
{
  if (rc) report_error(1, "%s: %s\n", what, gh_cuda_last_error());
}

/* src/grid_tools.c:24-101 */
void get_point_sources(ParamGetHI *par)
{
  gh_cuda_psources_params p;
  long long np_tot = 0;
  print_info("*** Getting point sources\n");
  if (NodeThis == 0) timer(0);
  print_info("Poisson-sampling\n");
  gh_fill_psources_params(par, &p);
  ps_check(gh_cuda_get_point_sources(par->cuda, &p, &np_tot), "get_point_sources");
  if (NodeThis == 0) timer(2);
  print_info("  There will be %ld particles in total \n", (long)np_tot);
}

/* src/pixelize.c:58-148 */
void mk_psources_maps(ParamGetHI *par)
{
  print_info("*** Making source maps\n");
  if (NodeThis == 0) timer(0);
  if (!par->maps_PS) {
    void *m = NULL;
    const size_t bytes = (size_t)(par->n_shells_here > 0 ? par->n_shells_here : 1) * 12 * par->n_side * par->n_side * sizeof(float);
    ps_check(gh_cuda_host_alloc(&m, bytes), "allocate_maps");
    par->maps_PS = (float *)m;
  }
  ps_check(gh_cuda_mk_psources_maps(par->cuda, par->maps_PS), "mk_psources_maps");
  if (NodeThis == 0) timer(2);
  print_info("\n");
}
