/* Field access by name for bindings and tests (so they need not mirror the struct layout). */
#include <string.h>
#include "gh_host.h"

double gh_param_double(const ParamGetHI *p, const char *name)
{
#define F(n, v) if (!strcmp(name, n)) return (double)(v)
  F("OmegaM", p->OmegaM); F("OmegaL", p->OmegaL); F("OmegaB", p->OmegaB); F("hhub", p->hhub); F("weos", p->weos);
  F("n_scal", p->n_scal); F("sig8", p->sig8); F("fgrowth_0", p->fgrowth_0); F("hubble_0", p->hubble_0);
  F("z_max", p->z_max); F("z_min", p->z_min); F("r_max", p->r_max); F("r_min", p->r_min); F("r2_smooth", p->r2_smooth);
  F("do_smoothing", p->do_smoothing); F("numk", p->numk); F("logkmax", p->logkmax); F("logkmin", p->logkmin);
  F("idlogk", p->idlogk); F("glob_idr", p->glob_idr); F("seed_rng", p->seed_rng); F("n_side", p->n_side);
  F("nu_max", p->nu_max); F("nu_min", p->nu_min); F("n_nu", p->n_nu); F("n_grid", p->n_grid); F("l_box", p->l_box);
  F("nz_here", p->nz_here); F("iz0_here", p->iz0_here); F("pos_obs0", p->pos_obs[0]); F("pos_obs1", p->pos_obs[1]);
  F("pos_obs2", p->pos_obs[2]); F("sigma2_gauss", p->sigma2_gauss); F("mean_gauss", p->mean_gauss);
  F("do_psources", p->do_psources); F("irregular_nutable", p->irregular_nutable); F("n_shells_here", p->n_shells_here);
  F("shell0_here", p->shell0_here); F("nz_tab", GH_NZ); F("dz_tab", GH_DZ);
#undef F
  return -1e300;
}

const double *gh_param_table(const ParamGetHI *p, const char *name, int *len)
{
#define T(n, ptr, l) if (!strcmp(name, n)) { *len = (l); return (ptr); }
  T("logkarr", p->logkarr, p->numk) T("pkarr", p->pkarr, p->numk)
  T("z_arr_z2r", p->z_arr_z2r, GH_NZ) T("r_arr_z2r", p->r_arr_z2r, GH_NZ)
  T("z_arr_r2z", p->z_arr_r2z, GH_NZ) T("r_arr_r2z", p->r_arr_r2z, GH_NZ)
  T("growth_d_arr", p->growth_d_arr, GH_NZ) T("growth_v_arr", p->growth_v_arr, GH_NZ)
  T("frac_HI_arr", p->frac_HI_arr, GH_NZ) T("bias_HI_arr", p->bias_HI_arr, GH_NZ)
  T("nz_psources_arr", p->nz_psources_arr, GH_NZ_PSOURCES) T("max_Lpdf_arr", p->max_Lpdf_arr, GH_NZ_PSOURCES)
  T("nu0_arr", p->nu0_arr, p->nu0_arr ? p->n_nu : 0) T("nuf_arr", p->nuf_arr, p->nuf_arr ? p->n_nu : 0)
#undef T
  *len = 0;
  return NULL;
}

const float *gh_param_maps(const ParamGetHI *p) { return p->maps_HI; }
const char *gh_param_prefix(const ParamGetHI *p) { return p->prefixOut; }

/* setup_psources on a parsed parameter set, and the block that crosses the C-ABI (valid while `p` lives) */
void gh_inspect_psources(ParamGetHI *p, gh_cuda_psources_params *out)
{
  setup_psources(p);
  gh_fill_psources_params(p, out);
}
