/* TEST INFRASTRUCTURE ONLY -- flat C API over the UNMODIFIED reference objects (compiled from
 * /root/reference/src by oracle/Makefile into oracle/_ref/libgethi_ref.so).  It lets tests/ and
 * bench.py's cpu_baseline / --impl reference arm drive the reference's own functions stage by stage
 * (common_gh.h:212-251) and read the state they produce.  Nothing in the product path may link this. */
#include <string.h>
#include <stdio.h>
#include "common_gh.h" /* the reference's header, found via -I/root/reference/src at build time */

void *ref_read_run_params(const char *fname)
{
  char buf[256];
  snprintf(buf, sizeof(buf), "%s", fname);
  return read_run_params(buf);
}

void ref_create_d_and_vr_fields(void *p) { create_d_and_vr_fields((ParamGetHI *)p); }
void ref_get_HI(void *p) { get_HI((ParamGetHI *)p); }
void ref_mk_T_maps(void *p) { mk_T_maps((ParamGetHI *)p); }
void ref_write_maps(void *p) { write_maps((ParamGetHI *)p); }
void ref_free(void *p) { param_gethi_free((ParamGetHI *)p); }

double ref_pk_linear0(void *p, double lgk) { return pk_linear0((ParamGetHI *)p, lgk); }
double ref_z_of_r(void *p, double r) { return z_of_r((ParamGetHI *)p, r); }
double ref_r_of_z(void *p, double z) { return r_of_z((ParamGetHI *)p, z); }
double ref_dgrowth_of_r(void *p, double r) { return dgrowth_of_r((ParamGetHI *)p, r); }
double ref_vgrowth_of_r(void *p, double r) { return vgrowth_of_r((ParamGetHI *)p, r); }
double ref_fraction_HI(double z) { return fraction_HI(z); }
double ref_bias_HI(double z) { return bias_HI(z); }

/* ---- point sources (do_psources = 1): the reference's own functions, plus a density rescaling for tests ---- */
void ref_setup_psources(void *p) { setup_psources((ParamGetHI *)p); }
void ref_get_point_sources(void *p) { get_point_sources((ParamGetHI *)p); }
void ref_mk_psources_maps(void *p) { mk_psources_maps((ParamGetHI *)p); }
double ref_n_of_z_psources(void *p, double z) { return n_of_z_psources((ParamGetHI *)p, z); }
double ref_temp_of_l(void *p, double l0, double nu, double z, double r, double domega) { return temp_of_l((ParamGetHI *)p, l0, nu, z, r, domega); }
double ref_draw_luminosity(void *p, double z, unsigned int seed, int n, double *out)
{
  gsl_rng *rng = init_rng(seed);
  double sum = 0;
  for (int i = 0; i < n; i++) { out[i] = draw_luminosity((ParamGetHI *)p, z, rng); sum += out[i]; }
  end_rng(rng);
  return sum;
}
/* thin the catalogue by f: n(z) -> f n(z); the rejection envelope max_Lpdf = 1.1 max / n(z) scales by 1/f so that
 * draw_luminosity's distribution stays what it was (l_distribution divides by n_of_z_psources) */
void ref_scale_psources(void *vp, double f)
{
  ParamGetHI *p = (ParamGetHI *)vp;
  for (int i = 0; i < NZ_PSOURCES; i++) { p->nz_psources_arr[i] *= f; p->max_Lpdf_arr[i] /= f; }
}
int *ref_nsources(void *vp) { return ((ParamGetHI *)vp)->nsources; }
float *ref_maps_PS(void *vp) { return (float *)((ParamGetHI *)vp)->maps_PS; }

double ref_get_double(void *vp, const char *name)
{
  ParamGetHI *p = (ParamGetHI *)vp;
#define F(n, v) if (!strcmp(name, n)) return (double)(v)
  F("OmegaM", p->OmegaM); F("OmegaL", p->OmegaL); F("OmegaB", p->OmegaB); F("hhub", p->hhub);
  F("weos", p->weos); F("n_scal", p->n_scal); F("sig8", p->sig8); F("fgrowth_0", p->fgrowth_0);
  F("hubble_0", p->hubble_0); F("z_max", p->z_max); F("z_min", p->z_min); F("r_max", p->r_max);
  F("r_min", p->r_min); F("r2_smooth", p->r2_smooth); F("do_smoothing", p->do_smoothing);
  F("numk", p->numk); F("logkmax", p->logkmax); F("logkmin", p->logkmin); F("idlogk", p->idlogk);
  F("glob_idr", p->glob_idr); F("seed_rng", p->seed_rng); F("n_side", p->n_side); F("nu_max", p->nu_max);
  F("nu_min", p->nu_min); F("n_nu", p->n_nu); F("n_grid", p->n_grid); F("l_box", p->l_box);
  F("nz_here", p->nz_here); F("iz0_here", p->iz0_here); F("pos_obs0", p->pos_obs[0]);
  F("pos_obs1", p->pos_obs[1]); F("pos_obs2", p->pos_obs[2]); F("sigma2_gauss", p->sigma2_gauss);
  F("do_psources", p->do_psources);
#undef F
  fprintf(stderr, "ref_get_double: unknown field %s\n", name);
  return -1e300;
}

void ref_set_double(void *vp, const char *name, double v)
{
  ParamGetHI *p = (ParamGetHI *)vp;
  if (!strcmp(name, "sigma2_gauss")) p->sigma2_gauss = v;
  else if (!strcmp(name, "seed_rng")) p->seed_rng = (unsigned int)v;
  else fprintf(stderr, "ref_set_double: unknown field %s\n", name);
}

const double *ref_get_table(void *vp, const char *name, int *len)
{
  ParamGetHI *p = (ParamGetHI *)vp;
#define T(n, ptr, l) if (!strcmp(name, n)) { *len = (l); return (ptr); }
  T("logkarr", p->logkarr, p->numk) T("pkarr", p->pkarr, p->numk)
  T("z_arr_z2r", p->z_arr_z2r, NZ) T("r_arr_z2r", p->r_arr_z2r, NZ)
  T("z_arr_r2z", p->z_arr_r2z, NZ) T("r_arr_r2z", p->r_arr_r2z, NZ)
  T("growth_d_arr", p->growth_d_arr, NZ) T("growth_v_arr", p->growth_v_arr, NZ)
  T("nz_psources_arr", p->nz_psources_arr, NZ_PSOURCES) T("max_Lpdf_arr", p->max_Lpdf_arr, NZ_PSOURCES)
#ifdef _IRREGULAR_NUTABLE
  T("nu0_arr", p->nu0_arr, p->n_nu) T("nuf_arr", p->nuf_arr, p->n_nu)
#endif
#undef T
  *len = 0;
  return NULL;
}

float *ref_grid(void *vp, const char *name)
{
  ParamGetHI *p = (ParamGetHI *)vp;
  if (!strcmp(name, "dens")) return (float *)p->grid_dens;
  if (!strcmp(name, "vpot")) return (float *)p->grid_vpot;
  if (!strcmp(name, "rvel")) return (float *)p->grid_rvel;
  if (!strcmp(name, "maps_HI")) return (float *)p->maps_HI;
  return NULL;
}

/* ---- FFT-boundary injection / capture (through the FFTW shim's hooks) ---- */
static float _Complex *inject_k[2], *capture_k[2];

static void before_fft(int call, int n, fftwf_complex *k, float *unused)
{
  (void)unused;
  size_t cnt = (size_t)n * n * (n / 2 + 1);
  int which = call & 1; /* call 0: density, call 1: velocity potential (fourier.c:391-392) */
  if (inject_k[which]) memcpy(k, inject_k[which], cnt * sizeof(float _Complex));
  if (capture_k[which]) memcpy(capture_k[which], k, cnt * sizeof(float _Complex));
}

void ref_set_fft_io(float _Complex *inject_dens_k, float _Complex *inject_vpot_k, float _Complex *capture_dens_k,
                    float _Complex *capture_vpot_k)
{
  inject_k[0] = inject_dens_k; inject_k[1] = inject_vpot_k;
  capture_k[0] = capture_dens_k; capture_k[1] = capture_vpot_k;
  shim_fftw_reset_call_index();
  shim_fftw_set_hooks((inject_dens_k || inject_vpot_k || capture_dens_k || capture_vpot_k) ? before_fft : NULL, NULL);
}
