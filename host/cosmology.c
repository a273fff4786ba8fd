/* Host cosmology of GetHI: background distances, growth, the interpolation tables the device reads, and
 * the sigma_8 normalisation of the input P(k).  Same definitions as reference src/cosmo.c:232-413 and
 * src/cosmo_mad.c:148-409 (flat or curved w0 dark energy, no radiation), written from the formulae with
 * an own adaptive Gauss-Legendre integrator instead of GSL's qng / qagil. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "gh_host.h"

#define HMPC 2997.92458 /* c/H0 in Mpc/h */

typedef struct { double OM, OL, OK, w0; int lambda; } Bg;

/* ---- adaptive 8-point Gauss-Legendre ---- */
static const double GLX[4] = {0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363};
static const double GLW[4] = {0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763};

typedef double (*fn1)(double, const void *);

static double gl8(fn1 f, const void *p, double a, double b)
{
  const double c = 0.5 * (a + b), h = 0.5 * (b - a);
  double s = 0;
  for (int i = 0; i < 4; i++) s += GLW[i] * (f(c - h * GLX[i], p) + f(c + h * GLX[i], p));
  return s * h;
}

static double adaptive(fn1 f, const void *p, double a, double b, double whole, double tol, int depth)
{
  const double c = 0.5 * (a + b);
  const double l = gl8(f, p, a, c), r = gl8(f, p, c, b);
  if (depth > 40 || fabs(l + r - whole) <= tol) return l + r;
  return adaptive(f, p, a, c, l, 0.5 * tol, depth + 1) + adaptive(f, p, c, b, r, 0.5 * tol, depth + 1);
}

static double integrate(fn1 f, const void *p, double a, double b, double rel)
{
  const double whole = gl8(f, p, a, b);
  double res = adaptive(f, p, a, b, whole, fabs(whole) * rel + 1e-300, 0);
  return adaptive(f, p, a, b, whole, fabs(res) * rel + 1e-300, 0);
}

/* ---- background ---- */
static double de_term(const Bg *bg, double a) { return bg->lambda ? bg->OL * a * a * a : bg->OL * pow(a, -3 * bg->w0); }
/* a^3 E^2(a) / ... : OM + OL a^{-3w} + OK a */
static double e2a3(const Bg *bg, double a) { return bg->OM + de_term(bg, a) + bg->OK * a; }
static double f_dchi(double a, const void *p) { return 1.0 / sqrt(a * e2a3((const Bg *)p, a)); }       /* H0/(a^2 H) */
static double f_growth(double a, const void *p) { const double q = sqrt(a / e2a3((const Bg *)p, a)); return q * q * q; } /* (H0/(aH))^3 */

static double a_equality(const Bg *bg)
{
  double ak = 1, al = 1;
  if (fabs(bg->OK) >= 1e-6) ak = bg->OM / fabs(bg->OK);
  if (bg->OL != 0) al = bg->lambda ? pow(bg->OM / bg->OL, 0.333) : pow(bg->OM / bg->OL, -1 / (3 * bg->w0));
  return ak < al ? ak : al;
}

/* comoving radial distance chi(a) in Mpc/h */
static double chi_of_a(const Bg *bg, double a) { return a >= 1 ? 0.0 : HMPC * integrate(f_dchi, bg, a, 1.0, 1e-10); }

/* growth factor normalised to D ~ a deep in matter domination */
static double growth_of_a(const Bg *bg, double a, double alim)
{
  if (a <= alim) return a;
  const double int0 = 0.4 * sqrt(alim * alim * alim * alim * alim / (bg->OM * bg->OM * bg->OM));
  const double in = integrate(f_growth, bg, alim, a, 1e-10);
  return (int0 + in) * 2.5 * bg->OM / (a * sqrt(a / e2a3(bg, a)));
}

static double hubble_of_a(const Bg *bg, double a) { return sqrt(e2a3(bg, a) / (a * a * a)) / HMPC; }

/* f = dlnD/dlna from the integral form of D */
static double fgrowth_of_a(const Bg *bg, double a, double alim)
{
  const double D = growth_of_a(bg, a, alim);
  double coeff = 0, apow = a * a * a;
  if (!bg->lambda) { coeff = 1 + bg->w0; apow = pow(a, -3 * bg->w0); }
  return 0.5 * (5 * bg->OM * a / D - (3 * bg->OM + 3 * coeff * bg->OL * apow + 2 * bg->OK * a)) /
         (bg->OM + bg->OL * apow + bg->OK * a);
}

/* ---- table look-ups (what the device also evaluates per cell) ---- */
double r_of_z(const ParamGetHI *p, double z)
{
  if (z <= 0) return 0;
  if (z >= p->z_arr_z2r[GH_NZ - 1]) return p->r_arr_z2r[GH_NZ - 1];
  const int iz = (int)(z / GH_DZ);
  return p->r_arr_z2r[iz] + (p->r_arr_z2r[iz + 1] - p->r_arr_z2r[iz]) * (z - p->z_arr_z2r[iz]) / GH_DZ;
}

static double lerp_r(const ParamGetHI *p, const double *tab, double r, double at0)
{
  if (r <= 0) return at0;
  if (r >= p->r_arr_r2z[GH_NZ - 1]) return tab[GH_NZ - 1];
  const int ir = (int)(r * p->glob_idr);
  return tab[ir] + (tab[ir + 1] - tab[ir]) * (r - p->r_arr_r2z[ir]) * p->glob_idr;
}
double z_of_r(const ParamGetHI *p, double r) { return lerp_r(p, p->z_arr_r2z, r, 0); }
double dgrowth_of_r(const ParamGetHI *p, double r) { return lerp_r(p, p->growth_d_arr, r, 1); }
double vgrowth_of_r(const ParamGetHI *p, double r) { return lerp_r(p, p->growth_v_arr, r, 1); }

double pk_linear0(const ParamGetHI *p, double lgk)
{
  const int ik = (int)((lgk - p->logkmin) * p->idlogk);
  if (ik < 0) return p->pkarr[0] * pow(10, p->n_scal * (lgk - p->logkmin));
  if (ik < p->numk) {
    const double hi = ik + 1 < p->numk ? p->pkarr[ik + 1] : p->pkarr[ik]; /* the reference reads one past the end here */
    return p->pkarr[ik] + (lgk - p->logkarr[ik]) * (hi - p->pkarr[ik]) * p->idlogk;
  }
  return p->pkarr[p->numk - 1] * pow(10, -3 * (lgk - p->logkmax));
}

/* ---- sigma_8 ---- */
static double tophat(double x)
{
  if (x < 0.1) {
    const double x2 = x * x;
    return 1. - 0.1 * x2 + 0.003571429 * x2 * x2 - 6.61376E-5 * x2 * x2 * x2 + 7.51563E-7 * x2 * x2 * x2 * x2;
  }
  return 3 * (sin(x) - x * cos(x)) / (x * x * x);
}

typedef struct { const ParamGetHI *par; double R; } SigArg;
static double f_sigma(double logk, const void *vp)
{
  const SigArg *a = (const SigArg *)vp;
  const double k = pow(10, logk), w = tophat(k * a->R);
  return 0.1166503235296796 /* ln10/(2 pi^2) */ * pk_linear0(a->par, logk) * k * k * k * w * w;
}

/* variance in top-hat spheres of radius R: integral over log10 k from -inf to logkmax (src/cosmo.c:232-303) */
static double sigma2_tophat(const ParamGetHI *par, double R)
{
  SigArg a = {par, R};
  double sum = 0, hi = par->logkmax;
  /* the integrand falls like k^{3+ns}: sweep decades downwards until they stop contributing */
  for (int d = 0; d < 40; d++) {
    const double lo = hi - 0.5;
    const double part = integrate(f_sigma, &a, lo, hi, 1e-9);
    sum += part;
    hi = lo;
    if (d > 8 && fabs(part) < 1e-14 * fabs(sum)) break;
  }
  return sum;
}

static int count_lines(FILE *f)
{
  int n = 0;
  char buf[1000];
  while (fgets(buf, sizeof(buf), f)) n++;
  return n;
}

static void read_pk(ParamGetHI *par)
{
  print_info("Reading P_k from file: %s\n", par->fnamePk);
  FILE *f = fopen(par->fnamePk, "r");
  if (!f) { fprintf(stderr, "CRIME: Couldn't open file %s \n", par->fnamePk); exit(1); }
  par->numk = count_lines(f);
  rewind(f);
  par->logkarr = malloc(sizeof(double) * par->numk);
  par->pkarr = malloc(sizeof(double) * par->numk);
  if (!par->logkarr || !par->pkarr) { fprintf(stderr, "out of memory\n"); exit(1); }
  for (int i = 0; i < par->numk; i++) {
    double k, pk;
    if (fscanf(f, "%lf %lf", &k, &pk) != 2) { fprintf(stderr, "CRIME: Error reading file %s, line %d \n", par->fnamePk, i + 1); exit(1); }
    par->pkarr[i] = pk;
    par->logkarr[i] = log10(k);
  }
  fclose(f);
  par->logkmin = par->logkarr[0];
  par->logkmax = par->logkarr[par->numk - 1];
  par->idlogk = (par->numk - 1) / (par->logkarr[par->numk - 1] - par->logkarr[0]);
  const double s2 = sigma2_tophat(par, 8.0);
  print_info("  Original sigma8=%lf\n", sqrt(s2));
  const double norm = par->sig8 * par->sig8 / s2;
  for (int i = 0; i < par->numk; i++) par->pkarr[i] *= norm;
}

void cosmo_set(ParamGetHI *par)
{
  Bg bg;
  bg.OM = par->OmegaM; bg.OL = par->OmegaL; bg.OK = 1 - par->OmegaM - par->OmegaL; bg.w0 = par->weos;
  bg.lambda = fabs(par->weos + 1) < 1e-6;
  if (fabs(bg.OK) < 1e-6) bg.OK = bg.OK; /* the reference keeps the tiny residual too */
  if (par->OmegaM <= 0) { fprintf(stderr, "CRIME: Wrong matter parameter %.3lf \n", par->OmegaM); exit(1); }
  if (par->OmegaM < par->OmegaB) { fprintf(stderr, "CRIME: Wrong M/B parameter %.3lf > %.3lf \n", par->OmegaB, par->OmegaM); exit(1); }
  if (par->weos > -0.333333) { fprintf(stderr, "CRIME: DE is too exotic (w=%.3lf \n", par->weos); exit(1); }
  print_info("The cosmological model is:\n");
  print_info(" O_M=%.3f O_L=%.3f O_K=%.3f\n", bg.OM, bg.OL, bg.OK);
  print_info(" O_B=%.3f w=%.3f h=%.3f\n", par->OmegaB, par->weos, par->hhub);
  print_info(fabs(bg.OK) < 1e-6 ? " Flat universe, " : (bg.OK > 0 ? " Open universe, " : " Closed universe, "));
  print_info(bg.lambda ? "standard cosmological constant\n" : "non-standard dark energy\n");
  const double aeq = a_equality(&bg), alim = 0.01 * aeq;
  const double growth0 = growth_of_a(&bg, 1.0, alim);
  print_info("\n Time of equality: a_eq=%.5lf\n", aeq);
  print_info(" Particle horizon: chi_H(0)=%.3lE Mpc/h\n", HMPC * (2 * sqrt(alim / bg.OM) + integrate(f_dchi, &bg, alim, 1.0, 1e-10)));
  print_info(" Present growth factor: D_0=%.3lf\n\n", growth0);

  par->fgrowth_0 = fgrowth_of_a(&bg, 1.0, alim);
  par->hubble_0 = hubble_of_a(&bg, 1.0);
  par->z_min = GH_NU_21 / par->nu_max - 1;
  par->z_max = GH_NU_21 / par->nu_min - 1;
  par->r_min = chi_of_a(&bg, 1 / (1 + par->z_min));
  par->r_max = chi_of_a(&bg, 1 / (1 + par->z_max));
  par->l_box = 2 * par->r_max * (1 + 2. / par->n_grid);
  for (int i = 0; i < 3; i++) par->pos_obs[i] = 0.5 * par->l_box;

  for (int i = 0; i < GH_NZ; i++) {
    const double z = i * GH_DZ;
    par->z_arr_z2r[i] = z;
    par->r_arr_z2r[i] = chi_of_a(&bg, 1 / (1 + z));
  }
  if (par->z_arr_z2r[GH_NZ - 1] <= par->z_max || par->r_arr_z2r[GH_NZ - 1] <= par->r_max) report_error(1, "OMG!\n");
  par->glob_idr = (GH_NZ - 1) / (par->r_arr_z2r[GH_NZ - 1] - par->r_arr_z2r[0]);
  int iz = 0;
  for (int i = 0; i < GH_NZ; i++) {
    const double r = i / par->glob_idr;
    double z;
    if (r <= 0) z = 0;
    else if (r >= par->r_arr_z2r[GH_NZ - 1]) z = par->z_arr_z2r[GH_NZ - 1];
    else {
      while (r >= par->r_arr_z2r[iz]) iz++; /* monotone in i: resume the search */
      z = par->z_arr_z2r[iz - 1] + (par->z_arr_z2r[iz] - par->z_arr_z2r[iz - 1]) * (r - par->r_arr_z2r[iz - 1]) /
                                       (par->r_arr_z2r[iz] - par->r_arr_z2r[iz - 1]);
    }
    const double a = 1 / (1 + z);
    par->z_arr_r2z[i] = z;
    par->r_arr_r2z[i] = r;
    const double gz = growth_of_a(&bg, a, alim) / growth0;
    par->growth_d_arr[i] = gz;
    par->growth_v_arr[i] = gz * hubble_of_a(&bg, a) * fgrowth_of_a(&bg, a, alim) / (par->fgrowth_0 * par->hubble_0);
  }
  if (par->z_arr_r2z[GH_NZ - 1] <= par->z_max || par->r_arr_r2z[GH_NZ - 1] <= par->r_max) report_error(1, "OMG!\n");
  /* the user hooks (host/user_defined.c) on the radial grid: get_HI on the device interpolates these */
  for (int i = 0; i < GH_NZ; i++) {
    par->frac_HI_arr[i] = fraction_HI(par->z_arr_r2z[i]);
    par->bias_HI_arr[i] = bias_HI(par->z_arr_r2z[i]);
  }
  read_pk(par);
}
