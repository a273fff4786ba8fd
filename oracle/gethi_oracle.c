/* TEST INFRASTRUCTURE ONLY (oracle) -- see gethi_oracle.h for the contract and provenance. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "gethi_oracle.h"
#include "mt19937.h"
#include "fft3d.h"
#include "healpix_ring.h"

#define NH(n) ((n) / 2 + 1)

/* ------------------------------------------------------------------ table look-ups */

/* cosmo.c:153-170: uniform-grid bin index, but the node abscissa is the table's own logkarr[ik] */
double oracle_pk_linear0(const gh_cuda_params *p, double lgk)
{
  int ik = (int)((lgk - p->logkmin) * p->idlogk);
  if (ik < 0) return p->pkarr[0] * pow(10, p->n_scal * (lgk - p->logkmin));
  if (ik < p->numk) {
    /* for logkmax <= lgk < logkmax + 1/idlogk the reference has ik == numk-1 and reads pkarr[numk], one
     * past the end of its table (undefined behaviour; k >= kmax of the input file, far beyond any grid's
     * Nyquist frequency).  The oracle clamps the upper node instead. */
    double hi = (ik + 1 < p->numk) ? p->pkarr[ik + 1] : p->pkarr[ik];
    return p->pkarr[ik] + (lgk - p->logkarr[ik]) * (hi - p->pkarr[ik]) * p->idlogk;
  }
  return p->pkarr[p->numk - 1] * pow(10, -3 * (lgk - p->logkmax));
}

/* cosmo.c:40-50 */
double oracle_r_of_z(const gh_cuda_params *p, double z)
{
  if (z <= 0) return 0;
  if (z >= p->z_arr_z2r[p->nz_tab - 1]) return p->r_arr_z2r[p->nz_tab - 1];
  int iz = (int)(z / p->dz_tab);
  return p->r_arr_z2r[iz] + (p->r_arr_z2r[iz + 1] - p->r_arr_z2r[iz]) * (z - p->z_arr_z2r[iz]) / p->dz_tab;
}

static double lerp_r(const gh_cuda_params *p, const double *tab, double r, double at_zero)
{
  if (r <= 0) return at_zero;
  if (r >= p->r_arr_r2z[p->nz_tab - 1]) return tab[p->nz_tab - 1];
  int ir = (int)(r * p->glob_idr);
  return tab[ir] + (tab[ir + 1] - tab[ir]) * (r - p->r_arr_r2z[ir]) * p->glob_idr;
}
double oracle_z_of_r(const gh_cuda_params *p, double r) { return lerp_r(p, p->z_arr_r2z, r, 0); }          /* cosmo.c:52-62 */
double oracle_dgrowth_of_r(const gh_cuda_params *p, double r) { return lerp_r(p, p->growth_d_arr, r, 1); } /* cosmo.c:64-74 */
double oracle_vgrowth_of_r(const gh_cuda_params *p, double r) { return lerp_r(p, p->growth_v_arr, r, 1); } /* cosmo.c:76-86 */

/* user_defined.c:27-35 -- a file the reference tells its users to edit.  The shipped functions are
 * x_HI = a (1+z)^p and b_HI = b0 + b1 (1+z)^q; oracle_set_user_defined changes the five numbers so that the tests can
 * follow a user's edit (pinned by the reference compiled with oracle/userdef_variant.c in place of its user_defined.c). */
static double ud_a = 0.008, ud_p = 0.6, ud_b0 = 0.904, ud_b1 = 0.135, ud_q = 1.696;
void oracle_set_user_defined(double a, double p, double b0, double b1, double q) { ud_a = a; ud_p = p; ud_b0 = b0; ud_b1 = b1; ud_q = q; }
double oracle_fraction_HI(double z) { return ud_a * pow(1 + z, ud_p); }        /* user_defined.c:27-30 */
double oracle_bias_HI(double z) { return ud_b0 + ud_b1 * pow(1 + z, ud_q); }  /* user_defined.c:32-35 */

/* ------------------------------------------------------------------ k-space realisation */

static double signed_wavenumber(int i, int n, double dk)
{
  return (2 * i <= n) ? i * dk : -(n - i) * dk; /* fourier.c:263-267,271-274,280-283 */
}

/* one stored mode given its two uniforms (u1 first: phase; u2: modulus), fourier.c:285-299 + common.c:154-164 */
static void mode_from_uniforms(const gh_cuda_params *p, double k2, double idk3, double factor, double u1, double u2,
                               float _Complex *dk_out, float _Complex *vk_out)
{
  if (k2 <= 0) { *dk_out = 0; *vk_out = 0; return; }
  double lgk = 0.5 * log10(k2);
  double sigma2 = oracle_pk_linear0(p, lgk) * idk3;
  if (p->do_smoothing) sigma2 *= exp(-p->r2_smooth * k2);
  double phase = 2 * M_PI * u1;
  double mod = sqrt(-sigma2 * log(1 - u2));
  float _Complex d = (float _Complex)(mod * cexp(I * phase));
  *dk_out = d;
  *vk_out = (float _Complex)((double _Complex)d * factor / k2); /* reads back the rounded float */
}

void oracle_kgen_mt19937(const gh_cuda_params *p, oracle_slab s, int n_threads, float _Complex *dens_k,
                         float _Complex *vpot_k)
{
  const int n = p->n_grid, nh = NH(n);
  const double dk = 2 * M_PI / p->l_box, idk3 = 1. / (dk * dk * dk), factor = p->fgrowth_0 * p->hubble_0;
  /* libgomp's default static schedule: the first (nz % nthr) threads get one extra plane */
  int q = s.nz_here / n_threads, rem = s.nz_here % n_threads, start = 0;
  for (int t = 0; t < n_threads; t++) {
    int cnt = q + (t < rem ? 1 : 0);
    oracle_mt19937 g;
    oracle_mt_seed(&g, p->seed_rng + (unsigned)t); /* IThread0 = 0 without MPI, fourier.c:253 */
    for (int ii = start; ii < start + cnt; ii++) {
      double kz = signed_wavenumber(s.iz0_here + ii, n, dk);
      for (int jj = 0; jj < n; jj++) {
        double ky = signed_wavenumber(jj, n, dk);
        for (int kk = 0; kk < nh; kk++) {
          double kx = signed_wavenumber(kk, n, dk);
          double k2 = kx * kx + ky * ky + kz * kz;
          size_t idx = kk + (size_t)nh * (jj + (size_t)n * ii);
          double u1 = 0, u2 = 0;
          if (k2 > 0) { u1 = oracle_mt_uniform(&g); u2 = oracle_mt_uniform(&g); }
          mode_from_uniforms(p, k2, idk3, factor, u1, u2, &dens_k[idx], &vpot_k[idx]);
        }
      }
    }
    start += cnt;
  }
}

void oracle_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
  uint32_t c[4] = {ctr_in[0], ctr_in[1], ctr_in[2], ctr_in[3]}, k[2] = {key_in[0], key_in[1]};
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof(c));
}

/* Product stream: counter = (g>>1 lo, g>>1 hi, 0, 0) for global mode index g, key = (seed, 'GetH'); the even
 * mode of a pair uses words 0,1 and the odd one words 2,3: u1 = (w_a>>8)/2^24 (phase), u2 = (w_b>>8)/2^24 (modulus) -- 24-bit uniforms, exact in float, u2 < 1 so ln(1-u2) is finite.  The global mode index is the reference's single-process index
 * kk + nh*(jj + n*ii) (fourier.c:278), so the realisation does not depend on the number of GPUs.
 * Rows ky in [ky0, ky0+nky).  transposed_layout=0: reference layout restricted to those rows,
 * out[(kz*nky + (ky-ky0))*nh + kx] -- for nky==n this is exactly [kz][ky][kx]. */
void oracle_kgen_philox(const gh_cuda_params *p, int ky0, int nky, float _Complex *dens_k, float _Complex *vpot_k,
                        int transposed_layout)
{
  (void)transposed_layout;
  const int n = p->n_grid, nh = NH(n);
  const double dk = 2 * M_PI / p->l_box, idk3 = 1. / (dk * dk * dk), factor = p->fgrowth_0 * p->hubble_0;
  const uint32_t key[2] = {p->seed_rng, 0x47657448u};
#pragma omp parallel for schedule(static)
  for (int ii = 0; ii < n; ii++) {
    double kz = signed_wavenumber(ii, n, dk);
    for (int jl = 0; jl < nky; jl++) {
      int jj = ky0 + jl;
      double ky = signed_wavenumber(jj, n, dk);
      for (int kk = 0; kk < nh; kk++) {
        double kx = signed_wavenumber(kk, n, dk);
        double k2 = kx * kx + ky * ky + kz * kz;
        uint64_t gidx = (uint64_t)kk + (uint64_t)nh * ((uint64_t)jj + (uint64_t)n * ii);
        uint64_t blk = gidx >> 1; /* one Philox block serves the two modes of a pair */
        uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), 0, 0}, r4[4], r[2];
        oracle_philox4x32_10(ctr, key, r4);
        r[0] = r4[2 * (gidx & 1)];
        r[1] = r4[2 * (gidx & 1) + 1];
        size_t o = ((size_t)ii * nky + jl) * nh + kk;
        /* phase from 24 bits; the modulus draw keeps all 32 bits, like gsl_rng_uniform (u32 / 2^32, src/common.c:154-164):
         * a 24-bit u2 would clip the Rayleigh tail at sqrt(ln 2^24) = 4.08 sigma_k */
        mode_from_uniforms(p, k2, idk3, factor, (r[0] >> 8) / 16777216.0, r[1] / 4294967296.0, &dens_k[o], &vpot_k[o]);
      }
    }
  }
}

/* ------------------------------------------------------------------ real-space field stages */

void oracle_normalize(const gh_cuda_params *p, oracle_slab s, float *dens, float *vpot)
{
  const size_t tot = 2 * (size_t)NH(p->n_grid) * p->n_grid * s.nz_here; /* padding included, fourier.c:381 */
  const double norm = pow(sqrt(2 * M_PI) / p->l_box, 3);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < tot; i++) {
    dens[i] *= norm; /* float * double -> double -> float, fourier.c:409-410 */
    vpot[i] *= norm;
  }
}

void oracle_radial_velocity(const gh_cuda_params *p, oracle_slab s, const float *vpot, const float *slice_left,
                            const float *slice_right, float *rvel)
{
  const int n = p->n_grid, ngx = 2 * NH(n);
  const double dx = p->l_box / n, idx = 1. / dx;
#pragma omp parallel for schedule(static)
  for (int iz = 0; iz < s.nz_here; iz++) {
    double z = dx * (iz + s.iz0_here + 0.5) - p->pos_obs[2];
    size_t pz = (size_t)iz * ngx * n;
    for (int iy = 0; iy < n; iy++) {
      double y = dx * (iy + 0.5) - p->pos_obs[1];
      size_t py = (size_t)iy * ngx, py_hi = (size_t)((iy + 1) % n) * ngx, py_lo = (size_t)((iy + n - 1) % n) * ngx;
      for (int ix = 0; ix < n; ix++) {
        double x = dx * (ix + 0.5) - p->pos_obs[0];
        double irr = 1. / sqrt(x * x + y * y + z * z);
        int ix_hi = (ix + 1) % n, ix_lo = (ix + n - 1) % n;
        double vx = 0.5 * idx * (vpot[ix_hi + py + pz] - vpot[ix_lo + py + pz]); /* float difference first */
        double vy = 0.5 * idx * (vpot[ix + py_hi + pz] - vpot[ix + py_lo + pz]);
        /* z neighbours: halo planes at the slab edges (fourier.c:361-366); the lower edge test wins
         * when the slab is one plane thick, as in the reference's if / else-if chain */
        float up = (iz == s.nz_here - 1) ? slice_right[ix + py] : vpot[ix + py + pz + (size_t)ngx * n];
        float dn = (iz == 0) ? slice_left[ix + py] : vpot[ix + py + pz - (size_t)ngx * n];
        double vz = 0.5 * idx * (up - dn);
        rvel[ix + py + pz] = (float)(vx * (x * irr) + vy * (y * irr) + vz * (z * irr));
      }
    }
  }
}

void oracle_sigma_partial(const gh_cuda_params *p, oracle_slab s, const float *dens, double *mean_part,
                          double *sigma2_part)
{
  const int n = p->n_grid, ngx = 2 * NH(n);
  const double ng_tot = (double)n * ((double)n * n);
  double s1 = 0, s2 = 0;
  for (int iz = 0; iz < s.nz_here; iz++)
    for (int iy = 0; iy < n; iy++) {
      const float *row = dens + ((size_t)iz * n + iy) * ngx;
      for (int ix = 0; ix < n; ix++) {
        float sq = row[ix] * row[ix]; /* float product, then promoted (fourier.c:51) */
        s2 += sq;
        s1 += row[ix];
      }
    }
  *mean_part = s1 / ng_tot;
  *sigma2_part = s2 / ng_tot;
}

void oracle_fields_from_k(const gh_cuda_params *p, float _Complex *dens_k, float _Complex *vpot_k, float *rvel,
                          double *sigma2_gauss, double *mean_gauss)
{
  const int n = p->n_grid;
  oracle_slab s = {n, 0};
  oracle_c2r_3d_inplace(n, dens_k); /* fourier.c:391-392 */
  oracle_c2r_3d_inplace(n, vpot_k);
  float *dens = (float *)dens_k, *vpot = (float *)vpot_k;
  oracle_normalize(p, s, dens, vpot);
  size_t plane = (size_t)2 * NH(n) * n;
  oracle_radial_velocity(p, s, vpot, vpot + (size_t)(n - 1) * plane, vpot, rvel); /* fourier.c:425-427 */
  double m, s2;
  oracle_sigma_partial(p, s, dens, &m, &s2);
  *sigma2_gauss = s2 - m * m; /* fourier.c:74 */
  if (mean_gauss) *mean_gauss = m;
}

/* ------------------------------------------------------------------ get_HI */

void oracle_get_HI(const gh_cuda_params *p, oracle_slab s, double sigma2_gauss, float *dens, float *rvel)
{
  const int n = p->n_grid, ngx = 2 * NH(n);
  const double dx = p->l_box / n, mass_prefac = dx * dx * dx;
#pragma omp parallel for schedule(static)
  for (int iz = 0; iz < s.nz_here; iz++) {
    double z = dx * (iz + s.iz0_here + 0.5) - p->pos_obs[2];
    for (int iy = 0; iy < n; iy++) {
      double y = dx * (iy + 0.5) - p->pos_obs[1];
      float *drow = dens + ((size_t)iz * n + iy) * ngx, *vrow = rvel + ((size_t)iz * n + iy) * ngx;
      for (int ix = 0; ix < n; ix++) {
        double x = dx * (ix + 0.5) - p->pos_obs[0];
        double r = sqrt(x * x + y * y + z * z);
        double redshift = oracle_z_of_r(p, r);
        double gfd = oracle_dgrowth_of_r(p, r) * oracle_bias_HI(redshift);
        double delta_gauss = drow[ix];
        double dz_rsd = vrow[ix] * oracle_vgrowth_of_r(p, r);
        double dens_LN = exp(gfd * (delta_gauss - 0.5 * gfd * sigma2_gauss));
        drow[ix] = (float)(mass_prefac * oracle_fraction_HI(redshift) * dens_LN);
        vrow[ix] = (float)dz_rsd;
      }
    }
  }
}

/* ------------------------------------------------------------------ mk_T_maps */

void oracle_subparticle_offsets(const gh_cuda_params *p, double *xyz30)
{
  oracle_mt19937 g;
  oracle_mt_seed(&g, p->seed_rng);
  const double lcell = p->l_box / p->n_grid;
  for (int i = 0; i < GH_CUDA_N_SUBPART; i++) { /* interleaved x,y,z draws, pixelize.c:160-164 */
    xyz30[i] = lcell * (oracle_mt_uniform(&g) - 0.5);
    xyz30[GH_CUDA_N_SUBPART + i] = lcell * (oracle_mt_uniform(&g) - 0.5);
    xyz30[2 * GH_CUDA_N_SUBPART + i] = lcell * (oracle_mt_uniform(&g) - 0.5);
  }
}

int oracle_get_inu(const gh_cuda_params *p, double nu, int inu_start)
{
  int inu = inu_start < 0 ? 0 : (inu_start >= p->n_nu ? p->n_nu - 1 : inu_start);
  for (;;) {
    if (inu == -1 || inu == p->n_nu) return inu;
    if (nu < p->nu0_arr[inu]) inu--;
    else if (nu >= p->nuf_arr[inu]) inu++;
    else return inu;
  }
}

int oracle_shell_of_nu(const gh_cuda_params *p, double nu, int inu_prev)
{
  if (p->irregular_nutable) return oracle_get_inu(p, nu, inu_prev);
  double inv_dnu = p->n_nu / (p->nu_max - p->nu_min); /* pixelize.c:176-178 */
  return (int)(inv_dnu * (nu - p->nu_min));           /* C truncation toward zero, pixelize.c:216 */
}

static inline void point_to_shell_pixel(const gh_cuda_params *p, double x, double y, double z, double dz_rsd,
                                        int *inu_io, long *ipix)
{
  double r = sqrt(x * x + y * y + z * z);
  double redshift = oracle_z_of_r(p, r) + dz_rsd;
  double nu = GH_CUDA_NU_21 / (1 + redshift);
  int inu = oracle_shell_of_nu(p, nu, *inu_io);
  *inu_io = inu;
  *ipix = -1;
  if (inu >= 0 && inu < p->n_nu) {
    double pos[3] = {x, y, z};
    *ipix = oracle_vec2pix_ring(p->n_side, pos);
  }
}

void oracle_accumulate_maps(const gh_cuda_params *p, oracle_slab s, const float *mass, const float *dz_rsd,
                            float *maps)
{
  const int n = p->n_grid, ngx = 2 * NH(n);
  const long npix = 12 * p->n_side * p->n_side;
  const double dx = p->l_box / n;
  double off[3 * GH_CUDA_N_SUBPART];
  oracle_subparticle_offsets(p, off);
  for (int iz = 0; iz < s.nz_here; iz++) {
    int inu = 0;
    double z0 = dx * (iz + s.iz0_here + 0.5) - p->pos_obs[2];
    for (int iy = 0; iy < n; iy++) {
      double y0 = dx * (iy + 0.5) - p->pos_obs[1];
      for (int ix = 0; ix < n; ix++) {
        size_t idx = ix + ((size_t)iz * n + iy) * ngx;
        double x0 = dx * (ix + 0.5) - p->pos_obs[0];
        double mass_sub = mass[idx] / GH_CUDA_N_SUBPART; /* float -> double / int, pixelize.c:203 */
        double dz = (double)dz_rsd[idx];
        for (int isub = 0; isub < GH_CUDA_N_SUBPART; isub++) {
          long ipix;
          point_to_shell_pixel(p, x0 + off[isub], y0 + off[GH_CUDA_N_SUBPART + isub],
                               z0 + off[2 * GH_CUDA_N_SUBPART + isub], dz, &inu, &ipix);
          if (ipix >= 0) maps[ipix + npix * (size_t)inu] += mass_sub; /* float += double -> float */
        }
      }
    }
  }
}

void oracle_shell_prefactors(const gh_cuda_params *p, double *prefac)
{
  const long npix = 12 * p->n_side * p->n_side;
  const double m2t = 90.057156 * p->OmegaB * p->hhub * npix / (4 * M_PI); /* pixelize.c:155 */
  for (int inu = 0; inu < p->n_nu; inu++) {
    double dnu, nu;
    if (p->irregular_nutable) {
      dnu = p->nuf_arr[inu] - p->nu0_arr[inu];
      nu = (p->nuf_arr[inu] + p->nu0_arr[inu]) * 0.5;
    } else {
      dnu = (p->nu_max - p->nu_min) / p->n_nu;
      nu = p->nu_min + (inu + 0.5) * dnu;
    }
    double r = oracle_r_of_z(p, GH_CUDA_NU_21 / nu - 1);
    prefac[inu] = m2t / (r * r * dnu);
  }
}

void oracle_normalize_maps(const gh_cuda_params *p, float *maps)
{
  const long npix = 12 * p->n_side * p->n_side;
  double *pf = malloc(sizeof(double) * p->n_nu);
  oracle_shell_prefactors(p, pf);
#pragma omp parallel for schedule(static)
  for (int inu = 0; inu < p->n_nu; inu++)
    for (long ip = 0; ip < npix; ip++) maps[ip + npix * (size_t)inu] *= pf[inu]; /* float*double -> float */
  free(pf);
}

void oracle_points_to_shell_pixel(const gh_cuda_params *p, const double *pos, const double *dz_rsd, long long n,
                                  int *shell_out, long long *pix_out)
{
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; i++) {
    int inu = 0;
    long ipix;
    point_to_shell_pixel(p, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dz_rsd ? dz_rsd[i] : 0.0, &inu, &ipix);
    shell_out[i] = inu;
    pix_out[i] = ipix;
  }
}
