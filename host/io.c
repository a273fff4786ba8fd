/* Parameter file, frequency table, FITS / nuTable output of GetHI.
 * Behaviour follows reference src/io_gh.c:29-324 and src/healpix_extra.c:132-164; the FITS writer is a
 * from-scratch minimal BINTABLE emitter (cfitsio is not a dependency here). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "gh_host.h"

static int count_lines(FILE *f)
{
  int n = 0;
  char buf[1000];
  while (fgets(buf, sizeof(buf), f)) n++;
  return n;
}

static void read_nutable(ParamGetHI *par)
{
  FILE *f = fopen(par->fnameNuTable, "r");
  if (!f) { fprintf(stderr, "CRIME: Couldn't open file %s \n", par->fnameNuTable); exit(1); }
  par->n_nu = count_lines(f) - 1;
  rewind(f);
  if (par->n_nu < 1) report_error(1, "Error reading file %s, line %d\n", par->fnameNuTable, 1);
  par->nu0_arr = malloc(sizeof(double) * par->n_nu);
  par->nuf_arr = malloc(sizeof(double) * par->n_nu);
  for (int i = 0; i <= par->n_nu; i++) {
    double nu;
    if (fscanf(f, "%lf ", &nu) != 1) report_error(1, "Error reading file %s, line %d\n", par->fnameNuTable, i + 1);
    if (i != par->n_nu) par->nu0_arr[i] = nu;
    if (i != 0) par->nuf_arr[i - 1] = nu;
  }
  fclose(f);
  for (int i = 0; i < par->n_nu; i++)
    if (par->nuf_arr[i] <= par->nu0_arr[i]) report_error(1, "Frequency bins don't make sense\n");
  par->nu_max = par->nuf_arr[par->n_nu - 1];
  par->nu_min = par->nu0_arr[0];
}

static ParamGetHI *param_gethi_new(void)
{
  ParamGetHI *p = calloc(1, sizeof(ParamGetHI));
  if (!p) { fprintf(stderr, "out of memory\n"); exit(1); }
  /* defaults of src/io_gh.c:133-186 (do_psources has none there; 0 here) */
  strcpy(p->fnamePk, "default");
  p->OmegaM = 0.3; p->OmegaL = 0.7; p->OmegaB = 0.05; p->hhub = 0.7; p->weos = -1.; p->n_scal = 0.96; p->sig8 = 0.83;
  p->fgrowth_0 = -1; p->hubble_0 = -1; p->z_max = 1.5; p->z_min = 0.5; p->r_max = -1; p->r_min = -1;
  p->r2_smooth = 2.0; p->do_smoothing = 1;
  p->logkmax = 1; p->logkmin = -3; p->idlogk = 100; p->glob_idr = -1;
  p->seed_rng = 1234; p->n_side = 128; p->nu_max = 1050.; p->nu_min = 350.; p->n_nu = 150;
  strcpy(p->fnameNuTable, "default");
  p->irregular_nutable = 1; /* the shipped Makefile defines _IRREGULAR_NUTABLE */
  p->n_grid = 512; p->l_box = -1; p->nz_here = 512;
  strcpy(p->prefixOut, "default");
  p->sigma2_gauss = -1;
  return p;
}

ParamGetHI *read_run_params_ex(const char *fname, int with_device)
{
  ParamGetHI *par = param_gethi_new();
  int have_nutable = 0;
  print_info("*** Reading run parameters \n");
  FILE *fi = fopen(fname, "r");
  if (!fi) { fprintf(stderr, "CRIME: Couldn't open file %s \n", fname); exit(1); }
  const int n_lin = count_lines(fi);
  rewind(fi);
  double nu_min_key = par->nu_min, nu_max_key = par->nu_max;
  int n_nu_key = par->n_nu;
  for (int ii = 0; ii < n_lin; ii++) {
    char s0[512], s1[64], s2[256];
    if (!fgets(s0, sizeof(s0), fi)) { fprintf(stderr, "CRIME: Error reading file %s, line %d \n", fname, ii + 1); exit(1); }
    if (s0[0] == '#' || s0[0] == '\n') continue;
    if (sscanf(s0, "%63s %255s", s1, s2) != 2) { fprintf(stderr, "CRIME: Error reading file %s, line %d \n", fname, ii + 1); exit(1); }
    if (!strcmp(s1, "prefix_out=")) snprintf(par->prefixOut, sizeof(par->prefixOut), "%s", s2);
    else if (!strcmp(s1, "pk_filename=")) snprintf(par->fnamePk, sizeof(par->fnamePk), "%s", s2);
    else if (!strcmp(s1, "omega_M=")) par->OmegaM = atof(s2);
    else if (!strcmp(s1, "omega_L=")) par->OmegaL = atof(s2);
    else if (!strcmp(s1, "omega_B=")) par->OmegaB = atof(s2);
    else if (!strcmp(s1, "h=")) par->hhub = atof(s2);
    else if (!strcmp(s1, "w=")) par->weos = atof(s2);
    else if (!strcmp(s1, "ns=")) par->n_scal = atof(s2);
    else if (!strcmp(s1, "sigma_8=")) par->sig8 = atof(s2);
    else if (!strcmp(s1, "r_smooth=")) par->r2_smooth = atof(s2);
    else if (!strcmp(s1, "frequencies_filename=")) { snprintf(par->fnameNuTable, sizeof(par->fnameNuTable), "%s", s2); have_nutable = 1; }
    else if (!strcmp(s1, "nu_min=")) nu_min_key = atof(s2);
    else if (!strcmp(s1, "nu_max=")) nu_max_key = atof(s2);
    else if (!strcmp(s1, "n_nu=")) n_nu_key = atoi(s2);
    else if (!strcmp(s1, "n_grid=")) par->n_grid = atoi(s2);
    else if (!strcmp(s1, "n_side=")) par->n_side = atoi(s2);
    else if (!strcmp(s1, "seed=")) par->seed_rng = atoi(s2);
    else if (!strcmp(s1, "do_psources=")) par->do_psources = atoi(s2);
    else fprintf(stderr, "CRIME: Unknown parameter %s\n", s1);
  }
  fclose(fi);
  if (par->r2_smooth > 0) { par->r2_smooth = pow(par->r2_smooth, 2); par->do_smoothing = 1; }
  else par->do_smoothing = 0;
  /* The reference picks the frequency-table personality at compile time (-D_IRREGULAR_NUTABLE, on in the
   * shipped Makefile).  Here: a frequencies_filename selects it, otherwise nu_min / nu_max / n_nu. */
  par->irregular_nutable = have_nutable;
  if (have_nutable) read_nutable(par);
  else { par->nu_min = nu_min_key; par->nu_max = nu_max_key; par->n_nu = n_nu_key; }
  gh_phase("parameter file");
  cosmo_set(par);
  gh_phase("cosmo_set");
  if (with_device) init_fftw(par);
  gh_phase("init_fftw (CUDA context, device and pinned allocations, NCCL)");

  const double dk = 2 * M_PI / par->l_box;
  const double dtheta = sqrt(41253. / (12 * par->n_side * par->n_side));
  const double rtod = 57.2957795;
  print_info("Run parameters: \n");
  print_info("  %.3lf < nu/MHz < %.3lf\n", par->nu_min, par->nu_max);
  print_info("  %.3lf < z < %.3lf\n", par->z_min, par->z_max);
  print_info("  %.3lf < r/(Mpc/h) < %.3lf\n", par->r_min, par->r_max);
  print_info("  L_box = %.3lf Mpc/h, N_grid = %d \n", par->l_box, par->n_grid);
  print_info("  Scales resolved: %.3lE < k < %.3lE h/Mpc\n", dk, 0.5 * (par->n_grid - 1) * dk);
  print_info("  Fourier-space resolution: dk = %.3lE h/Mpc\n", dk);
  print_info("  Real-space resolution: dx = %.3lE Mpc/h\n", par->l_box / par->n_grid);
  if (par->do_smoothing) print_info("  Density field pre-smoothed on scales: x_s = %.3lE Mpc/h\n", sqrt(par->r2_smooth));
  else print_info("  No extra smoothing\n");
  print_info("  n_nu = %d, d_nu= %.3lf MHz, drL ~ %.3lf Mpc/h\n", par->n_nu, (par->nu_max - par->nu_min) / par->n_nu,
             (par->r_max - par->r_min) / par->n_nu);
  print_info("  n_side = %ld, dtheta = %.3lf deg, %.3lf < drT/(Mpc/h) < %.3lf\n", par->n_side, dtheta, par->r_min * dtheta / rtod,
             par->r_max * dtheta / rtod);
  print_info("  Estimated output size ~ %.1lf GB \n", 12.0 * par->n_side * par->n_side * par->n_nu * sizeof(float) / (1024. * 1024 * 1024));
  print_info("\n");
  return par;
}

ParamGetHI *read_run_params(const char *fname) { return read_run_params_ex(fname, 1); }

/* ---- FITS ---- */
static void put_card(FILE *fp, const char *key, const char *val, const char *comm, int is_str, long *nbytes)
{
  char body[200], line[96];
  if (is_str) {
    char q[80];
    snprintf(q, sizeof(q), "'%-8s'", val);
    snprintf(body, sizeof(body), "%-8.8s= %-20s / %s", key, q, comm);
  } else {
    snprintf(body, sizeof(body), "%-8.8s= %20s / %s", key, val, comm);
  }
  snprintf(line, sizeof(line), "%-80.80s", body);
  fwrite(line, 1, 80, fp);
  *nbytes += 80;
}

static void end_header(FILE *fp, long *nbytes)
{
  char line[96];
  snprintf(line, sizeof(line), "%-80s", "END");
  fwrite(line, 1, 80, fp);
  *nbytes += 80;
  while (*nbytes % 2880) { fputc(' ', fp); (*nbytes)++; }
}

int gh_write_healpix_map(const float *map, long nside, const char *fname)
{
  /* cfitsio's fits_create_file refuses to overwrite and the reference ignores the status, silently
   * writing nothing on a re-run (src/healpix_extra.c:145).  Here an existing file is an error the caller sees. */
  FILE *t = fopen(fname, "rb");
  if (t) { fclose(t); return 1; }
  FILE *fp = fopen(fname, "wb");
  if (!fp) return 2;
  const long npix = 12 * nside * nside;
  long nb = 0;
  char v[32];
  put_card(fp, "SIMPLE", "T", "file does conform to FITS standard", 0, &nb);
  put_card(fp, "BITPIX", "8", "number of bits per data pixel", 0, &nb);
  put_card(fp, "NAXIS", "0", "number of data axes", 0, &nb);
  put_card(fp, "EXTEND", "T", "FITS dataset may contain extensions", 0, &nb);
  end_header(fp, &nb);
  nb = 0;
  put_card(fp, "XTENSION", "BINTABLE", "binary table extension", 1, &nb);
  put_card(fp, "BITPIX", "8", "8-bit bytes", 0, &nb);
  put_card(fp, "NAXIS", "2", "2-dimensional binary table", 0, &nb);
  put_card(fp, "NAXIS1", "4", "width of table in bytes", 0, &nb);
  snprintf(v, sizeof(v), "%ld", npix);
  put_card(fp, "NAXIS2", v, "number of rows in table", 0, &nb);
  put_card(fp, "PCOUNT", "0", "size of special data area", 0, &nb);
  put_card(fp, "GCOUNT", "1", "one data group (required keyword)", 0, &nb);
  put_card(fp, "TFIELDS", "1", "number of fields in each row", 0, &nb);
  put_card(fp, "TTYPE1", "T", "label for field   1", 1, &nb);
  put_card(fp, "TFORM1", "1E", "data format of field: 4-byte REAL", 1, &nb);
  put_card(fp, "TUNIT1", "mK", "physical unit of field", 1, &nb);
  put_card(fp, "EXTNAME", "BINTABLE", "name of this binary table extension", 1, &nb);
  put_card(fp, "PIXTYPE", "HEALPIX", "HEALPIX Pixelisation", 1, &nb);
  put_card(fp, "ORDERING", "RING", "Pixel ordering scheme, either RING or NESTED", 1, &nb);
  snprintf(v, sizeof(v), "%ld", nside);
  put_card(fp, "NSIDE", v, "Resolution parameter for HEALPIX", 0, &nb);
  put_card(fp, "COORDSYS", "G", "Pixelisation coordinate system", 1, &nb);
  {
    char line[96];
    snprintf(line, sizeof(line), "%-80.80s", "COMMENT G = Galactic, E = ecliptic, C = celestial = equatorial");
    fwrite(line, 1, 80, fp);
    nb += 80;
  }
  end_header(fp, &nb);
  /* big-endian float32 rows: 32-bit byte swaps (the compiler turns the loop into vector shuffles), 1 MiB per write */
  enum { CH = 1 << 18 };
  uint32_t *buf = malloc(sizeof(uint32_t) * CH);
  if (!buf) { fclose(fp); return 3; }
  long written = 0;
  for (long i0 = 0; i0 < npix; i0 += CH) {
    const long n = npix - i0 < CH ? npix - i0 : CH;
    const uint32_t *src = (const uint32_t *)(const void *)(map + i0);
    for (long i = 0; i < n; i++) buf[i] = __builtin_bswap32(src[i]);
    if (fwrite(buf, 4, n, fp) != (size_t)n) { free(buf); fclose(fp); return 4; }
    written += 4 * n;
  }
  free(buf);
  while (written % 2880) { fputc(0, fp); written++; }
  return fclose(fp) ? 4 : 0;
}

static void write_nu_table(const ParamGetHI *par)
{
  char fn[300];
  snprintf(fn, sizeof(fn), "%s_nuTable.dat", par->prefixOut);
  FILE *f = fopen(fn, "w");
  if (!f) { fprintf(stderr, "CRIME: Couldn't open file %s \n", fn); exit(1); }
  for (int i = 0; i < par->n_nu; i++) {
    double nu0, nuf;
    if (par->irregular_nutable) { nu0 = par->nu0_arr[i]; nuf = par->nuf_arr[i]; }
    else {
      nu0 = par->nu_min + (par->nu_max - par->nu_min) * (i + 0.0) / par->n_nu;
      nuf = par->nu_min + (par->nu_max - par->nu_min) * (i + 1.0) / par->n_nu;
    }
    fprintf(f, "%d %lf %lf %lf %lf\n", i + 1, nu0, nuf, GH_NU_21 / nuf - 1, GH_NU_21 / nu0 - 1);
  }
  fclose(f);
}

/* Every rank writes the shells it owns (the reference gathers everything on rank 0 first and writes the files
 * one after the other, with a float copy per map, src/io_gh.c:69-131).  One file per shell, so the shells are
 * written concurrently: the byte-swap runs on several cores and the file system sees several streams.
 * After mk_T_maps_begin the writers also overlap the device->host copy: shell s is written as soon as the
 * chunk holding it has landed, while later shells are still on the wire (and, on several GPUs, while other
 * ranks are still reducing). */
void write_maps(ParamGetHI *par)
{
  if (par->rank == 0) write_nu_table(par); /* the shipped Makefile's -D_DEBUG output, src/io_gh.c:112-114 */
  print_info("*** Writing files %s_###.fits\n", par->prefixOut);
  const long npix = 12 * par->n_side * par->n_side;
  int n_exist = 0, n_fail = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : n_exist, n_fail)
  for (int s = 0; s < par->n_shells_here; s++) {
    char fn[300];
    snprintf(fn, sizeof(fn), "%s_%03d.fits", par->prefixOut, par->shell0_here + s + 1);
    if (par->maps_streaming && gh_cuda_wait_shells(par->cuda, s + 1)) { n_fail++; continue; }
    const int rc = gh_write_healpix_map(par->maps_HI + (size_t)s * npix, par->n_side, fn);
    if (rc == 1) n_exist++;
    else if (rc) n_fail++;
  }
  if (par->maps_streaming) {
    if (gh_cuda_wait(par->cuda, NULL)) report_error(1, "mk_T_maps: %s\n", gh_cuda_last_error());
    double ms[GH_T_NSLOTS];
    if (!gh_cuda_stage_times(par->cuda, ms))
      print_info(">    Relative time ellapsed %.1lf ms (maps; the download and the writers overlapped it)\n\n",
                 ms[GH_T_MAPS] + ms[GH_T_REDUCE] + ms[GH_T_D2H]);
    par->maps_streaming = 0;
  }
  if (par->do_psources && par->maps_PS) { /* src/io_gh.c:122-128 */
    print_info("*** Writing single slice files %s_ps_###.fits\n", par->prefixOut);
#pragma omp parallel for schedule(dynamic) reduction(+ : n_exist, n_fail)
    for (int s = 0; s < par->n_shells_here; s++) {
      char fn[300];
      snprintf(fn, sizeof(fn), "%s_ps_%03d.fits", par->prefixOut, par->shell0_here + s + 1);
      const int rc = gh_write_healpix_map(par->maps_PS + (size_t)s * npix, par->n_side, fn);
      if (rc == 1) n_exist++;
      else if (rc) n_fail++;
    }
  }
  if (n_exist) report_error(0, "%d of the %s_###.fits files exist and were left untouched\n", n_exist, par->prefixOut);
  if (n_fail) report_error(1, "could not write %d of the %s_###.fits files\n", n_fail, par->prefixOut);
}

void param_gethi_free(ParamGetHI *par)
{
  if (!par) return;
  if (par->maps_PS) gh_cuda_host_free(par->maps_PS);
  end_fftw(par);
  free(par->logkarr); free(par->pkarr); free(par->nu0_arr); free(par->nuf_arr); free(par->ps_lcdf); free(par->ps_sed);
  free(par);
}
