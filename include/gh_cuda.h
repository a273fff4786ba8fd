/*
 * gh_cuda.h -- C-ABI of the B200-native GetHI hot path (libgh_cuda.so).
 *
 * The reference (damonge/CRIME) has no plugin/FFI layer: its hot path sits behind five plain C
 * functions that take the single state struct `ParamGetHI *` and are called once each, in fixed
 * order, from main() (reference src/main_gh.c:24-80):
 *
 *     init_fftw(par)                 src/fourier.c:101      (via read_run_params, src/io_gh.c:265)
 *     create_d_and_vr_fields(par)    src/fourier.c:375      src/main_gh.c:52
 *     get_HI(par)                    src/grid_tools.c:103   src/main_gh.c:59
 *     mk_T_maps(par)                 src/pixelize.c:150     src/main_gh.c:62
 *     end_fftw()                     src/fourier.c:201      (via param_gethi_free, src/io_gh.c:323)
 *
 * The entry points below are what a reference-side binding replaces those five calls with.  Plain
 * pointers and sizes only; every call returns 0 on success, non-zero on error, and the message is
 * available from gh_cuda_last_error() (the reference's own convention is message + exit(1),
 * src/common_gh.c:104-120 -- the host wrapper turns non-zero into report_error(1, ...)).
 *
 * Process model: ONE PROCESS PER GPU.  Rank r of n owns z-planes [r*N/n, (r+1)*N/n) of the three
 * real-space grids, exactly the reference's FFTW-MPI slab decomposition (src/fourier.c:141-148), and
 * shells [r*ceil(n_nu/n), ...) of the reduced map stack.  The ranks are bound together by an NCCL
 * unique id that the launcher (torchrun + torch.distributed, or MPI_Bcast, or a pipe) distributes.
 * There is no CPU fallback: every entry point fails loudly without a CUDA device.
 */
#ifndef GH_CUDA_H
#define GH_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define GH_CUDA_UNIQUE_ID_BYTES 128
#define GH_CUDA_N_SUBPART 10 /* src/pixelize.c:25 */
#define GH_CUDA_NU_21 1420.40575177 /* src/common.h:38 */

/* POD mirror of the fields of ParamGetHI (src/common_gh.h:138-209) that the hot path reads.  All
 * pointers are HOST pointers to small tables filled by read_run_params -> cosmo_set
 * (src/io_gh.c:188-296, src/cosmo.c:341-413); they are copied to the device by gh_cuda_create. */
typedef struct gh_cuda_params {
  /* grid and geometry (src/cosmo.c:361-364).  n_grid: even, 8..4096; powers of two 32..4096 run the tuned FFT kernels and may be
   * split over a power-of-two number of ranks, other even sizes run general-length passes on one rank (the reference: any
   * size FFTW takes, src/fourier.c:85) */
  int n_grid;
  double l_box;
  double pos_obs[3];
  /* k-space realisation (src/fourier.c:234-305) */
  unsigned int seed_rng;
  int do_smoothing;
  double r2_smooth; /* already squared, src/io_gh.c:255-260 */
  double fgrowth_0;
  double hubble_0;
  /* P(k) table (src/cosmo.c:153-170, 306-339) */
  int numk;
  double logkmin, logkmax, idlogk, n_scal;
  const double *logkarr; /* [numk] */
  const double *pkarr;   /* [numk] */
  /* radial tables (src/cosmo.c:40-86, 366-393); nz_tab is the reference's NZ = 5001 */
  int nz_tab;
  double glob_idr;
  const double *z_arr_r2z;    /* [nz_tab] */
  const double *r_arr_r2z;    /* [nz_tab] */
  const double *growth_d_arr; /* [nz_tab] */
  const double *growth_v_arr; /* [nz_tab] */
  const double *z_arr_z2r;    /* [nz_tab], uniform in z with step dz_tab */
  const double *r_arr_z2r;    /* [nz_tab] */
  double dz_tab;              /* the reference's DZ = 0.001 */
  /* sky and frequency shells (src/pixelize.c:150-262, src/io_gh.c:29-57) */
  long n_side;
  int n_nu;
  int irregular_nutable; /* 1: -D_IRREGULAR_NUTABLE personality (nu0_arr/nuf_arr); 0: uniform nu_min..nu_max */
  const double *nu0_arr; /* [n_nu] lower edges (irregular only) */
  const double *nuf_arr; /* [n_nu] upper edges (irregular only) */
  double nu_min, nu_max;
  double OmegaB, hhub;
  /* the user hooks of src/user_defined.c:27-35 (a file GetHI users are told to edit), tabulated by the host on the
   * radial grid: frac_HI_arr[i] = fraction_HI(z_arr_r2z[i]), bias_HI_arr[i] = bias_HI(z_arr_r2z[i]).  get_HI
   * interpolates them in r exactly like the redshift and growth tables (src/cosmo.c:52-86).  NULL (both): the
   * formulas the reference ships, x_HI = 0.008 (1+z)^0.6 and b_HI = 0.904 + 0.135 (1+z)^1.696. */
  const double *frac_HI_arr; /* [nz_tab] or NULL */
  const double *bias_HI_arr; /* [nz_tab] or NULL */
} gh_cuda_params;

typedef struct gh_cuda_ctx gh_cuda_ctx;

/* grids addressable through the test/diagnostic entry points */
enum { GH_GRID_DENS = 0, GH_GRID_VPOT = 1, GH_GRID_RVEL = 2 };

/* stage timer slots (milliseconds, CUDA events on the context's stream); they line up one-to-one
 * with the reference's timer(2) brackets (src/fourier.c:385-433, src/grid_tools.c:112,151,
 * src/pixelize.c:168,263) */
enum {
  GH_T_KGEN = 0,    /* create_density_and_velpot_fourier */
  GH_T_FFT = 1,     /* both c2r transforms incl. transposes and the fused normalisation */
  GH_T_VEL = 2,     /* halo exchange + radial velocity */
  GH_T_SIGMA = 3,   /* compute_sigma_dens */
  GH_T_GETHI = 4,   /* get_HI */
  GH_T_MAPS = 5,    /* mk_T_maps accumulate + scale */
  GH_T_REDUCE = 6,  /* cross-GPU map reduction */
  GH_T_D2H = 7,     /* maps device -> host */
  GH_T_NSLOTS = 8
};

/* Rank 0 calls this and the launcher hands the 128 bytes to every rank (only needed for nranks>1). */
int gh_cuda_get_unique_id(void *id_out);

/* == mpi_init + init_fftw + allocate_maps (src/common_gh.c:31, src/fourier.c:101, src/io_gh.c:60).
 * Selects `device`, copies the tables, fixes this rank's slab (nz_here = N/nranks,
 * iz0_here = rank*nz_here; N % nranks must be 0), allocates the three grids, the map stack and,
 * for nranks>1, joins the NCCL communicator.  unique_id may be NULL when nranks==1. */
int gh_cuda_create(const gh_cuda_params *params, int rank, int nranks, const void *unique_id, int device,
                   gh_cuda_ctx **ctx_out);

/* Re-send the run parameters and small tables of an existing context from host memory (same n_grid,
 * n_side, n_nu, table lengths): what read_run_params hands over before each realisation.  Lets a caller
 * loop over seeds / spectra without re-allocating the grids. */
int gh_cuda_set_params(gh_cuda_ctx *ctx, const gh_cuda_params *params);

/* == end_fftw + the grid/map part of param_gethi_free (src/fourier.c:201, src/io_gh.c:298-324). */
int gh_cuda_destroy(gh_cuda_ctx *ctx);

/* slab owned by this rank (ParamGetHI.nz_here / iz0_here) and the shells it owns after the reduction */
int gh_cuda_slab(const gh_cuda_ctx *ctx, int *nz_here, int *iz0_here);
int gh_cuda_shells(const gh_cuda_ctx *ctx, int *n_shells_here, int *shell0_here);

/* Plane ranges of the map accumulation on nranks GPUs: rank r accumulates planes [bounds[r], bounds[r+1])
 * (bounds has nranks+1 entries).  The reference gives every rank its own slab (src/pixelize.c:186-232); its
 * cells' cost is uneven (cells outside the shells' radial window are skipped), so the ranges are chosen for equal
 * modelled cost and a rank pulls planes outside its slab from the owner over NVLink.  Host-only, no GPU needed. */
int gh_cuda_map_plane_bounds(const gh_cuda_params *params, int nranks, int *bounds);

/* == create_d_and_vr_fields (src/fourier.c:375-438): k-space realisation, two c2r FFTs, normalisation,
 * halo exchange, radial velocity, Gaussian variance.  *sigma2_gauss_out receives par->sigma2_gauss and
 * *mean_gauss_out (may be NULL) the <d> of the reference's log line (src/fourier.c:75).
 * If a k-space field was supplied with gh_cuda_set_delta_k it is used instead of the generator. */
int gh_cuda_create_d_and_vr_fields(gh_cuda_ctx *ctx, double *sigma2_gauss_out, double *mean_gauss_out);

/* == get_HI (src/grid_tools.c:103-153), in place on the device grids; uses the sigma2_gauss computed
 * above unless overridden by gh_cuda_set_sigma2_gauss. */
int gh_cuda_get_HI(gh_cuda_ctx *ctx);

/* == mk_T_maps (src/pixelize.c:150-286): accumulate, scale to mK, reduce across ranks.  On return
 * maps_host (may be NULL to leave the result on the device) holds this rank's shells,
 * [n_shells_here][12*n_side^2] floats, RING order; with nranks==1 that is the reference's full
 * par->maps_HI.  maps_host should be page-locked for full PCIe speed (gh_cuda_host_alloc). */
int gh_cuda_mk_T_maps(gh_cuda_ctx *ctx, float *maps_host);

/* Streaming form of mk_T_maps for an overlapped writer (the reference writes the files only after
 * everything has been gathered on rank 0, src/io_gh.c:109-131): _begin enqueues accumulation, reduction,
 * scaling and the download of this rank's shells in chunks of whole shells and returns at once;
 * gh_cuda_wait_shells(ctx, n) returns once this rank's first n shells (n < 0: all) of the most recently begun
 * download are complete in maps_host.  maps_host should come from gh_cuda_host_alloc.  Thread-safe for
 * several waiting threads. */
int gh_cuda_mk_T_maps_begin(gh_cuda_ctx *ctx, float *maps_host);
int gh_cuda_wait_shells(gh_cuda_ctx *ctx, int n_shells);

/* whole hot path, main_gh.c:52-62: the three calls above back to back */
int gh_cuda_run(gh_cuda_ctx *ctx, double *sigma2_gauss_out, float *maps_host);

/* The same without any host synchronisation: everything is enqueued (the variance stays on the device between
 * the stages) and the device->host copy of the maps runs on a second stream, so a caller producing many
 * realisations overlaps the copy of one with the computation of the next:
 *     gh_cuda_run_async(ctx, bufA); gh_cuda_run_async(ctx, bufB); gh_cuda_wait(ctx, &s2); ...
 * gh_cuda_wait returns when everything enqueued so far, copies included, has finished.  The next realisation
 * does not touch the device map stack before the pending copy has read it.  maps_host must be page-locked. */
int gh_cuda_run_async(gh_cuda_ctx *ctx, float *maps_host);
int gh_cuda_wait(gh_cuda_ctx *ctx, double *sigma2_gauss_out);

/* page-locked host memory for maps / injected fields */
int gh_cuda_host_alloc(void **ptr, unsigned long long bytes);
int gh_cuda_host_free(void *ptr);

/* ---- finer-grained stages (what create_d_and_vr_fields is made of), for parity tests and profiling ---- */
int gh_cuda_generate_k(gh_cuda_ctx *ctx);       /* src/fourier.c:234-305 with the counter-based stream */
int gh_cuda_fft_fields(gh_cuda_ctx *ctx);       /* src/fourier.c:390-413 */
int gh_cuda_radial_velocity(gh_cuda_ctx *ctx);  /* src/fourier.c:415-433 */
int gh_cuda_sigma_dens(gh_cuda_ctx *ctx, double *sigma2_gauss_out, double *mean_gauss_out); /* src/fourier.c:24-76 */
int gh_cuda_accumulate_maps(gh_cuda_ctx *ctx);  /* src/pixelize.c:186-231, un-normalised mass maps on the device */
int gh_cuda_synchronize(gh_cuda_ctx *ctx);

/* ---- injection / read-back (test and integration aids; host pointers) ---- */
/* Full half-spectra in the reference's layout [kz][ky][kx<=N/2] complex-float (src/fourier.c:278), as
 * FFTW sees them at src/fourier.c:391-392; every rank passes the same global arrays and keeps its part. */
int gh_cuda_set_delta_k(gh_cuda_ctx *ctx, const float *dens_k, const float *vpot_k);
int gh_cuda_clear_delta_k(gh_cuda_ctx *ctx);
/* k-space as generated, gathered back into the reference's global layout (this rank's ky rows only are
 * written; other elements of the output are left untouched) */
int gh_cuda_download_delta_k(gh_cuda_ctx *ctx, float *dens_k, float *vpot_k);
/* this rank's slab of a real-space grid, padded layout [nz_here][N][2(N/2+1)] floats */
int gh_cuda_download_grid(gh_cuda_ctx *ctx, int which, float *slab_out);
int gh_cuda_upload_grid(gh_cuda_ctx *ctx, int which, const float *slab_in);
int gh_cuda_set_sigma2_gauss(gh_cuda_ctx *ctx, double sigma2_gauss);
/* Position-weighted checksum of the real cells of planes [z0_local, z0_local+n_planes) of this rank's slab:
 * sum of bits(value) * (2*g + 1) mod 2^64, g = (z_global*N + y)*N + x.  Bit-identical fields give identical sums
 * whatever the slab decomposition: lets a test compare a 2048^3 run on eight GPUs with one GPU's plane by plane
 * without moving the grids (the reference has no counterpart; its fields are checked by eye). */
int gh_cuda_grid_checksum(gh_cuda_ctx *ctx, int which, int z0_local, int n_planes, unsigned long long *sum_out);
/* un-normalised or final map stack as it sits on this rank's device: [n_nu][npix] before the
 * reduction (full stack), use n_floats to bound the copy */
int gh_cuda_download_maps(gh_cuda_ctx *ctx, float *maps_out, unsigned long long first, unsigned long long n_floats);
int gh_cuda_zero_maps(gh_cuda_ctx *ctx);
/* the 3*GH_CUDA_N_SUBPART sub-particle offsets mk_T_maps draws from MT19937(seed_rng)
 * (src/pixelize.c:157-164), as x[10], y[10], z[10] */
int gh_cuda_subparticle_offsets(const gh_cuda_ctx *ctx, double *xyz_out);
/* (shell, RING pixel) of arbitrary points through the same device code mk_T_maps uses: pos is [n][3]
 * observer-centred coordinates, dz_rsd [n]; shell_out[i] is -1 / n_nu when out of range */
int gh_cuda_points_to_shell_pixel(gh_cuda_ctx *ctx, const double *pos, const double *dz_rsd, long long n,
                                  int *shell_out, long long *pix_out);

/* Audit of mk_T_maps' fp32 fast path against its exact fp64 path on arbitrary points (same inputs as
 * above).  eps_scale multiplies the fast path's error bounds (1 = production).  counts_out[4] = fast path
 * says out-of-range / in-range / unsure (-> exact path) / was sure but disagrees with the exact path; the
 * last one must be 0 at eps_scale 1, and how far eps_scale can be lowered before it is not measures the
 * head-room of the bounds. */
int gh_cuda_fastpath_audit(gh_cuda_ctx *ctx, const double *pos, const double *dz_rsd, long long n, double eps_scale,
                           unsigned long long *counts_out);

/* The same audit over every sub-particle of the grids currently on the device (HI mass in GH_GRID_DENS,
 * Delta z_RSD in GH_GRID_RVEL), through exactly the per-cell code mk_T_maps runs; nothing is deposited. */
int gh_cuda_accumulate_audit(gh_cuda_ctx *ctx, double eps_scale, unsigned long long *counts_out);

/* per-stage device times of the most recent calls, GH_T_NSLOTS doubles in ms */
int gh_cuda_stage_times(gh_cuda_ctx *ctx, double *ms_out);
/* Opt-in timers (GH_TIME_FFT_PASSES=1 in the environment when the context is created): device milliseconds of
 * each field's z pass, which on several GPUs includes the transpose fused into it (peer stores over NVLink) and
 * the barrier that closes it -- the interval the NVLink fraction of the roofline is computed from.
 * z_ms[0] = density, z_ms[1] = velocity potential; -1 when off. */
int gh_cuda_fft_pass_times(gh_cuda_ctx *ctx, double *z_ms);

/* ---- point sources on the same grids (do_psources = 1) ----
 * What setup_psources (src/psources.c:98-131) leaves in ParamGetHI, plus host-side tabulations of the three
 * user-definable functions of src/psources.c:33-71 (l_z_function through its cumulative distribution, spec_ed,
 * bias_psources), so that editing them stays a host-side change.  All pointers are HOST pointers. */
typedef struct gh_cuda_psources_params {
  int nz;                 /* NZ_PSOURCES: redshift bins of width z_max / nz (src/psources.c:106) */
  double z_max;           /* par->z_max */
  const double *nz_arr;   /* [nz] nz_psources_arr: sources per (Mpc/h)^3 integrated over luminosities */
  const double *bias_arr; /* [nz] bias_psources(z_i) */
  int nl;                 /* luminosity bins of the tabulated distribution */
  double logl_min, logl_max;  /* LLOGMIN, LLOGMAX (src/psources.c:25-26), log10 of L / 10^22 W/Hz */
  const double *lcdf;     /* [nz][nl + 1] cumulative P(log10 L | z_i): 0 at logl_min, 1 at logl_max */
  int nsed;               /* spec_ed on nsed points uniform in log10(nu / MHz) over [lognu_min, lognu_max] */
  double lognu_min, lognu_max;
  const double *sed_arr;  /* [nsed] */
  double hhub;
} gh_cuda_psources_params;
/* get_point_sources (src/grid_tools.c:24-101): Poisson-samples the number of sources of every cell from the Gaussian
 * density field -- call it between gh_cuda_create_d_and_vr_fields and gh_cuda_get_HI, as main_gh.c:55-59 does.
 * np_total_out: sources in the whole box (all ranks). */
int gh_cuda_get_point_sources(gh_cuda_ctx *ctx, const gh_cuda_psources_params *ps, long long *np_total_out);
/* mk_psources_maps (src/pixelize.c:58-148): call after gh_cuda_get_HI (it reads Delta z_RSD).  maps_ps_host receives
 * this rank's shells [n_shells_here][npix] of maps_PS in mK (may be NULL). */
int gh_cuda_mk_psources_maps(gh_cuda_ctx *ctx, float *maps_ps_host);
/* test aid: this rank's slab of source counts [nz_here][N][N] and of the Poisson means they were drawn from */
int gh_cuda_download_point_sources(gh_cuda_ctx *ctx, int *nsources_out, float *lambda_out);

/* ---- JoinT ingestion of GetHI output on the device ----
 * gh_cuda_jt_merge_maps replaces merge_maps (src/main_jt.c:98-211) for the shells this rank owns: for every shell,
 * map_in = 0, then the n_comp components are added in the order given (the reference's order: cosmological signal,
 * extragalactic free-free, galactic free-free, point sources, [polarised synchrotron * polarization_leakage,]
 * synchrotron, custom), then he_udgrade(map_in, n_side -> nside_out, RING) (src/healpix_extra.c:318-385) and the
 * result goes to out_host[n_shells_here][12 nside_out^2].  comp_host[k] is a HOST stack [n_shells_here][12 n_side^2]
 * of this rank's shells, or NULL for the stack gh_cuda_mk_T_maps / gh_cuda_run has just left on the device -- the
 * cosmological signal enters the sum without a FITS round trip.  scale[k] (may be NULL = all 1) multiplies
 * component k as `map_read[ii] *= leakage` does (float * double -> float, src/main_jt.c:183).
 * Bit-identical to the reference: float adds in the same order, double sum over the NEST children in child order.
 * gh_cuda_udgrade is he_udgrade alone on n_maps host maps (nest = 0: RING ordering). */
int gh_cuda_jt_merge_maps(gh_cuda_ctx *ctx, int n_comp, const float *const *comp_host, const double *scale, long nside_out,
                          float *out_host);
int gh_cuda_udgrade(gh_cuda_ctx *ctx, const float *maps_in, long nside_in, float *maps_out, long nside_out, int nest, int n_maps);
/* chealpix nest2ring (to_ring != 0) / ring2nest through the device code the two calls above use */
int gh_cuda_nest_ring(gh_cuda_ctx *ctx, long nside, const long long *pix_in, long long *pix_out, long long n, int to_ring);

/* launches of our own kernels issued by this context since creation */
unsigned long long gh_cuda_kernel_launches(const gh_cuda_ctx *ctx);
/* the CUDA stream (cudaStream_t as void*) all work of this context is enqueued on */
void *gh_cuda_stream(const gh_cuda_ctx *ctx);

const char *gh_cuda_last_error(void);
const char *gh_cuda_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GH_CUDA_H */
