// C-ABI of libgh_cuda.so (include/gh_cuda.h): context, memory, stage sequencing, NCCL plumbing.
// One process per GPU; rank r owns z planes [r*N/P,(r+1)*N/P) of the real-space grids, ky rows of the
// same range of the k-space grids, and ceil(n_nu/P) shells of the reduced map stack.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <new>
#include "gh_internal.cuh"

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";

void gh_set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *gh_cuda_last_error(void) { return g_err; }
extern "C" const char *gh_cuda_version(void) { return "gh_cuda 0.1 (sm_100a)"; }

#define GH_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      gh_set_error(__VA_ARGS__);   \
      return 1;                    \
    }                              \
  } while (0)

// ------------------------------------------------------------------------------------------------ MT19937
// mk_T_maps draws its 30 sub-particle offsets from gsl_rng_mt19937 seeded with seed_rng
// (src/pixelize.c:157-164, src/common.c:133-146).  30 draws on the host; GSL maps seed 0 to 4357 and
// returns u32 / 2^32.
namespace {
struct Mt19937 {
  uint32_t mt[624];
  int idx;
  explicit Mt19937(uint32_t seed)
  {
    if (seed == 0) seed = 4357;
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  uint32_t next()
  {
    if (idx >= 624) {
      for (int k = 0; k < 624; ++k) {
        const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  double uniform() { return next() / 4294967296.0; }
};

// src/cosmo.c:40-50
double host_r_of_z(const gh_cuda_params *p, double z)
{
  if (z <= 0) return 0;
  if (z >= p->z_arr_z2r[p->nz_tab - 1]) return p->r_arr_z2r[p->nz_tab - 1];
  const int iz = (int)(z / p->dz_tab);
  return p->r_arr_z2r[iz] + (p->r_arr_z2r[iz + 1] - p->r_arr_z2r[iz]) * (z - p->z_arr_z2r[iz]) / p->dz_tab;
}

struct StageTimer {
  gh_cuda_ctx *c;
  int slot;
  StageTimer(gh_cuda_ctx *ctx, int s) : c(ctx), slot(s) { cudaEventRecord(c->ev[2 * slot], c->stream); }
  ~StageTimer()
  {
    cudaEventRecord(c->ev[2 * slot + 1], c->stream);
    c->ev_used[slot] = true;
  }
};
}  // namespace

// ------------------------------------------------------------------------------------------------ create / destroy
extern "C" int gh_cuda_get_unique_id(void *id_out)
{
  GH_REQUIRE(id_out, "gh_cuda_get_unique_id: null output");
  static_assert(sizeof(ncclUniqueId) <= GH_CUDA_UNIQUE_ID_BYTES, "unique id size");
  ncclUniqueId id;
  GH_NCCL_OK(ncclGetUniqueId(&id));
  memset(id_out, 0, GH_CUDA_UNIQUE_ID_BYTES);
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

static int upload_tables(gh_cuda_ctx *c, const gh_cuda_params *p)
{
  GhDev &d = c->d;
  const int nz = p->nz_tab, nk = p->numk, nn = p->n_nu;
  // layout: doubles first, then floats
  const size_t n_dbl = (size_t)2 * nk + 4 * nz + 2 * nn;
  const size_t n_flt = (size_t)6 * nz + nn + 1;
  const size_t bytes = (n_dbl * sizeof(double) + n_flt * sizeof(float) + 7) & ~(size_t)7;
  // pinned staging, two buffers used alternately: the copy is ordered on the compute stream behind the previous
  // realisation's kernels and the host does not wait for it (a realisation can be queued while the last one runs)
  const int k = c->stage_next;
  c->stage_next ^= 1;
  if (!c->h_stage[k]) {
    GH_CUDA_OK(cudaMallocHost((void **)&c->h_stage[k], bytes + sizeof(c->h_prefac)));
    GH_CUDA_OK(cudaEventCreateWithFlags(&c->ev_stage[k], cudaEventDisableTiming));
  } else {
    GH_CUDA_OK(cudaEventSynchronize(c->ev_stage[k]));
  }
  GH_REQUIRE(!c->d_tables || bytes == c->tables_bytes, "table sizes changed; create a new context");
  c->stage_cur = k;
  char *h = c->h_stage[k];
  double *hd = (double *)h;
  float *hf = (float *)(h + n_dbl * sizeof(double));
  size_t o = 0;
  auto putd = [&](const double *src, int n) { size_t at = o; memcpy(hd + o, src, sizeof(double) * n); o += n; return at; };
  const size_t o_logk = putd(p->logkarr, nk), o_pk = putd(p->pkarr, nk);
  const size_t o_z = putd(p->z_arr_r2z, nz), o_r = putd(p->r_arr_r2z, nz);
  const size_t o_gd = putd(p->growth_d_arr, nz), o_gv = putd(p->growth_v_arr, nz);
  size_t o_nu0 = o, o_nuf = o;
  if (p->irregular_nutable) { o_nu0 = putd(p->nu0_arr, nn); o_nuf = putd(p->nuf_arr, nn); }
  for (int i = 0; i < nz; ++i) {
    hf[i] = (float)p->z_arr_r2z[i];
    hf[nz + i] = (float)p->growth_d_arr[i];
    hf[2 * nz + i] = (float)p->growth_v_arr[i];
    hf[3 * nz + nn + 1 + i] = (float)p->r_arr_z2r[i];
    // user hooks (src/user_defined.c:27-35) on the radial grid; the reference's shipped formulas when not supplied
    const double z1 = 1.0 + p->z_arr_r2z[i];
    hf[4 * nz + nn + 1 + i] = (float)(p->frac_HI_arr ? p->frac_HI_arr[i] : 0.008 * pow(z1, 0.6));
    hf[5 * nz + nn + 1 + i] = (float)(p->bias_HI_arr ? p->bias_HI_arr[i] : 0.904 + 0.135 * pow(z1, 1.696));
  }
  for (int i = 0; i <= nn; ++i) {  // shell edges for the fp32 fast path
    double e;
    if (p->irregular_nutable) e = (i < nn) ? p->nu0_arr[i] : p->nuf_arr[nn - 1];
    else e = p->nu_min + (p->nu_max - p->nu_min) * (double)(i == 0 ? -1 : i) / nn;  // shell 0 also takes (nu_min-dnu, nu_min): C truncation, src/pixelize.c:216
    hf[3 * nz + i] = (float)e;
  }
  if (!c->d_tables) { GH_CUDA_OK(cudaMalloc(&c->d_tables, bytes)); c->tables_bytes = bytes; }
  GH_CUDA_OK(cudaMemcpyAsync(c->d_tables, h, bytes, cudaMemcpyHostToDevice, c->stream));
  GH_CUDA_OK(cudaEventRecord(c->ev_stage[k], c->stream));
  const double *dd = (const double *)c->d_tables;
  const float *df = (const float *)((const char *)c->d_tables + n_dbl * sizeof(double));
  d.logkarr = dd + o_logk; d.pkarr = dd + o_pk;
  d.z_r2z = dd + o_z; d.r_r2z = dd + o_r; d.gd = dd + o_gd; d.gv = dd + o_gv;
  d.nu0 = dd + o_nu0; d.nuf = dd + o_nuf;
  d.z_r2z_f = df; d.gd_f = df + nz; d.gv_f = df + 2 * nz; d.nu_edges_f = df + 3 * nz; d.r_z2r_f = df + 3 * nz + nn + 1;
  d.frac_f = df + 4 * nz + nn + 1; d.bias_f = df + 5 * nz + nn + 1;
  d.inv_dz_tab = (float)(1.0 / p->dz_tab); d.z_tab_max = (float)p->z_arr_z2r[nz - 1];
  {
    double sv = 0;
    for (int i = 0; i + 2 < nz; ++i) sv = fmax(sv, fabs(p->r_arr_z2r[i + 2] - 2 * p->r_arr_z2r[i + 1] + p->r_arr_z2r[i]));
    d.rz_slope_var = (float)(1.001 * sv / p->dz_tab);
  }
  return 0;
}

// per-shell prefactors through the tail of the same staging buffer (call after upload_tables)
static int upload_prefac(gh_cuda_ctx *c)
{
  const int k = c->stage_cur;
  char *h = c->h_stage[k] + c->tables_bytes;
  memcpy(h, c->h_prefac, sizeof(double) * c->d.n_nu_pad);
  GH_CUDA_OK(cudaMemcpyAsync(c->d_prefac, h, sizeof(double) * c->d.n_nu_pad, cudaMemcpyHostToDevice, c->stream));
  GH_CUDA_OK(cudaEventRecord(c->ev_stage[k], c->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------ map-plane load balance
// mk_T_maps' cost per cell is far from uniform: cells outside the shells' radial window are dropped after a short
// test, the others run ten sub-particles through the pixelisation.  With equal z slabs the ranks holding the box's
// centre planes therefore work ~2.4x longer than the edge ranks (measured, 8 x B200, 2048^3).  The accumulation
// is re-partitioned into contiguous plane ranges of equal modelled cost; a rank whose range reaches outside its
// own slab pulls those planes' HI mass and Delta z_RSD from the owner over NVLink (enqueue_maps).
// Cost model per plane: 0.18 + 0.82 * (fraction of the plane's cells with r_lo - h < r < r_hi + h), the weights
// fitted to the per-slab times measured on 4 and 8 GPUs; the fraction is computed from the exact x-extent of
// the annulus on a sample of rows.  Pure host code; every rank computes the same bounds.
extern "C" int gh_cuda_map_plane_bounds(const gh_cuda_params *p, int nranks, int *bounds)
{
  GH_REQUIRE(p && bounds && nranks >= 1 && nranks <= GH_MAX_RANKS && p->n_grid > 0 && p->n_grid % nranks == 0,
             "gh_cuda_map_plane_bounds: bad arguments");
  const int n = p->n_grid;
  const double dx = p->l_box / n, h = 0.8660254 * dx;
  double nu_lo, nu_hi;
  if (p->irregular_nutable) { nu_lo = p->nu0_arr[0]; nu_hi = p->nuf_arr[p->n_nu - 1]; }
  else { nu_lo = p->nu_min - (p->nu_max - p->nu_min) / p->n_nu; nu_hi = p->nu_max; }
  const double r_lo = fmax(0.0, host_r_of_z(p, GH_CUDA_NU_21 / nu_hi - 1) - h);
  const double r_hi = host_r_of_z(p, GH_CUDA_NU_21 / nu_lo - 1) + h;
  const double xmin = -p->pos_obs[0], xmax = p->l_box - p->pos_obs[0];
  auto overlap = [&](double a, double b) { const double lo = a > xmin ? a : xmin, hi = b < xmax ? b : xmax; return hi > lo ? hi - lo : 0.0; };
  const int stride = n > 256 ? n / 256 : 1;
  double *w = (double *)malloc(sizeof(double) * (size_t)(n + 1));
  GH_REQUIRE(w, "out of host memory");
  w[0] = 0;
  for (int z = 0; z < n; ++z) {
    const double zc = dx * (z + 0.5) - p->pos_obs[2];
    double len = 0;
    int rows = 0;
    for (int y = stride / 2; y < n; y += stride, ++rows) {
      const double yc = dx * (y + 0.5) - p->pos_obs[1];
      const double rho2 = yc * yc + zc * zc, b2 = r_hi * r_hi - rho2;
      if (b2 <= 0) continue;
      const double b = sqrt(b2), a = r_lo * r_lo > rho2 ? sqrt(r_lo * r_lo - rho2) : 0.0;
      len += overlap(a, b) + overlap(-b, -a);
    }
    w[z + 1] = w[z] + 0.18 + 0.82 * len / (p->l_box * (rows > 0 ? rows : 1));
  }
  bounds[0] = 0;
  int z = 0;
  for (int r = 1; r < nranks; ++r) {
    const double target = w[n] * r / nranks;
    while (z < n && w[z] < target) ++z;
    // the plane boundary closest to the target
    if (z > bounds[r - 1] && target - w[z - 1] < w[z] - target) --z;
    bounds[r] = z;
  }
  bounds[nranks] = n;
  free(w);
  return 0;
}

// scalars, sub-particle offsets and per-shell prefactors derived from the parameter block
static int apply_params(gh_cuda_ctx *c, const gh_cuda_params *p, int rank, int nranks)
{
  GhDev &d = c->d;
  d.n = p->n_grid; d.nh = p->n_grid / 2 + 1;
  d.nranks = nranks; d.rank = rank;
  d.nz_here = d.n / nranks; d.iz0 = rank * d.nz_here;
  d.nky_here = d.nz_here; d.ky0 = d.iz0;
  d.l_box = p->l_box; d.dx = p->l_box / p->n_grid;
  d.half_inv_dx = (float)(0.5 / d.dx);
  for (int i = 0; i < 3; ++i) d.pos_obs[i] = p->pos_obs[i];
  d.seed = p->seed_rng; d.do_smoothing = p->do_smoothing; d.r2_smooth = p->r2_smooth;
  d.vfactor = p->fgrowth_0 * p->hubble_0;
  d.dk = 2 * M_PI / p->l_box; d.idk3 = 1. / (d.dk * d.dk * d.dk);
  d.numk = p->numk; d.logkmin = p->logkmin; d.logkmax = p->logkmax; d.idlogk = p->idlogk; d.n_scal = p->n_scal;
  d.nz_tab = p->nz_tab; d.glob_idr = p->glob_idr; d.r_tab_max = p->r_arr_r2z[p->nz_tab - 1];
  d.nside = p->n_side; d.npix = 12LL * p->n_side * p->n_side;
  d.n_nu = p->n_nu; d.irregular = p->irregular_nutable;
  const int shells_per_rank = (p->n_nu + nranks - 1) / nranks;
  d.n_nu_pad = shells_per_rank * nranks;
  if (p->irregular_nutable) { d.nu_min = p->nu0_arr[0]; d.nu_max = p->nuf_arr[p->n_nu - 1]; }
  else { d.nu_min = p->nu_min; d.nu_max = p->nu_max; }
  d.inv_dnu = p->n_nu / (d.nu_max - d.nu_min);  // src/pixelize.c:176-178
  {
    // redshift window that can still land in a shell: nu in [nu_lo, nu_hi); the regular-table personality
    // truncates toward zero, which also accepts (nu_min - dnu, nu_min) into shell 0 (src/pixelize.c:216)
    const double nu_lo = p->irregular_nutable ? d.nu_min : d.nu_min - 1.0 / d.inv_dnu;
    d.z_hi_cull = GH_CUDA_NU_21 / nu_lo - 1 + 1e-9;
    d.z_lo_cull = GH_CUDA_NU_21 / d.nu_max - 1 - 1e-9;
  }
  {
    Mt19937 g(p->seed_rng);
    const double lcell = p->l_box / p->n_grid;
    for (int i = 0; i < GH_CUDA_N_SUBPART; ++i) {  // interleaved x,y,z draws, src/pixelize.c:160-164
      d.sub_off[i] = lcell * (g.uniform() - 0.5);
      d.sub_off[GH_CUDA_N_SUBPART + i] = lcell * (g.uniform() - 0.5);
      d.sub_off[2 * GH_CUDA_N_SUBPART + i] = lcell * (g.uniform() - 0.5);
    }
    for (int i = 0; i < 3 * GH_CUDA_N_SUBPART; ++i) d.sub_off_f[i] = (float)d.sub_off[i];
    for (int i = 0; i < GH_CUDA_N_SUBPART; ++i) {
      const float ox = d.sub_off_f[i], oy = d.sub_off_f[GH_CUDA_N_SUBPART + i], oz = d.sub_off_f[2 * GH_CUDA_N_SUBPART + i];
      d.sub_c[i] = make_float4(ox, oy, oz, ox * ox + oy * oy + oz * oz);
    }
  }
  {
    // src/pixelize.c:155,246-256
    const double m2t = 90.057156 * p->OmegaB * p->hhub * (double)d.npix / (4 * M_PI);
    for (int inu = 0; inu < d.n_nu_pad; ++inu) {
      if (inu >= p->n_nu) { c->h_prefac[inu] = 0; c->h_nu_centre[inu] = 1; continue; }
      double dnu, nu;
      if (p->irregular_nutable) { dnu = p->nuf_arr[inu] - p->nu0_arr[inu]; nu = (p->nuf_arr[inu] + p->nu0_arr[inu]) * 0.5; }
      else { dnu = (p->nu_max - p->nu_min) / p->n_nu; nu = p->nu_min + (inu + 0.5) * dnu; }
      const double r = host_r_of_z(p, GH_CUDA_NU_21 / nu - 1);
      c->h_prefac[inu] = m2t / (r * r * dnu);
      c->h_nu_centre[inu] = nu;
    }
  }
  if (nranks > 1) {
    if (gh_cuda_map_plane_bounds(p, nranks, c->map_bounds)) return 1;
    if (const char *ov = getenv("GH_MAP_BOUNDS")) {  // test hook: "b1,b2,...": interior bounds, planes
      int prev = 0;
      for (int r = 1; r < nranks && ov && *ov; ++r) {
        int v = atoi(ov);
        v = v < prev ? prev : v > d.n ? d.n : v;
        c->map_bounds[r] = prev = v;
        ov = strchr(ov, ',');
        if (ov) ++ov;
      }
    }
  }
  return 0;
}

// Stream-ordered barrier across ranks: a 1-int all-reduce.  When it completes on this rank's stream every rank
// has enqueued-and-reached the same point, so memory that peers wrote or read before it is settled.
int gh_stream_barrier(gh_cuda_ctx *c)
{
  if (c->d.nranks <= 1) return 0;
  GH_NCCL_OK(ncclAllReduce(c->d_barrier, c->d_barrier, 1, ncclInt, ncclSum, c->comm, c->stream));
  return 0;
}

// Exchange CUDA IPC handles of the slab buffers through the communicator and map every peer's buffers, so
// kernels can address other GPUs' slabs directly (NVLink loads / stores).  One process per GPU, one node.
static int setup_peers(gh_cuda_ctx *c)
{
  const int P = c->d.nranks, me = c->d.rank;
  GH_REQUIRE(P <= GH_MAX_RANKS, "at most %d ranks", GH_MAX_RANKS);
  GH_CUDA_OK(cudaMalloc(&c->d_barrier, sizeof(int)));
  GH_CUDA_OK(cudaMemsetAsync(c->d_barrier, 0, sizeof(int), c->stream));
  const size_t hsz = sizeof(cudaIpcMemHandle_t);
  cudaIpcMemHandle_t mine[2];
  GH_CUDA_OK(cudaIpcGetMemHandle(&mine[0], c->gridA));
  GH_CUDA_OK(cudaIpcGetMemHandle(&mine[1], c->gridC));
  char *d_all = nullptr;
  GH_CUDA_OK(cudaMalloc(&d_all, 2 * hsz * P));
  GH_CUDA_OK(cudaMemcpyAsync(d_all + 2 * hsz * me, mine, 2 * hsz, cudaMemcpyHostToDevice, c->stream));
  GH_NCCL_OK(ncclAllGather(d_all + 2 * hsz * me, d_all, 2 * hsz, ncclChar, c->comm, c->stream));
  cudaIpcMemHandle_t *all = (cudaIpcMemHandle_t *)malloc(2 * hsz * P);
  GH_REQUIRE(all, "out of host memory");
  cudaError_t e = cudaMemcpyAsync(all, d_all, 2 * hsz * P, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_all);
  if (e != cudaSuccess) { free(all); gh_set_error("IPC handle exchange failed: %s", cudaGetErrorString(e)); return 1; }
  bool ok = true;
  for (int q = 0; q < P && ok; ++q) {
    if (q == me) { c->peers.A[q] = c->gridA; c->peers.C[q] = c->gridC; continue; }
    void *pa = nullptr, *pc = nullptr;
    if (cudaIpcOpenMemHandle(&pa, all[2 * q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&pc, all[2 * q + 1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      ok = false;
      break;
    }
    c->peers.A[q] = (float2 *)pa;
    c->peers.C[q] = (float2 *)pc;
  }
  free(all);
  // every rank must take the same route: agree on min(ok) over the ranks
  {
    cudaGetLastError();
    int flag = ok ? 1 : 0;
    GH_CUDA_OK(cudaMemcpyAsync(c->d_barrier, &flag, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    GH_NCCL_OK(ncclAllReduce(c->d_barrier, c->d_barrier, 1, ncclInt, ncclMin, c->comm, c->stream));
    GH_CUDA_OK(cudaMemcpyAsync(&flag, c->d_barrier, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GH_CUDA_OK(cudaStreamSynchronize(c->stream));
    GH_CUDA_OK(cudaMemsetAsync(c->d_barrier, 0, sizeof(int), c->stream));
    ok = flag == 1;
  }
  if (!ok) {
    // not fatal: fall back to NCCL-only data movement (still all on the GPUs)
    for (int q = 0; q < P; ++q) {
      if (q != me && c->peers.A[q]) cudaIpcCloseMemHandle(c->peers.A[q]);
      if (q != me && c->peers.C[q]) cudaIpcCloseMemHandle(c->peers.C[q]);
      c->peers.A[q] = c->peers.C[q] = nullptr;
    }
    cudaGetLastError();
    c->have_peers = false;
    return 0;
  }
  c->have_peers = getenv("GH_NO_PEER") == nullptr;
  c->rebalance = c->have_peers && getenv("GH_NO_REBALANCE") == nullptr;
  return 0;
}

// Map every peer's accumulation stack too, for the sparse map reduction (default; GH_NO_SPARSE_REDUCE=1 turns it off)
// (gh_pixelize.cu).  Same handle exchange as setup_peers; a failure anywhere turns the feature off everywhere.
static int setup_map_peers(gh_cuda_ctx *c)
{
  const int P = c->d.nranks, me = c->d.rank;
  const size_t hsz = sizeof(cudaIpcMemHandle_t);
  cudaIpcMemHandle_t mine;
  GH_CUDA_OK(cudaIpcGetMemHandle(&mine, c->maps));
  char *d_all = nullptr;
  GH_CUDA_OK(cudaMalloc(&d_all, hsz * P));
  GH_CUDA_OK(cudaMemcpyAsync(d_all + hsz * me, &mine, hsz, cudaMemcpyHostToDevice, c->stream));
  GH_NCCL_OK(ncclAllGather(d_all + hsz * me, d_all, hsz, ncclChar, c->comm, c->stream));
  cudaIpcMemHandle_t *all = (cudaIpcMemHandle_t *)malloc(hsz * P);
  GH_REQUIRE(all, "out of host memory");
  cudaError_t e = cudaMemcpyAsync(all, d_all, hsz * P, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_all);
  if (e != cudaSuccess) { free(all); gh_set_error("IPC handle exchange failed: %s", cudaGetErrorString(e)); return 1; }
  bool ok = true;
  for (int q = 0; q < P && ok; ++q) {
    if (q == me) { c->map_peers[q] = c->maps; continue; }
    void *pm = nullptr;
    if (cudaIpcOpenMemHandle(&pm, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
    else c->map_peers[q] = (float *)pm;
  }
  free(all);
  cudaGetLastError();
  int flag = ok ? 1 : 0;
  GH_CUDA_OK(cudaMemcpyAsync(c->d_barrier, &flag, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  GH_NCCL_OK(ncclAllReduce(c->d_barrier, c->d_barrier, 1, ncclInt, ncclMin, c->comm, c->stream));
  GH_CUDA_OK(cudaMemcpyAsync(&flag, c->d_barrier, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  GH_CUDA_OK(cudaMemsetAsync(c->d_barrier, 0, sizeof(int), c->stream));
  if (flag != 1) {
    for (int q = 0; q < P; ++q) {
      if (q != me && c->map_peers[q]) cudaIpcCloseMemHandle(c->map_peers[q]);
      c->map_peers[q] = nullptr;
    }
    cudaGetLastError();
    return 0;
  }
  GH_CUDA_OK(cudaMalloc(&c->d_ext, sizeof(int) * 2 * (size_t)c->d.n_nu_pad * (P + 1)));
  c->sparse_reduce = true;
  return 0;
}

// Collective: map `mine` (a cudaMalloc'ed buffer, or nullptr when this rank could not allocate one) on every peer.
// ok_out is the agreement of all ranks (min over "I have a buffer and mapped everybody's").
static int exchange_ipc(gh_cuda_ctx *c, void *mine_ptr, void **peers_out, bool *ok_out)
{
  const int P = c->d.nranks, me = c->d.rank;
  const size_t hsz = sizeof(cudaIpcMemHandle_t);
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  bool ok = mine_ptr != nullptr && cudaIpcGetMemHandle(&mine, mine_ptr) == cudaSuccess;
  // a first agreement: nobody opens handles unless everybody has one
  int flag = ok ? 1 : 0;
  GH_CUDA_OK(cudaMemcpyAsync(c->d_barrier, &flag, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  GH_NCCL_OK(ncclAllReduce(c->d_barrier, c->d_barrier, 1, ncclInt, ncclMin, c->comm, c->stream));
  GH_CUDA_OK(cudaMemcpyAsync(&flag, c->d_barrier, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  GH_CUDA_OK(cudaMemsetAsync(c->d_barrier, 0, sizeof(int), c->stream));
  cudaGetLastError();
  *ok_out = false;
  if (flag != 1) return 0;
  char *d_all = nullptr;
  GH_CUDA_OK(cudaMalloc(&d_all, hsz * P));
  GH_CUDA_OK(cudaMemcpyAsync(d_all + hsz * me, &mine, hsz, cudaMemcpyHostToDevice, c->stream));
  GH_NCCL_OK(ncclAllGather(d_all + hsz * me, d_all, hsz, ncclChar, c->comm, c->stream));
  cudaIpcMemHandle_t *all = (cudaIpcMemHandle_t *)malloc(hsz * P);
  GH_REQUIRE(all, "out of host memory");
  cudaError_t e = cudaMemcpyAsync(all, d_all, hsz * P, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_all);
  if (e != cudaSuccess) { free(all); gh_set_error("IPC handle exchange failed: %s", cudaGetErrorString(e)); return 1; }
  for (int q = 0; q < P && ok; ++q) {
    if (q == me) { peers_out[q] = mine_ptr; continue; }
    void *pm = nullptr;
    if (cudaIpcOpenMemHandle(&pm, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
    else peers_out[q] = pm;
  }
  free(all);
  cudaGetLastError();
  flag = ok ? 1 : 0;
  GH_CUDA_OK(cudaMemcpyAsync(c->d_barrier, &flag, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  GH_NCCL_OK(ncclAllReduce(c->d_barrier, c->d_barrier, 1, ncclInt, ncclMin, c->comm, c->stream));
  GH_CUDA_OK(cudaMemcpyAsync(&flag, c->d_barrier, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  GH_CUDA_OK(cudaMemsetAsync(c->d_barrier, 0, sizeof(int), c->stream));
  if (flag != 1) {
    for (int q = 0; q < P; ++q) {
      if (q != me && peers_out[q]) cudaIpcCloseMemHandle(peers_out[q]);
      peers_out[q] = nullptr;
    }
    cudaGetLastError();
    return 0;
  }
  *ok_out = true;
  return 0;
}

// Streams, events and the second receive buffer of the copy-engine transposes (fft_both_fields_ce, gh_fft.cu)
static int setup_ce_transpose(gh_cuda_ctx *c)
{
  const GhDev &d = c->d;
  const char *mode = getenv("GH_TRANSPOSE");  // "ce" (default) | "nccl" | "fused"
  c->ce_transpose = getenv("GH_FUSED_TRANSPOSE") == nullptr && !(mode && !strcmp(mode, "fused"));
  if (!c->ce_transpose) return 0;
  if (mode && !strcmp(mode, "nccl")) {
    // a second communicator for the transposes: its id travels through the first one
    ncclUniqueId id2;
    if (d.rank == 0) GH_NCCL_OK(ncclGetUniqueId(&id2));
    char *d_id = nullptr;
    GH_CUDA_OK(cudaMalloc(&d_id, sizeof(id2)));
    GH_CUDA_OK(cudaMemcpyAsync(d_id, &id2, sizeof(id2), cudaMemcpyHostToDevice, c->stream));
    GH_NCCL_OK(ncclBroadcast(d_id, d_id, sizeof(id2), ncclChar, 0, c->comm, c->stream));
    GH_CUDA_OK(cudaMemcpyAsync(&id2, d_id, sizeof(id2), cudaMemcpyDeviceToHost, c->stream));
    GH_CUDA_OK(cudaStreamSynchronize(c->stream));
    cudaFree(d_id);
    GH_NCCL_OK(ncclCommInitRank(&c->comm2, d.nranks, id2, d.rank));
    c->have_comm2 = true;
    c->nccl_transpose = true;
  }
  {
    // high priority: the transposes' CTAs (push kernel, NCCL) are scheduled ahead of the queued FFT CTAs
    int lo = 0, hi = 0;
    GH_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (int k = 0; k < GH_N_COPY_STREAMS; ++k) GH_CUDA_OK(cudaStreamCreateWithPriority(&c->ce_stream[k], cudaStreamNonBlocking, hi));
  }
  // One stream: the copies to the peers run one after the other in the staggered order, so at any moment the GPUs form a
  // perfect matching and every link carries one full-rate copy.  Spreading them over several streams (concurrent copies to
  // several peers) was measured to cut the rate to a third on 4 and 8 GPUs (4 GPUs, 1024^3: 685 GB/s with one stream, 370
  // with two, 268 with four; profiles/r2/bench_4gpu_ce_streams_*.json)
  c->ce_streams_used = getenv("GH_CE_STREAMS") ? atoi(getenv("GH_CE_STREAMS")) : 1;
  if (c->ce_streams_used < 1 || c->ce_streams_used > GH_N_COPY_STREAMS) c->ce_streams_used = 1;
  c->push_transpose = mode && !strcmp(mode, "push");
  c->push_ctas = getenv("GH_PUSH_CTAS") ? atoi(getenv("GH_PUSH_CTAS")) : 64;
  if (c->push_ctas < 1) c->push_ctas = 1;
  for (int f = 0; f < 2; ++f) {
    GH_CUDA_OK(cudaEventCreateWithFlags(&c->ev_z[f], cudaEventDisableTiming));
    GH_CUDA_OK(cudaEventCreateWithFlags(&c->ev_free[f], cudaEventDisableTiming));
    for (int k = 0; k < GH_N_COPY_STREAMS; ++k) GH_CUDA_OK(cudaEventCreateWithFlags(&c->ev_sent[f][k], cudaEventDisableTiming));
  }
  const size_t slab_bytes = c->slab_complex * sizeof(float2);
  const size_t maps_bytes = (size_t)d.n_nu_pad * d.npix * sizeof(float);
  if (getenv("GH_ONE_RECV_BUFFER")) return 0;  // test hook: the single-buffer sequence
  if (c->sparse_reduce && maps_bytes >= slab_bytes && !getenv("GH_OWN_RECV_BUFFER")) {
    // the map accumulation stack is idle during the FFTs and already mapped on every peer
    c->recv2 = reinterpret_cast<float2 *>(c->maps);
    for (int q = 0; q < d.nranks; ++q) c->recv2_peers[q] = reinterpret_cast<float2 *>(c->map_peers[q]);
    return 0;
  }
  size_t free_b = 0, total_b = 0;
  GH_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
  void *buf = nullptr;
  if (free_b > slab_bytes + total_b / 8 && cudaMalloc(&buf, slab_bytes) != cudaSuccess) buf = nullptr;
  cudaGetLastError();
  bool ok = false;
  void *peers[GH_MAX_RANKS] = {nullptr};
  if (exchange_ipc(c, buf, peers, &ok)) return 1;
  if (!ok) { if (buf) cudaFree(buf); return 0; }
  c->recv2 = (float2 *)buf;
  c->recv2_owned = true;
  for (int q = 0; q < d.nranks; ++q) c->recv2_peers[q] = (float2 *)peers[q];
  return 0;
}

extern "C" int gh_cuda_destroy(gh_cuda_ctx *c)
{
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  gh_psources_release(c);
  for (int k = 0; k < GH_N_COPY_STREAMS; ++k)
    if (c->ce_stream[k]) { cudaStreamSynchronize(c->ce_stream[k]); cudaStreamDestroy(c->ce_stream[k]); c->ce_stream[k] = nullptr; }
  for (int f = 0; f < 2; ++f) {
    if (c->ev_z[f]) cudaEventDestroy(c->ev_z[f]);
    if (c->ev_free[f]) cudaEventDestroy(c->ev_free[f]);
    for (int k = 0; k < GH_N_COPY_STREAMS; ++k)
      if (c->ev_sent[f][k]) cudaEventDestroy(c->ev_sent[f][k]);
  }
  if (c->have_comm && c->d_barrier) {
    // nobody may free a slab a peer could still be reading
    ncclAllReduce(c->d_barrier, c->d_barrier, 1, ncclInt, ncclSum, c->comm, c->stream);
    cudaStreamSynchronize(c->stream);
  }
  if (c->recv2_owned) {
    for (int q = 0; q < c->d.nranks && q < GH_MAX_RANKS; ++q)
      if (q != c->d.rank && c->recv2_peers[q]) cudaIpcCloseMemHandle(c->recv2_peers[q]);
    cudaFree(c->recv2);
  }
  for (int q = 0; q < c->d.nranks && q < GH_MAX_RANKS; ++q) {
    if (q == c->d.rank) continue;
    if (c->peers.A[q]) cudaIpcCloseMemHandle(c->peers.A[q]);
    if (c->peers.C[q]) cudaIpcCloseMemHandle(c->peers.C[q]);
  }
  for (int q = 0; q < c->d.nranks && q < GH_MAX_RANKS; ++q)
    if (q != c->d.rank && c->map_peers[q]) cudaIpcCloseMemHandle(c->map_peers[q]);
  cudaFree(c->d_ext);
  cudaFree(c->d_barrier);
  if (c->have_comm2) ncclCommDestroy(c->comm2);
  if (c->have_comm) ncclCommDestroy(c->comm);
  cudaFree(c->gridA); cudaFree(c->gridB); cudaFree(c->gridC);
  cudaFree(c->halo_lo); cudaFree(c->halo_hi);
  for (int b = 0; b < 2; ++b) {
    if (c->out_buf[b]) cudaFree(c->out_buf[b]);
    if (c->ev_copied[b]) cudaEventDestroy(c->ev_copied[b]);
    if (c->h_stage[b]) cudaFreeHost(c->h_stage[b]);
    if (c->ev_stage[b]) cudaEventDestroy(c->ev_stage[b]);
  }
  if (c->maps != c->out_buf[0] && c->maps != c->out_buf[1]) cudaFree(c->maps);  // one rank: maps is one of out_buf[]
  cudaFree(c->twiddle); cudaFree(c->d_partials); cudaFree(c->d_prefac); cudaFree(c->d_tables);
  for (int i = 0; i < 2 * GH_T_NSLOTS; ++i)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->pull_stream) { cudaStreamSynchronize(c->pull_stream); cudaStreamDestroy(c->pull_stream); }
  for (int i = 0; i < GH_MAX_CHUNKS; ++i)
    if (c->ev_chunk[i]) cudaEventDestroy(c->ev_chunk[i]);
  for (int f = 0; f < 2; ++f)
    for (int k = 0; k < 2; ++k)
      if (c->ev_pass[f][k]) cudaEventDestroy(c->ev_pass[f][k]);
  if (c->ev_bar) cudaEventDestroy(c->ev_bar);
  if (c->ev_pulled) cudaEventDestroy(c->ev_pulled);
  if (c->ev_chunk_free) cudaEventDestroy(c->ev_chunk_free);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  if (c->h_stats) cudaFreeHost(c->h_stats);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

extern "C" int gh_cuda_create(const gh_cuda_params *p, int rank, int nranks, const void *unique_id, int device,
                              gh_cuda_ctx **ctx_out)
{
  GH_REQUIRE(p && ctx_out, "gh_cuda_create: null argument");
  *ctx_out = nullptr;
  GH_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "gh_cuda_create: bad rank %d of %d", rank, nranks);
  GH_REQUIRE(gh_fft_supported(p->n_grid), "n_grid=%d unsupported (even, 8..4096)", p->n_grid);
  GH_REQUIRE(nranks == 1 || gh_fft_tuned(p->n_grid),
             "n_grid=%d unsupported on %d ranks (several ranks need a power of two 32..4096; other even sizes run on one)", p->n_grid, nranks);
  GH_REQUIRE((nranks & (nranks - 1)) == 0 && p->n_grid % nranks == 0 && p->n_grid / nranks >= 2,
             "n_grid=%d cannot be split into %d slabs (need a power-of-two rank count, >=2 planes each)", p->n_grid, nranks);
  GH_REQUIRE(p->n_side >= 1 && p->n_nu >= 1 && p->n_nu <= 4096, "bad n_side=%ld / n_nu=%d", p->n_side, p->n_nu);
  GH_REQUIRE(p->nz_tab >= 2 && p->nz_tab <= GH_NZ_TAB_MAX && p->numk >= 2, "bad table sizes");
  GH_REQUIRE(p->logkarr && p->pkarr && p->z_arr_r2z && p->r_arr_r2z && p->growth_d_arr && p->growth_v_arr &&
                 p->z_arr_z2r && p->r_arr_z2r, "gh_cuda_create: missing table pointer");
  GH_REQUIRE(!p->irregular_nutable || (p->nu0_arr && p->nuf_arr), "irregular nu table requested but not supplied");
  GH_REQUIRE((p->frac_HI_arr == nullptr) == (p->bias_HI_arr == nullptr), "supply both frac_HI_arr and bias_HI_arr, or neither");
  if (p->irregular_nutable)
    for (int i = 0; i + 1 < p->n_nu; ++i)
      GH_REQUIRE(p->nuf_arr[i] == p->nu0_arr[i + 1], "frequency bins must be contiguous (the reference's get_inu never terminates otherwise)");
  GH_REQUIRE(nranks == 1 || unique_id, "gh_cuda_create: nranks>1 needs the NCCL unique id");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  GH_REQUIRE(e == cudaSuccess && ndev > 0, "no CUDA device available (%s); libgh_cuda has no CPU fallback",
             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  GH_REQUIRE(device >= 0 && device < ndev, "device %d out of range (%d visible)", device, ndev);
  GH_CUDA_OK(cudaSetDevice(device));

  gh_cuda_ctx *c = new (std::nothrow) gh_cuda_ctx();
  GH_REQUIRE(c, "out of host memory");
  memset(c, 0, sizeof(*c));
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; gh_set_error("cudaGetDeviceProperties failed"); return 1; }
  c->n_sm = prop.multiProcessorCount;
  {
    // y/x FFT plane batch in MB (GH_FFT_BATCH_MB; 0 / unset = whole slab per pass)
    const char *e = getenv("GH_FFT_BATCH_MB");
    const double mb = e ? atof(e) : 0.0;  // measured: batching through L2 does not pay on B200 (profiles/), off by default
    c->fft_batch_bytes = mb > 0 ? (size_t)(mb * 1024.0 * 1024.0) : (size_t)1 << 60;
  }

  if (apply_params(c, p, rank, nranks)) { delete c; return 1; }
  GhDev &d = c->d;
  const int shells_per_rank = d.n_nu_pad / nranks;

#define CREATE_OK(call)                                                                                   \
  do {                                                                                                    \
    cudaError_t e__ = (call);                                                                             \
    if (e__ != cudaSuccess) {                                                                             \
      gh_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));           \
      gh_cuda_destroy(c);                                                                                 \
      return 1;                                                                                           \
    }                                                                                                     \
  } while (0)

  CREATE_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  if (getenv("GH_TIME_FFT_PASSES")) {
    for (int f = 0; f < 2; ++f)
      for (int k = 0; k < 2; ++k) CREATE_OK(cudaEventCreate(&c->ev_pass[f][k]));
    c->time_fft_passes = true;
  }
  // gh_cuda_run*: radial velocity and get_HI in one pass (the multi-GPU parity check compares it with the staged calls)
  c->fuse_vel = getenv("GH_NO_FUSE_VEL") == nullptr;
  c->defer_d2h = getenv("GH_NO_DEFER_D2H") == nullptr;  // measured: e2e loop 6.54 -> 6.27 ms per realisation at 512^3 (profiles/r2/e2e_probe*.log)
  // TMA-fed strided FFT passes: measured faster up to 512 (1.66 -> 1.56 ms both fields at 512^3), equal at 1024 (16.2 vs 16.4 ms),
  // slower at 2048 where the 64 KB tile is only 4 lines wide and the 32-byte store rows dominate (217 vs 168 ms on one GPU;
  // profiles/r2/): on by default for n_grid <= 512, GH_FFT_TMA=1 / GH_FFT_NO_TMA=1 force it on / off
  c->fft_tma = getenv("GH_FFT_NO_TMA") == nullptr && (p->n_grid <= 512 || getenv("GH_FFT_TMA") != nullptr);
  for (int i = 0; i < 4; ++i) c->fft_map_ok[i] = false;
  CREATE_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CREATE_OK(cudaStreamCreateWithFlags(&c->pull_stream, cudaStreamNonBlocking));
  CREATE_OK(cudaEventCreateWithFlags(&c->ev_bar, cudaEventDisableTiming));
  CREATE_OK(cudaEventCreateWithFlags(&c->ev_pulled, cudaEventDisableTiming));
  CREATE_OK(cudaEventCreateWithFlags(&c->ev_chunk_free, cudaEventDisableTiming));
  CREATE_OK(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
  for (int b = 0; b < 2; ++b) CREATE_OK(cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming));
  CREATE_OK(cudaHostAlloc((void **)&c->h_stats, 4 * sizeof(double), cudaHostAllocMapped));
  CREATE_OK(cudaHostGetDevicePointer((void **)&c->h_stats_dev, c->h_stats, 0));
  for (int i = 0; i < 2 * GH_T_NSLOTS; ++i) CREATE_OK(cudaEventCreate(&c->ev[i]));
  if (upload_tables(c, p)) { gh_cuda_destroy(c); return 1; }

  c->slab_complex = (size_t)d.nz_here * d.n * d.nh;
  const size_t slab_bytes = c->slab_complex * sizeof(float2);
  CREATE_OK(cudaMalloc(&c->gridA, slab_bytes));
  CREATE_OK(cudaMalloc(&c->gridB, slab_bytes));
  CREATE_OK(cudaMalloc(&c->gridC, slab_bytes));
  const size_t map_bytes = (size_t)d.n_nu_pad * d.npix * sizeof(float);
  CREATE_OK(cudaMalloc(&c->maps, map_bytes));
  CREATE_OK(cudaMemsetAsync(c->maps, 0, map_bytes, c->stream));
  // out_buf[]: what the device->host copy reads -- the stack itself on one rank, the reduce-scatter output on
  // several.  A second copy of it (when it fits comfortably) lets realisation i+1 write its result while the copy
  // of realisation i is still on the wire.
  size_t out_bytes = map_bytes;
  if (nranks > 1) {
    const size_t plane_bytes = (size_t)2 * d.nh * d.n * sizeof(float);
    CREATE_OK(cudaMalloc(&c->halo_lo, plane_bytes));
    CREATE_OK(cudaMalloc(&c->halo_hi, plane_bytes));
    out_bytes = (size_t)shells_per_rank * d.npix * sizeof(float);
    CREATE_OK(cudaMalloc(&c->out_buf[0], out_bytes));
    c->maps_recv = c->out_buf[0];
  } else {
    c->out_buf[0] = c->maps;
  }
  {
    size_t free_b = 0, total_b = 0;
    CREATE_OK(cudaMemGetInfo(&free_b, &total_b));
    const bool want = !getenv("GH_SINGLE_OUT") && free_b > out_bytes && free_b - out_bytes > total_b / 8;
    if (want && cudaMalloc(&c->out_buf[1], out_bytes) != cudaSuccess) {
      c->out_buf[1] = nullptr;
      (void)cudaGetLastError();
    }
    if (c->out_buf[1] && nranks == 1) CREATE_OK(cudaMemsetAsync(c->out_buf[1], 0, out_bytes, c->stream));
  }
  CREATE_OK(cudaMalloc(&c->d_partials, sizeof(double) * (8 + 2 * ((size_t)c->n_sm * 8 + (size_t)d.nz_here * d.n))));
  CREATE_OK(cudaMalloc(&c->d_prefac, sizeof(double) * d.n_nu_pad));
  if (upload_prefac(c)) { gh_cuda_destroy(c); return 1; }
  {
    float2 *tw = (float2 *)malloc(sizeof(float2) * d.n);
    if (!tw) { gh_cuda_destroy(c); gh_set_error("out of host memory"); return 1; }
    for (int j = 0; j < d.n; ++j) {
      const double a = 2.0 * M_PI * (double)j / (double)d.n;
      tw[j] = make_float2((float)cos(a), (float)sin(a));
    }
    cudaError_t e2 = cudaMalloc(&c->twiddle, sizeof(float2) * d.n);
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(c->twiddle, tw, sizeof(float2) * d.n, cudaMemcpyHostToDevice, c->stream);
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(c->stream);
    free(tw);
    CREATE_OK(e2);
  }
  if (nranks > 1) {
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclResult_t r = ncclCommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
      gh_set_error("ncclCommInitRank failed: %s", ncclGetErrorString(r));
      gh_cuda_destroy(c);
      return 1;
    }
    c->have_comm = true;
    if (setup_peers(c)) { gh_cuda_destroy(c); return 1; }
    // sparse map reduction over peer memory: validated against the single-GPU run on 2 and 4 GPUs and 1.9x faster than
    // ncclReduceScatter there (profiles/r2/); GH_NO_SPARSE_REDUCE=1 restores the NCCL collective
    if (c->have_peers && !getenv("GH_NO_SPARSE_REDUCE") && setup_map_peers(c)) { gh_cuda_destroy(c); return 1; }
    if (c->have_peers && setup_ce_transpose(c)) { gh_cuda_destroy(c); return 1; }
  }
  CREATE_OK(cudaStreamSynchronize(c->stream));
#undef CREATE_OK
  c->sigma2_gauss = -1;
  *ctx_out = c;
  return 0;
}

extern "C" int gh_cuda_set_params(gh_cuda_ctx *c, const gh_cuda_params *p)
{
  GH_REQUIRE(c && p, "gh_cuda_set_params: null argument");
  GH_CUDA_OK(cudaSetDevice(c->device));
  GH_REQUIRE(p->n_grid == c->d.n && p->n_side == c->d.nside && p->n_nu == c->d.n_nu && p->nz_tab == c->d.nz_tab &&
                 p->numk == c->d.numk && p->irregular_nutable == c->d.irregular,
             "gh_cuda_set_params: grid / sky / table sizes differ from the context's; create a new context");
  GH_REQUIRE(p->logkarr && p->pkarr && p->z_arr_r2z && p->r_arr_r2z && p->growth_d_arr && p->growth_v_arr &&
                 p->z_arr_z2r && p->r_arr_z2r && (!p->irregular_nutable || (p->nu0_arr && p->nuf_arr)),
             "gh_cuda_set_params: missing table pointer");
  if (apply_params(c, p, c->d.rank, c->d.nranks)) return 1;
  if (upload_tables(c, p)) return 1;
  if (upload_prefac(c)) return 1;
  c->sigma_overridden = false;
  return 0;
}

#define GH_CTX(c)                                             \
  GH_REQUIRE((c) != nullptr, "null gh_cuda context");         \
  GH_CUDA_OK(cudaSetDevice((c)->device))

extern "C" int gh_cuda_slab(const gh_cuda_ctx *c, int *nz_here, int *iz0_here)
{
  GH_REQUIRE(c, "null gh_cuda context");
  if (nz_here) *nz_here = c->d.nz_here;
  if (iz0_here) *iz0_here = c->d.iz0;
  return 0;
}

extern "C" int gh_cuda_shells(const gh_cuda_ctx *c, int *n_shells_here, int *shell0_here)
{
  GH_REQUIRE(c, "null gh_cuda context");
  const int per = c->d.n_nu_pad / c->d.nranks;
  const int s0 = c->d.rank * per;
  int n = c->d.n_nu - s0;
  if (n > per) n = per;
  if (n < 0) n = 0;
  if (n_shells_here) *n_shells_here = n;
  if (shell0_here) *shell0_here = s0;
  return 0;
}

extern "C" int gh_cuda_synchronize(gh_cuda_ctx *c)
{
  GH_CTX(c);
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" void *gh_cuda_stream(const gh_cuda_ctx *c) { return c ? (void *)c->stream : nullptr; }
// Opt-in (GH_TIME_FFT_PASSES=1 at context creation): device time of each field's z pass including, on several
// ranks, the transpose fused into it and the barrier that closes it.  z_ms[0] = density, z_ms[1] = potential;
// -1 when the timers are off or no FFT has run.  Synchronises the stream.
extern "C" int gh_cuda_fft_pass_times(gh_cuda_ctx *c, double *z_ms)
{
  GH_CTX(c);
  GH_REQUIRE(z_ms, "gh_cuda_fft_pass_times: null output");
  z_ms[0] = z_ms[1] = -1.0;
  if (!c->time_fft_passes) return 0;
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  for (int f = 0; f < 2; ++f) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev_pass[f][0], c->ev_pass[f][1]) == cudaSuccess) z_ms[f] = ms;
    else (void)cudaGetLastError();
  }
  return 0;
}

extern "C" unsigned long long gh_cuda_kernel_launches(const gh_cuda_ctx *c) { return c ? c->launches : 0ULL; }

extern "C" int gh_cuda_host_alloc(void **ptr, unsigned long long bytes)
{
  GH_REQUIRE(ptr, "null pointer");
  GH_CUDA_OK(cudaMallocHost(ptr, bytes));
  return 0;
}
extern "C" int gh_cuda_host_free(void *ptr)
{
  GH_CUDA_OK(cudaFreeHost(ptr));
  return 0;
}

// ------------------------------------------------------------------------------------------------ stages
extern "C" int gh_cuda_generate_k(gh_cuda_ctx *c)
{
  GH_CTX(c);
  c->sigma_ready = false;  // a new realisation: the measured variance no longer describes the density grid
  StageTimer t(c, GH_T_KGEN);
  return gh_launch_kgen(c);
}

extern "C" int gh_cuda_fft_fields(gh_cuda_ctx *c)
{
  GH_CTX(c);
  c->fft_stats_blocks = 0;
  c->sigma_ready = false;
  StageTimer t(c, GH_T_FFT);
  return gh_launch_fft_both_fields(c);  // src/fourier.c:391-392
}

extern "C" int gh_cuda_radial_velocity(gh_cuda_ctx *c)
{
  GH_CTX(c);
  StageTimer t(c, GH_T_VEL);
  return gh_launch_radial_velocity(c);
}

// variance chain, all on the device: per-CTA partials -> sums -> (all-reduce) -> mean / sigma2 in d_partials[4..5];
// the four numbers are also copied to pinned host memory for whoever asks later
static int enqueue_sigma(gh_cuda_ctx *c)
{
  StageTimer t(c, GH_T_SIGMA);
  if (gh_launch_sigma(c)) return 1;
  if (c->d.nranks > 1) GH_NCCL_OK(ncclAllReduce(c->d_partials, c->d_partials, 2, ncclDouble, ncclSum, c->comm, c->stream));
  if (gh_launch_sigma_finish(c)) return 1;
  c->sigma_ready = true;
  return 0;
}

extern "C" int gh_cuda_sigma_dens(gh_cuda_ctx *c, double *sigma2_out, double *mean_out)
{
  GH_CTX(c);
  if (enqueue_sigma(c)) return 1;
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  c->mean_gauss = c->h_stats[2];
  if (!c->sigma_overridden) c->sigma2_gauss = c->h_stats[3];
  if (sigma2_out) *sigma2_out = c->h_stats[3];
  if (mean_out) *mean_out = c->h_stats[2];
  return 0;
}

extern "C" int gh_cuda_create_d_and_vr_fields(gh_cuda_ctx *c, double *sigma2_out, double *mean_out)
{
  GH_CTX(c);
  if (!c->k_injected && gh_cuda_generate_k(c)) return 1;
  if (gh_cuda_fft_fields(c)) return 1;
  if (gh_cuda_radial_velocity(c)) return 1;
  return gh_cuda_sigma_dens(c, sigma2_out, mean_out);
}

extern "C" int gh_cuda_get_HI(gh_cuda_ctx *c)
{
  GH_CTX(c);
  c->fft_stats_blocks = 0;  // the density grid turns into HI mass
  GH_REQUIRE(c->sigma_ready || c->sigma_overridden,
             "gh_cuda_get_HI: sigma2_gauss not known for this density grid (run gh_cuda_sigma_dens / create_d_and_vr_fields, or set it)");
  StageTimer t(c, GH_T_GETHI);
  return gh_launch_get_HI(c);
}

extern "C" int gh_cuda_zero_maps(gh_cuda_ctx *c)
{
  GH_CTX(c);
  // one rank: a device->host copy of an earlier realisation may still be reading this buffer (several ranks:
  // copies read the reduce-scatter output, never the accumulation stack)
  if (c->d.nranks == 1 && c->copy_pending[c->out_cur]) GH_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_copied[c->out_cur], 0));
  GH_CUDA_OK(cudaMemsetAsync(c->maps, 0, (size_t)c->d.n_nu_pad * c->d.npix * sizeof(float), c->stream));
  return 0;
}

static int accumulate_own_slab(gh_cuda_ctx *c)
{
  return gh_launch_accumulate(c, reinterpret_cast<const float *>(c->gridA), reinterpret_cast<const float *>(c->gridC), 0, c->d.iz0,
                              c->d.nz_here);
}

// Several ranks with peer mappings: accumulate the plane range [map_bounds[rank], map_bounds[rank+1]) instead of
// the own slab.  Planes of the range that live on other ranks are copied (HI mass and Delta z_RSD, 8 B/cell) into
// the vpot slab, which is free by now, on the pull stream while the own planes are being accumulated.
// Ordering: the all-reduce barrier says every rank's get_HI has finished before anybody pulls; the reduce-scatter
// that follows the accumulation cannot complete anywhere before every rank has finished pulling, so no rank
// starts overwriting its slabs (next realisation) under a reader.
static int accumulate_rebalanced(gh_cuda_ctx *c)
{
  const GhDev &d = c->d;
  const int lo = c->map_bounds[d.rank], hi = c->map_bounds[d.rank + 1];
  const int o0 = d.iz0, o1 = d.iz0 + d.nz_here;
  const size_t plane = (size_t)d.n * 2 * d.nh;  // floats
  const float *own_m = reinterpret_cast<const float *>(c->gridA), *own_z = reinterpret_cast<const float *>(c->gridC);
  float *stage_m = reinterpret_cast<float *>(c->gridB);
  const int cap = d.nz_here / 2;                // planes of (mass, dz) pairs the vpot slab can stage
  float *stage_z = stage_m + (size_t)cap * plane;
  if (gh_stream_barrier(c)) return 1;
  GH_CUDA_OK(cudaEventRecord(c->ev_bar, c->stream));
  GH_CUDA_OK(cudaStreamWaitEvent(c->pull_stream, c->ev_bar, 0));
  const int foreign[2][2] = {{lo, hi < o0 ? hi : o0}, {lo > o1 ? lo : o1, hi}};
  // first chunk of pulls goes out before the own planes are launched, so the two overlap
  bool own_done = false, first = true;
  for (int f = 0; f < 2; ++f) {
    for (int z0 = foreign[f][0]; z0 < foreign[f][1]; z0 += cap) {
      const int z1 = z0 + cap < foreign[f][1] ? z0 + cap : foreign[f][1];
      if (!first) GH_CUDA_OK(cudaStreamWaitEvent(c->pull_stream, c->ev_chunk_free, 0));
      for (int z = z0; z < z1;) {  // one copy pair per owner
        const int q = z / d.nz_here, zq1 = (q + 1) * d.nz_here < z1 ? (q + 1) * d.nz_here : z1;
        const size_t src = (size_t)(z - q * d.nz_here) * plane, dst = (size_t)(z - z0) * plane, bytes = (size_t)(zq1 - z) * plane * sizeof(float);
        GH_CUDA_OK(cudaMemcpyAsync(stage_m + dst, reinterpret_cast<const float *>(c->peers.A[q]) + src, bytes, cudaMemcpyDefault, c->pull_stream));
        GH_CUDA_OK(cudaMemcpyAsync(stage_z + dst, reinterpret_cast<const float *>(c->peers.C[q]) + src, bytes, cudaMemcpyDefault, c->pull_stream));
        z = zq1;
      }
      GH_CUDA_OK(cudaEventRecord(c->ev_pulled, c->pull_stream));
      if (!own_done) {
        const int a = lo > o0 ? lo : o0, b = hi < o1 ? hi : o1;
        if (gh_launch_accumulate(c, own_m, own_z, a - o0, a, b - a)) return 1;
        own_done = true;
      }
      GH_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_pulled, 0));
      if (gh_launch_accumulate(c, stage_m, stage_z, 0, z0, z1 - z0)) return 1;
      GH_CUDA_OK(cudaEventRecord(c->ev_chunk_free, c->stream));
      first = false;
    }
  }
  if (!own_done) {
    const int a = lo > o0 ? lo : o0, b = hi < o1 ? hi : o1;
    if (gh_launch_accumulate(c, own_m, own_z, a - o0, a, b - a)) return 1;
  }
  return 0;
}

extern "C" int gh_cuda_accumulate_maps(gh_cuda_ctx *c)
{
  GH_CTX(c);
  StageTimer t(c, GH_T_MAPS);
  return accumulate_own_slab(c);
}

// accumulate -> (reduce-scatter) -> scale -> device->host copy on the copy stream; no host synchronisation
// device -> host copy of a finished result on the copy stream, in chunks of whole shells (>= 32 MiB, at most
// GH_MAX_CHUNKS of them), an event behind each, so that a consumer (the FITS writer) can start on the first shells
// while the rest is on the wire.  The caller has made the copy stream wait for the result.
static int issue_map_copy(gh_cuda_ctx *c, float *maps_host, const float *result, int n_here, int cur)
{
  const GhDev &d = c->d;
  GH_CUDA_OK(cudaEventRecord(c->ev[2 * GH_T_D2H], c->copy_stream));
  const size_t shell_bytes = (size_t)d.npix * sizeof(float);
  int per = (int)((((size_t)32 << 20) + shell_bytes - 1) / shell_bytes);
  if (per * GH_MAX_CHUNKS < n_here) per = (n_here + GH_MAX_CHUNKS - 1) / GH_MAX_CHUNKS;
  c->chunk_shells = per;
  c->n_chunks = 0;
  for (int s = 0; s < n_here; s += per) {
    const int ns = s + per < n_here ? per : n_here - s;
    GH_CUDA_OK(cudaMemcpyAsync(maps_host + (size_t)s * d.npix, result + (size_t)s * d.npix, (size_t)ns * shell_bytes,
                               cudaMemcpyDeviceToHost, c->copy_stream));
    if (!c->ev_chunk[c->n_chunks]) GH_CUDA_OK(cudaEventCreateWithFlags(&c->ev_chunk[c->n_chunks], cudaEventDisableTiming));
    GH_CUDA_OK(cudaEventRecord(c->ev_chunk[c->n_chunks], c->copy_stream));
    c->n_chunks++;
  }
  GH_CUDA_OK(cudaEventRecord(c->ev[2 * GH_T_D2H + 1], c->copy_stream));
  c->ev_used[GH_T_D2H] = true;
  GH_CUDA_OK(cudaEventRecord(c->ev_copied[cur], c->copy_stream));
  c->copy_pending[cur] = true;
  c->copy_enqueued[cur] = true;
  return 0;
}

// a download postponed by gh_cuda_run_async: enqueue it now, behind `after` on the compute stream if given
static int flush_deferred_copy(gh_cuda_ctx *c, bool at_accumulate)
{
  if (!c->deferred_pending) return 0;
  c->deferred_pending = false;
  if (at_accumulate) {
    if (!c->ev_acc) GH_CUDA_OK(cudaEventCreateWithFlags(&c->ev_acc, cudaEventDisableTiming));
    GH_CUDA_OK(cudaEventRecord(c->ev_acc, c->stream));
    GH_CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_acc, 0));
  }
  GH_CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_ready[c->deferred_cur], 0));
  return issue_map_copy(c, c->deferred_host, c->deferred_result, c->deferred_n, c->deferred_cur);
}

static int enqueue_maps(gh_cuda_ctx *c, float *maps_host, bool defer_copy = false)
{
  const GhDev &d = c->d;
  int n_here = 0, s0 = 0;
  gh_cuda_shells(c, &n_here, &s0);
  if (c->out_buf[1]) {  // write this realisation's result into the buffer the last copy is not reading
    c->out_cur ^= 1;
    if (d.nranks == 1) c->maps = c->out_buf[c->out_cur];
    else c->maps_recv = c->out_buf[c->out_cur];
  }
  const int cur = c->out_cur;
  if (d.nranks == 1) {
    StageTimer t(c, GH_T_MAPS);
    if (gh_cuda_zero_maps(c)) return 1;
    if (flush_deferred_copy(c, true)) return 1;  // the previous realisation's download rides along with this accumulation
    if (accumulate_own_slab(c)) return 1;
    if (gh_launch_scale_maps(c, c->maps, 0, d.n_nu)) return 1;
  } else {
    {
      StageTimer t(c, GH_T_MAPS);
      if (gh_cuda_zero_maps(c)) return 1;
      if (flush_deferred_copy(c, true)) return 1;
      if ((c->rebalance && d.nz_here >= 2) ? accumulate_rebalanced(c) : accumulate_own_slab(c)) return 1;
    }
    {
      // the reference sums full per-rank stacks onto rank 0 (src/pixelize.c:266-284); here every rank ends
      // up with the sum of its own shells, then scales just those
      StageTimer t(c, GH_T_REDUCE);
      const size_t per = (size_t)(d.n_nu_pad / d.nranks) * d.npix;
      if (c->sparse_reduce) {
        // own intervals -> all-gather (doubles as "everybody has finished accumulating") -> pull, sum, scale
        int *own = c->d_ext, *all = c->d_ext + 2 * (size_t)d.n_nu_pad;
        GH_CUDA_OK(cudaMemsetAsync(own, 0x7f, sizeof(int) * d.n_nu_pad, c->stream));               // lo = huge
        GH_CUDA_OK(cudaMemsetAsync(own + d.n_nu_pad, 0, sizeof(int) * d.n_nu_pad, c->stream));     // hi = 0: empty
        if (gh_launch_shell_extents(c, own, own + d.n_nu_pad)) return 1;
        GH_NCCL_OK(ncclAllGather(own, all, 2 * (size_t)d.n_nu_pad, ncclInt, c->comm, c->stream));
        if (c->copy_pending[cur]) GH_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_copied[cur], 0));
        if (gh_launch_sparse_reduce(c, all, c->maps_recv, s0, n_here)) return 1;
        // nobody may zero its stack (next realisation) while a peer is still reading it
        if (gh_stream_barrier(c)) return 1;
      } else {
        if (c->copy_pending[cur]) GH_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_copied[cur], 0));
        GH_NCCL_OK(ncclReduceScatter(c->maps, c->maps_recv, per, ncclFloat, ncclSum, c->comm, c->stream));
        if (gh_launch_scale_maps(c, c->maps_recv, s0, n_here)) return 1;
      }
    }
  }
  float *result = c->out_buf[cur];
  c->n_chunks = 0;
  if (maps_host && n_here > 0) {
    GH_CUDA_OK(cudaEventRecord(c->ev_done, c->stream));
    if (defer_copy && c->out_buf[1]) {  // (needs the second result buffer: the next realisation must not touch this one)
      // gh_cuda_run_async: the download starts when the NEXT realisation reaches its map accumulation (or at
      // gh_cuda_wait): next to the memory-bound FFT passes it costs them ~0.4 ms at 512^3, next to the issue-bound
      // accumulation much less
      if (!c->ev_ready[cur]) GH_CUDA_OK(cudaEventCreateWithFlags(&c->ev_ready[cur], cudaEventDisableTiming));
      GH_CUDA_OK(cudaEventRecord(c->ev_ready[cur], c->stream));
      c->deferred_host = maps_host; c->deferred_result = result; c->deferred_n = n_here; c->deferred_cur = cur;
      c->deferred_pending = true;
      c->copy_pending[cur] = true;  // nobody may zero this buffer before its copy, which is not even enqueued yet
      c->copy_enqueued[cur] = false;
      return 0;
    }
    GH_CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_done, 0));
    if (issue_map_copy(c, maps_host, result, n_here, cur)) return 1;
  }
  return 0;
}

extern "C" int gh_cuda_mk_T_maps_begin(gh_cuda_ctx *c, float *maps_host)
{
  GH_CTX(c);
  GH_REQUIRE(maps_host, "gh_cuda_mk_T_maps_begin: null host buffer");
  return enqueue_maps(c, maps_host);
}

extern "C" int gh_cuda_wait_shells(gh_cuda_ctx *c, int n_shells)
{
  GH_CTX(c);
  int n_here = 0;
  gh_cuda_shells(c, &n_here, nullptr);
  if (n_shells < 0 || n_shells > n_here) n_shells = n_here;
  if (n_shells == 0 || c->n_chunks == 0) return 0;
  int idx = (n_shells + c->chunk_shells - 1) / c->chunk_shells - 1;
  if (idx >= c->n_chunks) idx = c->n_chunks - 1;
  GH_CUDA_OK(cudaEventSynchronize(c->ev_chunk[idx]));
  return 0;
}

extern "C" int gh_cuda_wait(gh_cuda_ctx *c, double *sigma2_out)
{
  GH_CTX(c);
  if (flush_deferred_copy(c, false)) return 1;
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  for (int b = 0; b < 2; ++b) {
    if (c->copy_pending[b]) GH_CUDA_OK(cudaEventSynchronize(c->ev_copied[b]));
    c->copy_pending[b] = false;
  }
  if (c->sigma_ready) {
    c->mean_gauss = c->h_stats[2];
    if (!c->sigma_overridden) c->sigma2_gauss = c->h_stats[3];
    if (sigma2_out) *sigma2_out = c->h_stats[3];
  }
  return 0;
}

extern "C" int gh_cuda_mk_T_maps(gh_cuda_ctx *c, float *maps_host)
{
  GH_CTX(c);
  if (enqueue_maps(c, maps_host)) return 1;
  return gh_cuda_wait(c, nullptr);
}

extern "C" int gh_cuda_run_async(gh_cuda_ctx *c, float *maps_host)
{
  GH_CTX(c);
  if (!c->k_injected && gh_cuda_generate_k(c)) return 1;
  if (gh_cuda_fft_fields(c)) return 1;
  if (c->fuse_vel) {
    // nobody reads the radial velocity between the stages here: one pass does velocity + get_HI
    if (enqueue_sigma(c)) return 1;
    {
      StageTimer t(c, GH_T_VEL);  // what is left of the stage: the two-plane halo exchange
      if (gh_launch_halo_exchange(c)) return 1;
    }
    c->fft_stats_blocks = 0;
    StageTimer t(c, GH_T_GETHI);
    if (gh_launch_velocity_get_HI(c)) return 1;
  } else {
    if (gh_cuda_radial_velocity(c)) return 1;
    if (enqueue_sigma(c)) return 1;
    if (gh_cuda_get_HI(c)) return 1;
  }
  return enqueue_maps(c, maps_host, c->defer_d2h);
}

extern "C" int gh_cuda_run(gh_cuda_ctx *c, double *sigma2_out, float *maps_host)
{
  if (gh_cuda_run_async(c, maps_host)) return 1;
  return gh_cuda_wait(c, sigma2_out);
}

// ------------------------------------------------------------------------------------------------ injection / read-back
static float2 *grid_ptr(gh_cuda_ctx *c, int which)
{
  return which == GH_GRID_DENS ? c->gridA : which == GH_GRID_VPOT ? c->gridB : which == GH_GRID_RVEL ? c->gridC : nullptr;
}

extern "C" int gh_cuda_set_delta_k(gh_cuda_ctx *c, const float *dens_k, const float *vpot_k)
{
  GH_CTX(c);
  GH_REQUIRE(dens_k && vpot_k, "gh_cuda_set_delta_k: null field");
  const GhDev &d = c->d;
  // global [kz][ky][kx] -> local [kz][ky_local][kx]: per kz one contiguous run of nky_here*nh modes
  const size_t width = (size_t)d.nky_here * d.nh * sizeof(float2), spitch = (size_t)d.n * d.nh * sizeof(float2);
  const size_t off = (size_t)d.ky0 * d.nh * 2;  // floats
  GH_CUDA_OK(cudaMemcpy2DAsync(c->gridA, width, dens_k + off, spitch, width, d.n, cudaMemcpyHostToDevice, c->stream));
  GH_CUDA_OK(cudaMemcpy2DAsync(c->gridB, width, vpot_k + off, spitch, width, d.n, cudaMemcpyHostToDevice, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  c->k_injected = true;
  c->sigma_ready = false;
  return 0;
}

extern "C" int gh_cuda_clear_delta_k(gh_cuda_ctx *c)
{
  GH_REQUIRE(c, "null gh_cuda context");
  c->k_injected = false;
  return 0;
}

extern "C" int gh_cuda_download_delta_k(gh_cuda_ctx *c, float *dens_k, float *vpot_k)
{
  GH_CTX(c);
  const GhDev &d = c->d;
  const size_t width = (size_t)d.nky_here * d.nh * sizeof(float2), dpitch = (size_t)d.n * d.nh * sizeof(float2);
  const size_t off = (size_t)d.ky0 * d.nh * 2;
  if (dens_k) GH_CUDA_OK(cudaMemcpy2DAsync(dens_k + off, dpitch, c->gridA, width, width, d.n, cudaMemcpyDeviceToHost, c->stream));
  if (vpot_k) GH_CUDA_OK(cudaMemcpy2DAsync(vpot_k + off, dpitch, c->gridB, width, width, d.n, cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gh_cuda_download_grid(gh_cuda_ctx *c, int which, float *slab_out)
{
  GH_CTX(c);
  float2 *g = grid_ptr(c, which);
  GH_REQUIRE(g && slab_out, "gh_cuda_download_grid: bad grid id %d or null output", which);
  GH_CUDA_OK(cudaMemcpyAsync(slab_out, g, c->slab_complex * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gh_cuda_upload_grid(gh_cuda_ctx *c, int which, const float *slab_in)
{
  GH_CTX(c);
  float2 *g = grid_ptr(c, which);
  GH_REQUIRE(g && slab_in, "gh_cuda_upload_grid: bad grid id %d or null input", which);
  if (which == GH_GRID_DENS) { c->fft_stats_blocks = 0; c->sigma_ready = false; }  // the FFT's sums / the variance no longer describe this grid
  GH_CUDA_OK(cudaMemcpyAsync(g, slab_in, c->slab_complex * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gh_cuda_grid_checksum(gh_cuda_ctx *c, int which, int z0_local, int n_planes, unsigned long long *sum_out)
{
  GH_CTX(c);
  float2 *g = grid_ptr(c, which);
  GH_REQUIRE(g && sum_out, "gh_cuda_grid_checksum: bad grid id %d or null output", which);
  GH_REQUIRE(z0_local >= 0 && n_planes >= 0 && z0_local + n_planes <= c->d.nz_here, "gh_cuda_grid_checksum: planes [%d,%d) outside the slab",
             z0_local, z0_local + n_planes);
  *sum_out = 0ULL;
  if (n_planes == 0) return 0;
  unsigned long long *d_sum = reinterpret_cast<unsigned long long *>(c->d_partials + 7);  // a slot no kernel of the path uses
  GH_CUDA_OK(cudaMemsetAsync(d_sum, 0, sizeof(*d_sum), c->stream));
  if (gh_launch_checksum(c, reinterpret_cast<const float *>(g), z0_local, n_planes, d_sum)) return 1;
  GH_CUDA_OK(cudaMemcpyAsync(sum_out, d_sum, sizeof(*d_sum), cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gh_cuda_set_sigma2_gauss(gh_cuda_ctx *c, double sigma2)
{
  GH_REQUIRE(c, "null gh_cuda context");
  c->sigma2_gauss = sigma2;
  c->sigma_overridden = true;  // get_HI reads d_partials[6] from now on (until gh_cuda_set_params)
  GH_CUDA_OK(cudaSetDevice(c->device));
  GH_CUDA_OK(cudaMemcpyAsync(c->d_partials + 6, &c->sigma2_gauss, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gh_cuda_download_maps(gh_cuda_ctx *c, float *maps_out, unsigned long long first, unsigned long long n_floats)
{
  GH_CTX(c);
  const unsigned long long total = (unsigned long long)c->d.n_nu_pad * c->d.npix;
  GH_REQUIRE(maps_out && first + n_floats <= total, "gh_cuda_download_maps: range out of bounds");
  GH_CUDA_OK(cudaMemcpyAsync(maps_out, c->maps + first, n_floats * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gh_cuda_subparticle_offsets(const gh_cuda_ctx *c, double *xyz_out)
{
  GH_REQUIRE(c && xyz_out, "null argument");
  memcpy(xyz_out, c->d.sub_off, sizeof(c->d.sub_off));
  return 0;
}

extern "C" int gh_cuda_points_to_shell_pixel(gh_cuda_ctx *c, const double *pos, const double *dz_rsd, long long n,
                                             int *shell_out, long long *pix_out)
{
  GH_CTX(c);
  GH_REQUIRE(pos && shell_out && pix_out && n >= 0, "gh_cuda_points_to_shell_pixel: bad argument");
  if (n == 0) return 0;
  double *d_pos = nullptr, *d_dz = nullptr;
  int *d_sh = nullptr;
  long long *d_px = nullptr;
  int rc = 1;
  do {
    if (cudaMalloc(&d_pos, sizeof(double) * 3 * n) != cudaSuccess) break;
    if (dz_rsd && cudaMalloc(&d_dz, sizeof(double) * n) != cudaSuccess) break;
    if (cudaMalloc(&d_sh, sizeof(int) * n) != cudaSuccess) break;
    if (cudaMalloc(&d_px, sizeof(long long) * n) != cudaSuccess) break;
    if (cudaMemcpyAsync(d_pos, pos, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) break;
    if (dz_rsd && cudaMemcpyAsync(d_dz, dz_rsd, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) break;
    if (gh_launch_points(c, d_pos, d_dz, n, d_sh, d_px)) break;
    if (cudaMemcpyAsync(shell_out, d_sh, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) break;
    if (cudaMemcpyAsync(pix_out, d_px, sizeof(long long) * n, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) break;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) break;
    rc = 0;
  } while (0);
  if (rc) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) gh_set_error("gh_cuda_points_to_shell_pixel: %s", cudaGetErrorString(e));
  }
  cudaFree(d_pos); cudaFree(d_dz); cudaFree(d_sh); cudaFree(d_px);
  return rc;
}

extern "C" int gh_cuda_fastpath_audit(gh_cuda_ctx *c, const double *pos, const double *dz_rsd, long long n, double eps_scale,
                                      unsigned long long *counts_out)
{
  GH_CTX(c);
  GH_REQUIRE(pos && counts_out && n >= 0, "gh_cuda_fastpath_audit: bad argument");
  memset(counts_out, 0, 4 * sizeof(unsigned long long));
  if (n == 0) return 0;
  double *d_pos = nullptr, *d_dz = nullptr;
  unsigned long long *d_cnt = nullptr;
  int rc = 1;
  do {
    if (cudaMalloc(&d_pos, sizeof(double) * 3 * n) != cudaSuccess) break;
    if (dz_rsd && cudaMalloc(&d_dz, sizeof(double) * n) != cudaSuccess) break;
    if (cudaMalloc(&d_cnt, 4 * sizeof(unsigned long long)) != cudaSuccess) break;
    if (cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned long long), c->stream) != cudaSuccess) break;
    if (cudaMemcpyAsync(d_pos, pos, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) break;
    if (dz_rsd && cudaMemcpyAsync(d_dz, dz_rsd, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) break;
    if (gh_launch_fastpath_audit(c, d_pos, d_dz, n, (float)eps_scale, d_cnt)) break;
    if (cudaMemcpyAsync(counts_out, d_cnt, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) break;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) break;
    rc = 0;
  } while (0);
  if (rc) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) gh_set_error("gh_cuda_fastpath_audit: %s", cudaGetErrorString(e));
  }
  cudaFree(d_pos); cudaFree(d_dz); cudaFree(d_cnt);
  return rc;
}

extern "C" int gh_cuda_accumulate_audit(gh_cuda_ctx *c, double eps_scale, unsigned long long *counts_out)
{
  GH_CTX(c);
  GH_REQUIRE(counts_out, "gh_cuda_accumulate_audit: null output");
  unsigned long long *d_cnt = nullptr;
  GH_CUDA_OK(cudaMalloc(&d_cnt, 4 * sizeof(unsigned long long)));
  int rc = 1;
  do {
    if (cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned long long), c->stream) != cudaSuccess) break;
    if (gh_launch_accumulate_audit(c, (float)eps_scale, d_cnt)) break;
    if (cudaMemcpyAsync(counts_out, d_cnt, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) break;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) break;
    rc = 0;
  } while (0);
  if (rc) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) gh_set_error("gh_cuda_accumulate_audit: %s", cudaGetErrorString(e));
  }
  cudaFree(d_cnt);
  return rc;
}

extern "C" int gh_cuda_stage_times(gh_cuda_ctx *c, double *ms_out)
{
  GH_CTX(c);
  GH_REQUIRE(ms_out, "null output");
  GH_CUDA_OK(cudaStreamSynchronize(c->stream));
  for (int s = 0; s < GH_T_NSLOTS; ++s) {
    ms_out[s] = 0;
    if (!c->ev_used[s]) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev[2 * s], c->ev[2 * s + 1]) == cudaSuccess) ms_out[s] = ms;
  }
  return 0;
}
