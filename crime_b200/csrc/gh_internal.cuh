// Internal declarations shared by the translation units of libgh_cuda.so.
// B200 (sm_100a) only; no CPU fallback exists anywhere in this library.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/gh_cuda.h"

#define GH_NZ_TAB_MAX 8192

// Everything a kernel needs, passed by value (small tables live in device memory).
struct GhDev {
  int n, nh;           // n_grid, n_grid/2+1
  int nz_here, iz0;    // real-space slab (z planes) of this rank
  int nky_here, ky0;   // k-space slab (ky rows) of this rank, before the transpose
  int nranks, rank;
  double l_box, dx, pos_obs[3];
  float half_inv_dx;   // 0.5/dx
  // k-space realisation
  unsigned int seed;
  int do_smoothing;
  double r2_smooth, vfactor, dk, idk3;
  int numk;
  double logkmin, logkmax, idlogk, n_scal;
  const double *logkarr, *pkarr;
  // radial tables
  int nz_tab;
  double glob_idr, r_tab_max;
  const double *z_r2z, *r_r2z, *gd, *gv;  // double, for the exact index path
  const float *z_r2z_f, *gd_f, *gv_f;     // float copies for the field kernels
  const float *frac_f, *bias_f;           // fraction_HI / bias_HI on the same radial grid (src/user_defined.c:27-35)
  // sky
  long long nside, npix;
  int n_nu, n_nu_pad, irregular;
  const double *nu0, *nuf;
  const float *nu_edges_f;  // n_nu+1 float shell edges for the fp32 fast path
  const float *r_z2r_f;     // float copy of r_arr_z2r (uniform in z, step dz_tab)
  float inv_dz_tab, z_tab_max;
  float rz_slope_var;       // largest change of dr/dz between adjacent intervals of r_z2r_f (block shell thresholds)
  double nu_min, nu_max, inv_dnu;
  double z_lo_cull, z_hi_cull; // redshift window outside of which no sub-particle can land in a shell
  double sub_off[3 * GH_CUDA_N_SUBPART];
  float sub_off_f[3 * GH_CUDA_N_SUBPART];
  // (ox, oy, oz, |o|^2) of each float offset: one 16-byte constant load per sub-particle in the block-expansion loops
  float4 sub_c[GH_CUDA_N_SUBPART];
};

// d_partials layout (doubles): [0..1] sum, sum of squares (all-reduced); [4] mean; [5] measured variance;
// [6] caller-supplied variance (gh_cuda_set_sigma2_gauss); per-CTA partial sums start at GH_PARTIALS_BASE, so no
// partial-sum kernel ever touches the slots get_HI reads
#define GH_PARTIALS_BASE 8

#define GH_MAX_RANKS 16
#define GH_MAX_CHUNKS 64
#define GH_N_COPY_STREAMS 4

// Peer views of the slab buffers of every rank (CUDA IPC mappings; entry [rank] is the local buffer).
// Kernels use them to read or write other GPUs' memory over NVLink directly.
struct GhPeers {
  float2 *A[GH_MAX_RANKS];  // dens / HI mass
  float2 *C[GH_MAX_RANKS];  // transpose receive buffer / radial velocity / Delta z_RSD
};

struct gh_cuda_ctx {
  GhDev d;
  int device;
  cudaStream_t stream;
  cudaStream_t copy_stream;        // device->host copy of the finished maps, overlaps the next realisation
  cudaEvent_t ev_done, ev_copied[2];
  bool copy_pending[2], sigma_ready;
  bool copy_enqueued[2];           // the pending copy of that buffer has been handed to the copy stream (see defer_d2h)
  bool defer_d2h;                  // gh_cuda_run_async postpones a download to the next realisation's accumulation (default; GH_NO_DEFER_D2H=1 turns it off)
  bool deferred_pending;
  float *deferred_host;
  const float *deferred_result;
  int deferred_n, deferred_cur;
  cudaEvent_t ev_ready[2], ev_acc;
  float *out_buf[2];               // what the device->host copy reads (maps on one rank, maps_recv on several), doubled when it fits
  int out_cur;
  char *h_stage[2];                // pinned staging of the tables + prefactors, used alternately
  cudaEvent_t ev_stage[2];
  int stage_next, stage_cur;
  cudaEvent_t ev_chunk[GH_MAX_CHUNKS];  // one behind every chunk of the last map download
  int n_chunks, chunk_shells;
  double *h_stats;                 // mapped pinned: sum, sumsq, mean, sigma2 of the last realisation (written by the kernel)
  double *h_stats_dev;             // its device-side address
  ncclComm_t comm;
  bool have_comm;
  bool have_peers;                 // peer mappings established (nranks>1, same node)
  bool time_fft_passes;            // opt-in (GH_TIME_FFT_PASSES=1): events around each field's z pass + transpose
  cudaEvent_t ev_pass[2][2];
  bool fuse_vel;                   // gh_cuda_run*: radial velocity and get_HI in one pass
  bool sparse_reduce;              // opt-in: map reduction by pulling the peers' touched pixel intervals (GH_SPARSE_REDUCE=1)
  float *map_peers[GH_MAX_RANKS];  // every rank's accumulation stack (peer-mapped), sparse_reduce only
  int *d_ext;                      // [2][n_nu_pad] own intervals, then [nranks][2][n_nu_pad] gathered
  bool rebalance;                  // accumulate equal-cost plane ranges, pulling foreign planes from peers
  int map_bounds[GH_MAX_RANKS + 1];
  cudaStream_t pull_stream;
  cudaEvent_t ev_bar, ev_pulled, ev_chunk_free;
  GhPeers peers;
  int *d_barrier;                  // one int, all-reduced as a stream-ordered cross-rank barrier
  size_t slab_complex;  // complex elements per slab = nz_here*n*nh (== n*nky_here*nh)
  float2 *gridA, *gridB, *gridC;  // dens, vpot, rvel/transposition scratch
  float *halo_lo, *halo_hi;       // neighbour planes of vpot (nranks>1)
  float *maps;                    // [n_nu_pad][npix] accumulation stack (== out_buf[out_cur] on one rank)
  float *maps_recv;               // reduce-scatter output (nranks>1) == out_buf[out_cur]
  float2 *twiddle;                // exp(+2 pi i j/n), j<n
  double *d_partials;             // reduction scratch
  double *d_prefac;               // per-shell mass->temperature factors
  void *d_tables;                 // one allocation holding all small tables
  size_t tables_bytes;
  double h_prefac[4096];
  double h_nu_centre[4096];       // shell centres (src/pixelize.c:63-70), for the point-source maps
  bool k_injected;
  bool sigma_overridden;
  double sigma2_gauss, mean_gauss;
  cudaEvent_t ev[2 * GH_T_NSLOTS];
  bool ev_used[GH_T_NSLOTS];
  unsigned long long launches;
  int n_sm;
  int fft_stats_blocks;    // >0: the density FFT left that many per-CTA (sum, sumsq) partials in d_partials
  size_t fft_batch_bytes;  // plane batch of the fused y/x FFT passes (kept L2-resident)
  // transposes on the copy engines, pipelined against the other field's passes (several ranks; gh_fft.cu)
  bool ce_transpose;       // GH_FUSED_TRANSPOSE=1 restores the transpose fused into the z pass
  bool nccl_transpose;     // the pipelined transposes as ncclSend/ncclRecv groups on a second communicator (GH_TRANSPOSE=nccl)
  bool push_transpose;     // the pipelined transposes as a store kernel on a high-priority stream (GH_TRANSPOSE=push)
  int push_ctas;
  ncclComm_t comm2;        // used by the transposes only, on ce_stream[0]: never in flight together with `comm`
  bool have_comm2;
  cudaStream_t ce_stream[GH_N_COPY_STREAMS];
  int ce_streams_used;     // GH_CE_STREAMS=1..4: how many of them the peer copies are spread over
  cudaEvent_t ev_z[2], ev_free[2], ev_sent[2][GH_N_COPY_STREAMS];
  float2 *recv2;           // second receive buffer (the idle map stack, or an extra slab), nullptr: none
  float2 *recv2_peers[GH_MAX_RANKS];
  bool recv2_owned;        // recv2 is an allocation of its own
  bool fft_tma;            // strided FFT passes fetch their tiles with the TMA unit (GH_FFT_NO_TMA=1 turns it off)
  CUtensorMap fft_map[6];  // z pass of A, of B; y pass of A, of B: even rows; odd rows (wider box, see gh_fft.cu)
  bool fft_map_ok[4];
};

void gh_set_error(const char *fmt, ...);

#define GH_CUDA_OK(call)                                                                          \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      gh_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

#define GH_NCCL_OK(call)                                                                          \
  do {                                                                                            \
    ncclResult_t r__ = (call);                                                                    \
    if (r__ != ncclSuccess) {                                                                     \
      gh_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__));   \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

#define GH_LAUNCH_CHECK(ctx)                                                                      \
  do {                                                                                            \
    (ctx)->launches++;                                                                            \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess) {                                                                     \
      gh_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

// stage launchers (each enqueues on ctx->stream and returns 0 / non-zero)
int gh_launch_kgen(gh_cuda_ctx *c);
int gh_launch_fft_field(gh_cuda_ctx *c, float2 *field);  // full c2r of one field incl. transpose + normalisation
int gh_launch_fft_both_fields(gh_cuda_ctx *c);           // density then potential; pipelined transposes on several ranks
int gh_fft_supported(int n);  // any even n_grid in 8..4096
int gh_fft_tuned(int n);      // powers of two 32..4096: tuned kernels, any power-of-two rank count
int gh_launch_radial_velocity(gh_cuda_ctx *c);
int gh_launch_sigma(gh_cuda_ctx *c);  // leaves (sum, sumsq) in c->d_partials[0..1]
int gh_launch_sigma_finish(gh_cuda_ctx *c);  // d_partials[4] = mean, [5] = measured sigma2_gauss
int gh_launch_get_HI(gh_cuda_ctx *c);
int gh_launch_halo_exchange(gh_cuda_ctx *c);
int gh_launch_checksum(gh_cuda_ctx *c, const float *grid, int z0_local, int nplanes, unsigned long long *d_out);
int gh_launch_velocity_get_HI(gh_cuda_ctx *c);
int gh_launch_accumulate(gh_cuda_ctx *c, const float *mass, const float *dzrsd, int iz_base, int zg_base, int nplanes);
int gh_stream_barrier(gh_cuda_ctx *c);  // every rank has reached this point of its stream
int gh_launch_accumulate_audit(gh_cuda_ctx *c, float eps_scale, unsigned long long *d_counts);
int gh_launch_scale_maps(gh_cuda_ctx *c, float *maps, int shell0, int nshells);
int gh_launch_shell_extents(gh_cuda_ctx *c, int *ext_lo, int *ext_hi);
int gh_launch_sparse_reduce(gh_cuda_ctx *c, const int *all_ext, float *out, int shell0, int nshells);
int gh_launch_fastpath_audit(gh_cuda_ctx *c, const double *d_pos, const double *d_dz, long long n, float eps_scale,
                              unsigned long long *d_counts);
void gh_psources_release(gh_cuda_ctx *c);  // frees the point-source state of a context (gh_psources.cu)
int gh_launch_points(gh_cuda_ctx *c, const double *d_pos, const double *d_dz, long long n, int *d_shell,
                     long long *d_pix);
