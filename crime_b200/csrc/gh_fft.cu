// 3-D complex-to-real inverse FFT of one half-spectrum field, hand-written for sm_100a.
//
// Replaces the reference's fftw_wrap (src/fourier.c:78-99, FFTW c2r 3-D, unnormalised, sign +) and the
// normalisation loop that follows it (src/fourier.c:397-412), slab-decomposed like FFTW-MPI
// (src/fourier.c:141-148) but with ONE transpose per field: k-space is held ky-distributed
// ([kz][ky_local][kx]) so the z transform is local, a single all-to-all turns it into z-distributed
// planes ([z_local][ky][kx]) and the y transform + x half-complex-to-real transform finish locally.
//
// Kernels (all HBM-bound; 8 B/mode read + 8 B/mode written per pass, i.e. 8 B per real cell):
//   fft_strided_kernel : radix-8/4 decimation-in-frequency over a strided axis.  One CTA owns a tile of
//                        W adjacent lines (W*8 B contiguous per element row -> full 32 B sectors), keeps
//                        it in shared memory between radix passes, first pass straight from global
//                        into registers, last pass straight to global (digit-reversed scatter).
//   fft_c2r_rows_kernel: x axis, rows are contiguous.  Half-size complex transform with the Hermitian
//                        pre-twist fused into the load (Im of DC and Nyquist are never read, as FFTW's
//                        c2r); a tile of W rows is transposed into the same [position][line] shared
//                        layout, transformed in place, and gathered back in natural order so that both
//                        the global loads and the global stores are coalesced; in place, scaled by
//                        (sqrt(2 pi)/L)^3.
// The row kernel addresses shared memory through an XOR swizzle of the 16 float2-banks chosen (by brute force
// over every access phase) so that staging, radix passes and gather all run at the 2-wavefront optimum of
// 8-byte accesses; the strided kernel's [position][line] tile is conflict-free as it is (lines are the fast
// index in both global and shared memory), so it uses plain addressing.
#include "gh_internal.cuh"
#include "gh_fft_generic.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace {

// ---- radix plan: 8s first, then 4s (n = 2^e, e >= 4) -------------------------------------------
__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }
__host__ __device__ constexpr int n_eights(int n) { return (ilog2(n) % 3 == 1) ? ilog2(n) / 3 - 1 : ilog2(n) / 3; }
__host__ __device__ constexpr int n_fours(int n) { return (ilog2(n) % 3 == 1) ? 2 : (ilog2(n) % 3 == 2 ? 1 : 0); }
__host__ __device__ constexpr int n_steps(int n) { return n_eights(n) + n_fours(n); }
__host__ __device__ constexpr int rad_at(int n, int s) { return s < n_eights(n) ? 8 : 4; }
// product of radices of steps [0, s)
__host__ __device__ constexpr int rad_prod(int n, int s) { return s <= 0 ? 1 : rad_prod(n, s - 1) * rad_at(n, s - 1); }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }  // * (+i)

// in-register DFT, exponent sign +
__device__ __forceinline__ void dft4(float2 &u0, float2 &u1, float2 &u2, float2 &u3)
{
  float2 a = cadd(u0, u2), b = csub(u0, u2), c = cadd(u1, u3), d = mul_pi(csub(u1, u3));
  u0 = cadd(a, c);
  u1 = cadd(b, d);
  u2 = csub(a, c);
  u3 = csub(b, d);
}

template <int R> __device__ __forceinline__ void dft(float2 (&u)[R]);
template <> __device__ __forceinline__ void dft<4>(float2 (&u)[4]) { dft4(u[0], u[1], u[2], u[3]); }
template <> __device__ __forceinline__ void dft<8>(float2 (&u)[8])
{
  dft4(u[0], u[2], u[4], u[6]);  // even samples -> E[0..3] in u0,u2,u4,u6
  dft4(u[1], u[3], u[5], u[7]);  // odd samples  -> O[0..3] in u1,u3,u5,u7
  const float h = 0.70710678118654752440f;
  float2 o0 = u[1];
  float2 o1 = make_float2((u[3].x - u[3].y) * h, (u[3].x + u[3].y) * h);   // * (1+i)/sqrt2
  float2 o2 = mul_pi(u[5]);                                                // * i
  float2 o3 = make_float2((-u[7].x - u[7].y) * h, (u[7].x - u[7].y) * h);  // * (-1+i)/sqrt2
  float2 e0 = u[0], e1 = u[2], e2 = u[4], e3 = u[6];
  u[0] = cadd(e0, o0); u[4] = csub(e0, o0);
  u[1] = cadd(e1, o1); u[5] = csub(e1, o1);
  u[2] = cadd(e2, o2); u[6] = csub(e2, o2);
  u[3] = cadd(e3, o3); u[7] = csub(e3, o3);
}

// u[q] *= w^q, q = 1..R-1
template <int R> __device__ __forceinline__ void twiddle_powers(float2 (&u)[R], float2 w1)
{
  float2 w2 = cmul(w1, w1);
  u[1] = cmul(u[1], w1);
  u[2] = cmul(u[2], w2);
  float2 w3 = cmul(w2, w1);
  u[3] = cmul(u[3], w3);
  if constexpr (R == 8) {
    float2 w4 = cmul(w2, w2);
    u[4] = cmul(u[4], w4);
    u[5] = cmul(u[5], cmul(w4, w1));
    u[6] = cmul(u[6], cmul(w3, w3));
    u[7] = cmul(u[7], cmul(w4, w3));
  }
}

// position in the decimation-in-frequency output <-> natural index (mixed-radix digit reversal):
// natural f = q0 + R0*(q1 + R1*(q2 ...)),  position = q0*(n/R0) + q1*(n/(R0 R1)) + ...
// Template recursion keeps every divisor a compile-time constant (shifts and masks in SASS).
template <int N, int S = 0> __device__ __forceinline__ int dif_pos_to_freq(int pos)
{
  if constexpr (S >= n_steps(N)) {
    return 0;
  } else {
    constexpr int sub = N / rad_prod(N, S + 1), R = rad_at(N, S), mult = rad_prod(N, S);
    return ((pos / sub) % R) * mult + dif_pos_to_freq<N, S + 1>(pos);
  }
}
template <int N, int S = 0> __device__ __forceinline__ int freq_to_dif_pos(int f)
{
  if constexpr (S >= n_steps(N)) {
    return 0;
  } else {
    constexpr int sub = N / rad_prod(N, S + 1), R = rad_at(N, S), mult = rad_prod(N, S);
    return ((f / mult) % R) * sub + freq_to_dif_pos<N, S + 1>(f);
  }
}

// ---- shared-memory tile [position][line], XOR-swizzled over the 16 float2 banks ---------------------
// address a = pos*W + w; its 16-element line index l = a >> 4 picks an XOR mask for the bank bits.
// Shift triples found by exhaustive simulation of every access phase (stage / radix passes / gather) with
// the half-warp conflict model of 8-byte accesses; 31 = unused.
template <int LEN, int W> struct Swz { static constexpr int a = 0, b = 1, c = 31; };
template <> struct Swz<64, 16> { static constexpr int a = 0, b = 2, c = 31; };
template <> struct Swz<128, 16> { static constexpr int a = 0, b = 1, c = 5; };
template <> struct Swz<256, 16> { static constexpr int a = 0, b = 2, c = 4; };
template <> struct Swz<512, 16> { static constexpr int a = 0, b = 6, c = 31; };
template <> struct Swz<1024, 16> { static constexpr int a = 0, b = 1, c = 7; };
template <> struct Swz<1024, 8> { static constexpr int a = 0, b = 6, c = 31; };
template <> struct Swz<2048, 8> { static constexpr int a = 0, b = 1, c = 7; };
template <> struct Swz<2048, 4> { static constexpr int a = 0, b = 6, c = 31; };

// tile layouts: LAY_PLAIN [position][line]; LAY_SWZ the same, XOR-swizzled; LAY_PAIR [position parity][position / 2][line],
// which is how the TMA-fed y pass receives its tile (even and odd rows arrive as two boxes, see fft_strided_tma_kernel)
// LAY_PAD: [position][line] with a pitch of W + 1 float2 (odd pitch -> conflict-free along either index for one multiply-add
// instead of the XOR swizzle's six integer operations per address).  Measured slower than the swizzle in the x pass
// (512^3: FFT 1.65 vs 1.56 ms, 1024^3: 18.1 vs 16.3 ms; profiles/r2/ab_fft_rows_pad_*.log), so off: -DGH_FFT_ROWS_PAD=1 builds it
enum { LAY_PLAIN = 0, LAY_SWZ = 1, LAY_PAIR = 2, LAY_PAD = 3 };
#ifndef GH_FFT_ROWS_PAD
#define GH_FFT_ROWS_PAD 0
#endif
template <int W> struct RowsLay { static constexpr int value = (W == 16 && GH_FFT_ROWS_PAD) ? LAY_PAD : LAY_SWZ; };

template <int LEN, int W, int LAY = LAY_SWZ> __device__ __forceinline__ int phys(int pos, int w)
{
  using S = Swz<LEN, W>;
  // odd rows: pitch W + 2, data from column 1 (their box starts one mode early to be 16-byte aligned)
  if constexpr (LAY == LAY_PAD) return pos * (W + 1) + w;
  if constexpr (LAY == LAY_PAIR) return (pos & 1) ? (LEN / 2) * W + (pos >> 1) * (W + 2) + 1 + w : (pos >> 1) * W + w;
  const int a = pos * W + w;
  if constexpr (LAY == LAY_PLAIN) return a;
  const int l = a >> 4;
  int m = l >> S::a;
  if constexpr (S::b < 31) m ^= l >> S::b;
  if constexpr (S::c < 31) m ^= l >> S::c;
  return a ^ (m & 15);
}

// One in-place radix pass of a LEN-point decimation-in-frequency transform over a [LEN][W] tile.
// NTW is the length of the twiddle table (exp(+2 pi i j/NTW)); LEN divides NTW.
template <int LEN, int NTW, int W, int NT, int S, int LAY>
__device__ __forceinline__ void dif_pass_smem(float2 *sm, const float2 *__restrict__ tw)
{
  constexpr int R = rad_at(LEN, S);
  constexpr int M = LEN / rad_prod(LEN, S);
  constexpr int SUB = M / R;
#pragma unroll 1
  for (int item = threadIdx.x; item < W * (LEN / R); item += NT) {
    const int w = item % W, ii = item / W;
    const int b = ii / SUB, i = ii % SUB;
    const int p0 = b * M + i;
    int addr[R];
    float2 u[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      addr[r] = phys<LEN, W, LAY>(p0 + r * SUB, w);
      u[r] = sm[addr[r]];
    }
    dft<R>(u);
    if constexpr (SUB > 1) twiddle_powers<R>(u, __ldg(tw + i * (NTW / M)));
#pragma unroll
    for (int q = 0; q < R; ++q) sm[addr[q]] = u[q];
  }
}

// ---- strided axis -------------------------------------------------------------------------------
struct StridedGeom {
  int lines_per_group;      // lines sharing one base offset rule (a plane's kx columns, or all columns flat)
  int tiles_per_group;
  long long src_group_stride, dst_group_stride;
  long long src_stride, dst_stride;  // element stride along the transformed axis
  int grp0;                 // first group handled by this launch (plane batches)
  int blk_shift;            // source index n is split as (n >> blk_shift, n & mask): the received
  long long src_chunk;      // all-to-all blocks sit src_chunk apart (blk_shift = 30 disables the split)
  // fused transpose (z pass, nranks > 1): output plane f belongs to rank f / nz_peer and is stored straight into
  // that rank's receive buffer over NVLink, at block `me` of its [peer][z_local][ky_local][kx] layout
  int peer_mode, nz_peer, me;
  long long peer_chunk;
};

template <int N, int W, int NT, int S, int LAY = LAY_PLAIN>
__device__ __forceinline__ void strided_middle(float2 *sm, const float2 *__restrict__ tw)
{
  if constexpr (S < n_steps(N) - 1) {
    dif_pass_smem<N, N, W, NT, S, LAY>(sm, tw);
    __syncthreads();
    strided_middle<N, W, NT, S + 1, LAY>(sm, tw);
  }
}

template <int N, int W, int NT>
__global__ void __launch_bounds__(NT) fft_strided_kernel(const float2 *__restrict__ src, float2 *__restrict__ dst,
                                                         const float2 *__restrict__ tw, StridedGeom g, GhPeers peers)
{
  extern __shared__ float2 sm[];
  const int tile = blockIdx.x;
  const int grp_rel = tile / g.tiles_per_group;
  const int grp = g.grp0 + grp_rel;
  const int l0 = (tile - grp_rel * g.tiles_per_group) * W;
  const float2 *sp = src + ((long long)grp * g.src_group_stride + l0);
  float2 *dp = dst + ((long long)grp * g.dst_group_stride + l0);
  const int nvalid = min(W, g.lines_per_group - l0);
  const int blk_mask = (1 << g.blk_shift) - 1;
  const int tid = threadIdx.x;
  constexpr int NS = n_steps(N);
  {
    // first pass: global -> registers -> shared (block length N, no block offset)
    constexpr int R = rad_at(N, 0), SUB = N / R;
#pragma unroll 1
    for (int item = tid; item < W * SUB; item += NT) {
      const int w = item % W, i = item / W;
      float2 u[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int n = i + r * SUB;
        const long long a = (long long)(n >> g.blk_shift) * g.src_chunk + (long long)(n & blk_mask) * g.src_stride;
        u[r] = (w < nvalid) ? __ldg(sp + a + w) : make_float2(0.f, 0.f);
      }
      dft<R>(u);
      twiddle_powers<R>(u, __ldg(tw + i));
#pragma unroll
      for (int q = 0; q < R; ++q) sm[phys<N, W, LAY_PLAIN>(i + q * SUB, w)] = u[q];
    }
  }
  __syncthreads();
  strided_middle<N, W, NT, 1>(sm, tw);
  {
    // last pass: shared -> registers -> global, scattered to natural order (block length R, SUB = 1)
    constexpr int R = rad_at(N, NS - 1);
#pragma unroll 1
    for (int item = tid; item < W * (N / R); item += NT) {
      const int w = item % W, b = item / W;
      float2 u[R];
#pragma unroll
      for (int r = 0; r < R; ++r) u[r] = sm[phys<N, W, LAY_PLAIN>(b * R + r, w)];
      dft<R>(u);
      if (w < nvalid) {
        const int f0 = dif_pos_to_freq<N>(b * R);
        if (!g.peer_mode) {
          float2 *o = dp + (long long)f0 * g.dst_stride + w;
#pragma unroll
          for (int q = 0; q < R; ++q) o[(long long)(q * (N / R)) * g.dst_stride] = u[q];
        } else {
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const int f = f0 + q * (N / R);
            const int owner = f / g.nz_peer, zl = f - owner * g.nz_peer;
            peers.C[owner][(long long)g.me * g.peer_chunk + (long long)zl * g.dst_stride + l0 + w] = u[q];
          }
        }
      }
    }
  }
}

// ---- strided axis, tile fetched by the TMA unit ----------------------------------------------------
// Same transform as fft_strided_kernel, but the [N][W] tile arrives in shared memory through cp.async.bulk.tensor
// boxes issued by one thread and signalled on an mbarrier: no load instructions, no registers and no address
// arithmetic are spent on the 8-byte-per-thread strided reads, and rows of W * 8 = 32..128 bytes stream at the
// rate of a bulk copy (tools/tma_tile_probe.cu: 3.9-4.6 TB/s where the per-thread loads reach 2.1 at N >= 1024).
// All radix passes then run out of shared memory and the last one scatters to global memory as before (to the
// peers' receive buffers in the fused transpose).
//   PAIR = false (z pass): the slab is a 2-D tensor {line (flat ky_local * nh + kx), kz}; its row pitch lines * 8 B
//     is a multiple of 16 B because nky_here is even.
//   PAIR = true (y pass): rows of one z plane are nh = N/2 + 1 modes long, an odd number, so the 8 nh-byte row pitch
//     violates the TMA's 16-byte stride rule.  Two consecutive rows, however, form one 16 nh-byte "row pair": the
//     tensor is {4 nh floats, rows / 2 pairs, plane, chunk} and a tile takes two boxes per block of pairs: the even
//     rows at inner offset 2 kx0, W modes wide, and the odd rows through a second map whose box is W + 2 modes wide
//     and starts at 2 (nh + kx0 - 1) -- box starts must be 16-byte aligned too (an odd start faults with "illegal
//     instruction"), and nh + kx0 is odd --, so the odd rows sit one column to the right in their half of the
//     LAY_PAIR tile.  chunk = the peer blocks of the received transpose ([q][z_local][ky_local][kx]).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned phase)
{
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
               : "memory");
}

struct TmaGeom {
  int box_rows;   // rows (PAIR: row pairs) per box, <= 256
  int nchunks;    // PAIR: peer blocks along the transformed axis (1 on one rank)
  int nh;
};

template <int N, int W, int NT, bool PAIR>
__global__ void __launch_bounds__(NT) fft_strided_tma_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap map_odd,
                                                             float2 *__restrict__ dst,
                                                             const float2 *__restrict__ tw, StridedGeom g, TmaGeom tg, GhPeers peers)
{
  extern __shared__ __align__(128) float2 sm[];
  __shared__ uint64_t bar;
  constexpr int LAY = PAIR ? LAY_PAIR : LAY_PLAIN;
  const int tile = blockIdx.x;
  const int grp_rel = tile / g.tiles_per_group;
  const int grp = g.grp0 + grp_rel;
  const int l0 = (tile - grp_rel * g.tiles_per_group) * W;
  float2 *dp = dst + ((long long)grp * g.dst_group_stride + l0);
  const int nvalid = min(W, g.lines_per_group - l0);
  const int tid = threadIdx.x;
  constexpr int NS = n_steps(N);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bar, (unsigned)((N * W + (PAIR ? N : 0)) * sizeof(float2)));
    if constexpr (!PAIR) {
      for (int r = 0; r < N; r += tg.box_rows) tma_load_2d(sm + (size_t)r * W, &map, &bar, 2 * l0, r);
    } else {
      const int half = N / 2, ppc = half / tg.nchunks;  // row pairs per chunk
      for (int q = 0; q < tg.nchunks; ++q) {
        for (int j = 0; j < ppc; j += tg.box_rows) {
          tma_load_4d(sm + (size_t)(q * ppc + j) * W, &map, &bar, 2 * l0, j, grp, q);
          tma_load_4d(sm + (size_t)half * W + (size_t)(q * ppc + j) * (W + 2), &map_odd, &bar, 2 * (tg.nh + l0 - 1), j, grp, q);
        }
      }
    }
  }
  __syncthreads();  // the barrier's initialisation is visible to everybody
  while (!mbar_try_wait(&bar, 0)) {}
  strided_middle<N, W, NT, 0, LAY>(sm, tw);
  {
    // last pass: shared -> registers -> global, scattered to natural order (block length R, SUB = 1)
    constexpr int R = rad_at(N, NS - 1);
#pragma unroll 1
    for (int item = tid; item < W * (N / R); item += NT) {
      const int w = item % W, b = item / W;
      float2 u[R];
#pragma unroll
      for (int r = 0; r < R; ++r) u[r] = sm[phys<N, W, LAY>(b * R + r, w)];
      dft<R>(u);
      if (w < nvalid) {
        const int f0 = dif_pos_to_freq<N>(b * R);
        if (!g.peer_mode) {
          float2 *o = dp + (long long)f0 * g.dst_stride + w;
#pragma unroll
          for (int q = 0; q < R; ++q) o[(long long)(q * (N / R)) * g.dst_stride] = u[q];
        } else {
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const int f = f0 + q * (N / R);
            const int owner = f / g.nz_peer, zl = f - owner * g.nz_peer;
            peers.C[owner][(long long)g.me * g.peer_chunk + (long long)zl * g.dst_stride + l0 + w] = u[q];
          }
        }
      }
    }
  }
}

// ---- contiguous rows: half-complex -> real, in place ----------------------------------------------
template <int H, int NTW, int W, int NT, int S>
__device__ __forceinline__ void rows_passes(float2 *sm, const float2 *__restrict__ tw)
{
  if constexpr (S < n_steps(H)) {
    dif_pass_smem<H, NTW, W, NT, S, RowsLay<W>::value>(sm, tw);
    __syncthreads();
    rows_passes<H, NTW, W, NT, S + 1>(sm, tw);
  }
}

// STATS: also leave this CTA's sum and sum of squares of the real cells it produced in stats[GH_PARTIALS_BASE + 2*blockIdx.x ..]
// (float product, double accumulation, fixed order: compute_sigma_dens, src/fourier.c:24-76, without re-reading
// the density field).
template <int N, int W, int NT, bool STATS>
__global__ void __launch_bounds__(NT) fft_c2r_rows_kernel(float2 *__restrict__ data, const float2 *__restrict__ tw,
                                                          long long nrows, float norm, double *__restrict__ stats)
{
  constexpr int H = N / 2;
  constexpr int ROW_STRIDE = N / 2 + 1;  // complex elements per row (= 2(N/2+1) floats)
  extern __shared__ float2 sm[];
  const long long row0 = (long long)blockIdx.x * W;
  const int tid = threadIdx.x;
  // stage (coalesced along the row, each input element loaded once): with A = X[k], B = conj X[H-k], E = A + B,
  // T = w^k (A - B):  Z[k] = E + iT  and  Z[H-k] = conj(E - iT).  k = 0 pairs X[0] with X[H] (their imaginary
  // parts are not part of a half-complex spectrum and are never read) and also owns the self-paired Z[H/2].
#pragma unroll 4
  for (int idx = tid; idx < W * (H / 2); idx += NT) {
    const int row = idx / (H / 2), k = idx % (H / 2);
    float2 zk = make_float2(0.f, 0.f), zm = make_float2(0.f, 0.f);
    if (row0 + row < nrows) {
      const float2 *x = data + (row0 + row) * ROW_STRIDE;
      float2 a = x[k], bb = x[H - k];
      if (k == 0) {
        const float2 mid = x[H / 2];
        zk = make_float2(a.x + bb.x, a.x - bb.x);
        zm = make_float2(2.f * mid.x, -2.f * mid.y);   // Z[H/2] = 2 conj X[H/2]
      } else {
        bb.y = -bb.y;
        const float2 e = cadd(a, bb), d = csub(a, bb);
        const float2 t = cmul(__ldg(tw + k), d);
        zk = make_float2(e.x - t.y, e.y + t.x);
        zm = make_float2(e.x + t.y, t.x - e.y);
      }
    }
    sm[phys<H, W, RowsLay<W>::value>(k, row)] = zk;
    sm[phys<H, W, RowsLay<W>::value>(k == 0 ? H / 2 : H - k, row)] = zm;
  }
  __syncthreads();
  rows_passes<H, N, W, NT, 0>(sm, tw);
  // gather in natural order (coalesced along the row): z[m] = x[2m] + i x[2m+1]
  double s1 = 0.0, s2 = 0.0;
#pragma unroll 2
  for (int idx = tid; idx < W * H; idx += NT) {
    const int row = idx / H, m = idx % H;
    if (row0 + row < nrows) {
      const float2 z = sm[phys<H, W, RowsLay<W>::value>(freq_to_dif_pos<H>(m), row)];
      const float2 v = make_float2(z.x * norm, z.y * norm);
      data[(row0 + row) * ROW_STRIDE + m] = v;
      if (STATS) {
        s1 += (double)v.x + (double)v.y;
        s2 += (double)__fmul_rn(v.x, v.x) + (double)__fmul_rn(v.y, v.y);
      }
    }
  }
  if (STATS) {
    __shared__ double red[2][NT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_down_sync(0xffffffffu, s1, o);
      s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = s1; red[1][tid >> 5] = s2; }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int i = 0; i < NT / 32; ++i) { a += red[0][i]; b += red[1][i]; }
      stats[GH_PARTIALS_BASE + 2 * (size_t)blockIdx.x] = a;
      stats[GH_PARTIALS_BASE + 1 + 2 * (size_t)blockIdx.x] = b;
    }
  }
}

template <int N, int WSEL = 0, int NTSEL = 0> struct FftCfg {
  static constexpr int W = WSEL ? WSEL : ((N <= 512) ? 16 : (N <= 2048 ? 8 : 4));  // strided tile width (lines)
  static constexpr int NT_S0 = (W * N / 16) < 64 ? 64 : ((W * N / 16) > 512 ? 512 : (W * N / 16));
  static constexpr int NT_S = NTSEL ? NTSEL : ((N <= 1024 && NT_S0 > 256) ? 256 : NT_S0);  // 3 CTAs/SM beat 2 fatter ones (profiles/)
  static constexpr int WT = (N <= 512) ? 16 : (8192 / N);               // TMA-fed strided tile: 64 KB, 3 CTAs per SM
  static constexpr int NT_T = (WT * N / 16) < 64 ? 64 : ((WT * N / 16) > 256 ? 256 : (WT * N / 16));
  static constexpr int WR = (N <= 1024) ? 16 : (N <= 2048 ? 8 : 4);     // rows per tile of the x pass
  static constexpr int NT_R = (WR * (N / 2) / 16) < 64 ? 64 : ((WR * (N / 2) / 16) > 512 ? 512 : (WR * (N / 2) / 16));
};

template <int N, int WSEL, int NTSEL>
int launch_strided(gh_cuda_ctx *c, const float2 *src, float2 *dst, const StridedGeom &g, int ngroups)
{
  using Cfg = FftCfg<N, WSEL, NTSEL>;
  auto kern = fft_strided_kernel<N, Cfg::W, Cfg::NT_S>;
  const size_t smem = (size_t)N * Cfg::W * sizeof(float2);
  GH_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (long long)g.tiles_per_group * ngroups;
  kern<<<(unsigned)tiles, Cfg::NT_S, smem, c->stream>>>(src, dst, c->twiddle, g, c->peers);
  GH_LAUNCH_CHECK(c);
  return 0;
}

// ---- tensor maps of the TMA-fed strided passes ----------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 tma_encoder()
{
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

// z pass: {2 * lines floats, n rows}, row pitch lines * 8 B
static bool make_map_z(CUtensorMap *m, const float2 *base, long long lines, int n, int w, int box_rows)
{
  auto enc = tma_encoder();
  if (!enc || (lines & 1) || 2 * lines > 0xffffffffLL) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)(2 * lines), (cuuint64_t)n};
  const cuuint64_t strides[1] = {(cuuint64_t)lines * sizeof(float2)};
  const cuuint32_t box[2] = {(cuuint32_t)(2 * w), (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// y pass: {4 nh floats (a pair of rows), rows_per_chunk / 2 pairs, planes, chunks}
static bool make_map_y(CUtensorMap *m, const float2 *base, int nh, int rows_per_chunk, int planes, int nchunks, int w, int box_rows)
{
  auto enc = tma_encoder();
  if (!enc || (rows_per_chunk % 16)) return false;  // whole boxes of >= 8 row pairs keep every box 128-byte aligned in shared memory
  const cuuint64_t row = (cuuint64_t)nh * sizeof(float2);
  const cuuint64_t dims[4] = {(cuuint64_t)(4 * nh), (cuuint64_t)(rows_per_chunk / 2), (cuuint64_t)planes, (cuuint64_t)nchunks};
  const cuuint64_t strides[3] = {2 * row, row * rows_per_chunk, row * rows_per_chunk * planes};
  const cuuint32_t box[4] = {(cuuint32_t)(2 * w), (cuuint32_t)box_rows, 1, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int N, bool PAIR>
int launch_strided_tma(gh_cuda_ctx *c, const CUtensorMap &map, const CUtensorMap &map_odd, float2 *dst, StridedGeom g, const TmaGeom &tg,
                       int ngroups)
{
  using Cfg = FftCfg<N>;
  auto kern = fft_strided_tma_kernel<N, Cfg::WT, Cfg::NT_T, PAIR>;
  const size_t smem = ((size_t)N * Cfg::WT + (PAIR ? N : 0)) * sizeof(float2);
  GH_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  g.tiles_per_group = (g.lines_per_group + Cfg::WT - 1) / Cfg::WT;
  const long long tiles = (long long)g.tiles_per_group * ngroups;
  kern<<<(unsigned)tiles, Cfg::NT_T, smem, c->stream>>>(map, map_odd, dst, c->twiddle, g, tg, c->peers);
  GH_LAUNCH_CHECK(c);
  return 0;
}

template <int N, bool STATS>
int launch_rows(gh_cuda_ctx *c, float2 *data, long long nrows, float norm, long long first_block)
{
  using Cfg = FftCfg<N>;
  auto kern = fft_c2r_rows_kernel<N, Cfg::WR, Cfg::NT_R, STATS>;
  const size_t smem = (size_t)(RowsLay<Cfg::WR>::value == LAY_PAD ? Cfg::WR + 1 : Cfg::WR) * (N / 2) * sizeof(float2);
  GH_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (nrows + Cfg::WR - 1) / Cfg::WR;
  kern<<<(unsigned)blocks, Cfg::NT_R, smem, c->stream>>>(data, c->twiddle, nrows, norm, c->d_partials + 2 * first_block);
  GH_LAUNCH_CHECK(c);
  if (STATS) c->fft_stats_blocks = (int)(first_block + blocks);
  return 0;
}

// (1) z axis, local because k-space is ky-distributed: all nky_here*nh columns as one flat group.
// peer_store: the last radix pass stores every output plane straight into the owning rank's receive buffer (the
// transpose fused into the pass); otherwise in place, which leaves block q (the kz planes of rank q's z slab)
// contiguous in the slab, ready to be shipped as one piece.
template <int N, int WSEL = 0, int NTSEL = 0>
int fft_z_pass(gh_cuda_ctx *c, float2 *field, bool peer_store)
{
  const GhDev &d = c->d;
  using Cfg = FftCfg<N, WSEL, NTSEL>;
  const int nh = d.nh;
  StridedGeom g;
  g.lines_per_group = d.nky_here * nh;
  g.tiles_per_group = (g.lines_per_group + Cfg::W - 1) / Cfg::W;
  g.src_group_stride = g.dst_group_stride = 0;
  g.src_stride = g.dst_stride = (long long)d.nky_here * nh;
  g.grp0 = 0;
  g.blk_shift = 30;
  g.src_chunk = 0;
  g.peer_mode = peer_store ? 1 : 0;
  g.nz_peer = d.nz_here;
  g.me = d.rank;
  g.peer_chunk = (long long)d.nz_here * d.nky_here * nh;
  const int mi = (field == c->gridA) ? 0 : 1;
  TmaGeom tg;
  tg.box_rows = N < 256 ? N : 256; tg.nchunks = 1; tg.nh = nh;
  if (c->fft_tma && (field == c->gridA || field == c->gridB) && !c->fft_map_ok[mi])
    c->fft_map_ok[mi] = make_map_z(&c->fft_map[mi], field, g.lines_per_group, N, Cfg::WT, tg.box_rows);
  if (c->fft_tma && (field == c->gridA || field == c->gridB) && c->fft_map_ok[mi])
    return launch_strided_tma<N, false>(c, c->fft_map[mi], c->fft_map[mi], field, g, tg, 1);
  return launch_strided<N, WSEL, NTSEL>(c, field, field, g, 1);
}

// (3) y axis per z plane, (4) x axis half-complex -> real with the normalisation fused.  ysrc: the field itself on
// one rank, a transpose receive buffer [q][z_local][ky_local][kx] (ky = q*nky_here + ky_local) on several.
template <int N, int WSEL = 0, int NTSEL = 0>
int fft_yx_passes(gh_cuda_ctx *c, float2 *field, const float2 *ysrc)
{
  const GhDev &d = c->d;
  using Cfg = FftCfg<N, WSEL, NTSEL>;
  const int nh = d.nh;
  const double normd = pow(sqrt(2.0 * 3.14159265358979323846) / d.l_box, 3.0);  // src/fourier.c:403
  StridedGeom g;
  g.peer_mode = 0; g.nz_peer = 1; g.me = 0; g.peer_chunk = 0;
  g.lines_per_group = nh;
  g.tiles_per_group = (nh + Cfg::W - 1) / Cfg::W;
  g.dst_group_stride = (long long)d.n * nh;
  g.dst_stride = nh;
  if (d.nranks > 1) {
    g.src_group_stride = (long long)d.nky_here * nh;
    g.src_stride = nh;
    g.blk_shift = ilog2(d.nky_here);
    g.src_chunk = (long long)d.nz_here * d.nky_here * nh;
  } else {
    g.src_group_stride = (long long)d.n * nh;
    g.src_stride = nh;
    g.blk_shift = 30;
    g.src_chunk = 0;
  }
  // The two passes can run over batches of planes small enough to stay in the 126 MB L2 (GH_FFT_BATCH_MB; measured: it
  // does not pay on B200, off by default)
  const size_t plane_bytes = (size_t)d.n * nh * sizeof(float2);
  const size_t nb_sz = c->fft_batch_bytes / plane_bytes;
  const int nb = nb_sz < 1 ? 1 : (nb_sz > (size_t)d.nz_here ? d.nz_here : (int)nb_sz);
  for (int z0 = 0; z0 < d.nz_here; z0 += nb) {
    const int nz = (d.nz_here - z0 < nb) ? d.nz_here - z0 : nb;
    g.grp0 = z0;
    // TMA-fed y pass
    const int mi = (field == c->gridA) ? 2 : 3;
    const int nchunks = d.nranks > 1 ? d.nranks : 1, rows_per_chunk = N / nchunks;
    TmaGeom tg;
    tg.box_rows = rows_per_chunk / 2 < 256 ? rows_per_chunk / 2 : 256; tg.nchunks = nchunks; tg.nh = nh;
    const bool tma_src = (field == c->gridA || field == c->gridB);
    if (c->fft_tma && tma_src && !c->fft_map_ok[mi])
      c->fft_map_ok[mi] = make_map_y(&c->fft_map[mi], ysrc, nh, rows_per_chunk, d.nz_here, nchunks, Cfg::WT, tg.box_rows) &&
                          make_map_y(&c->fft_map[mi + 2], ysrc, nh, rows_per_chunk, d.nz_here, nchunks, Cfg::WT + 2, tg.box_rows);
    if (c->fft_tma && tma_src && c->fft_map_ok[mi]) {
      if (launch_strided_tma<N, true>(c, c->fft_map[mi], c->fft_map[mi + 2], field, g, tg, nz)) return 1;
    } else if (launch_strided<N, WSEL, NTSEL>(c, ysrc, field, g, nz)) return 1;
    const long long first_block = ((long long)z0 * d.n) / Cfg::WR;
    if (field == c->gridA) {
      if (launch_rows<N, true>(c, field + (size_t)z0 * d.n * nh, (long long)nz * d.n, (float)normd, first_block)) return 1;
    } else {
      if (launch_rows<N, false>(c, field + (size_t)z0 * d.n * nh, (long long)nz * d.n, (float)normd, 0)) return 1;
    }
  }
  return 0;
}

// One field, start to finish (one rank; several ranks with the transpose fused into the z pass, or through NCCL when
// the peers' buffers cannot be mapped).
template <int N, int WSEL = 0, int NTSEL = 0>
int fft_field(gh_cuda_ctx *c, float2 *field)
{
  const GhDev &d = c->d;
  const int fi = field == c->gridA ? 0 : 1;
  const bool fused = d.nranks > 1 && c->have_peers;
  // peers must be done with their receive buffers (previous field's y pass, previous realisation's maps)
  if (fused && gh_stream_barrier(c)) return 1;
  if (c->time_fft_passes) cudaEventRecord(c->ev_pass[fi][0], c->stream);
  if (fft_z_pass<N, WSEL, NTSEL>(c, field, fused)) return 1;
  const float2 *ysrc = field;
  if (d.nranks > 1) {
    // (2) the one transpose of this field: block q (kz in q's z slab) goes to rank q
    const size_t chunk = (size_t)d.nz_here * d.nky_here * d.nh;
    if (fused) {
      // already done: the z pass stored its output into the peers' receive buffers; wait until all have
      if (gh_stream_barrier(c)) return 1;
    } else {
      GH_NCCL_OK(ncclGroupStart());
      for (int q = 0; q < d.nranks; ++q) {
        GH_NCCL_OK(ncclSend(field + q * chunk, chunk * 2, ncclFloat, q, c->comm, c->stream));
        GH_NCCL_OK(ncclRecv(c->gridC + q * chunk, chunk * 2, ncclFloat, q, c->comm, c->stream));
      }
      GH_NCCL_OK(ncclGroupEnd());
    }
    ysrc = c->gridC;
  }
  if (c->time_fft_passes) cudaEventRecord(c->ev_pass[fi][1], c->stream);
  return fft_yx_passes<N, WSEL, NTSEL>(c, field, ysrc);
}

// The all-to-all as a kernel: a few persistent CTAs push this rank's blocks into the peers' receive buffers with 16-byte
// loads and stores (NVLink writes are posted, so a store stream hides the link latency), on a high-priority stream next
// to the FFT kernels of the other field.  Used instead of the copy engines with GH_TRANSPOSE=push: eight GPUs exchanging
// at once brought the copy engines down to a third of the link rate.
struct PushDst {
  uint4 *p[GH_MAX_RANKS];
};

__global__ void __launch_bounds__(256) transpose_push_kernel(const uint4 *__restrict__ src, PushDst dst, long long chunk16, int me, int P)
{
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int j = 0; j < P; ++j) {
    const int q = (me + 1 + j) % P;  // staggered: at any moment every rank writes to a different peer
    const uint4 *s = src + (long long)q * chunk16;
    uint4 *d = dst.p[q] + (long long)me * chunk16;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < chunk16; i += 4 * stride) {
      const uint4 a = __ldcs(s + i), b = __ldcs(s + i + stride), c = __ldcs(s + i + 2 * stride), e = __ldcs(s + i + 3 * stride);
      __stcs(d + i, a); __stcs(d + i + stride, b); __stcs(d + i + 2 * stride, c); __stcs(d + i + 3 * stride, e);
    }
    for (; i < chunk16; i += stride) __stcs(d + i, __ldcs(s + i));
  }
}

// Both fields on several ranks with the transposes on the copy engines, pipelined against the compute passes:
//   compute stream : barrier | z(A) | z(B)            | wait A, barrier | y(A) x(A) | wait B, barrier | y(B) x(B)
//   copy streams   :         |      | A -> peers' C   | B -> peers' D   (NVLink, 1 large contiguous block per peer)
// The in-place z pass leaves block q of the slab -- the kz planes rank q will own -- contiguous, so the all-to-all is
// nranks - 1 plain peer copies of nz_here * nky_here * nh modes each (plus one local copy), which the copy engines move
// at NVLink line rate without occupying an SM, while the SMs transform the other field.  D is a second receive
// buffer: the (idle) map accumulation stack when it is large enough, else an extra slab when memory allows, else the
// velocity potential's copies wait until every rank has consumed C (one more barrier, less overlap).
// A barrier (1-int all-reduce) after a rank has waited for its own copies means all copies into every rank have landed.
template <int N>
int fft_both_fields_ce(gh_cuda_ctx *c)
{
  const GhDev &d = c->d;
  const int P = d.nranks, me = d.rank;
  const size_t chunk = (size_t)d.nz_here * d.nky_here * d.nh;
  float2 *fields[2] = {c->gridA, c->gridB};
  const bool two = c->recv2 != nullptr;
  auto ship = [&](int fi) -> int {
    if (c->nccl_transpose) {
      // two-sided: a rank's buffer is written only once it has posted the receives, so no barrier is involved
      cudaStream_t s2 = c->ce_stream[0];
      GH_CUDA_OK(cudaStreamWaitEvent(s2, c->ev_z[fi], 0));
      GH_CUDA_OK(cudaStreamWaitEvent(s2, c->ev_free[fi], 0));
      if (c->time_fft_passes) GH_CUDA_OK(cudaEventRecord(c->ev_pass[fi][0], s2));
      float2 *recv = (two && fi == 1) ? c->recv2 : c->gridC;
      GH_NCCL_OK(ncclGroupStart());
      for (int q = 0; q < P; ++q) {
        GH_NCCL_OK(ncclSend(fields[fi] + (size_t)q * chunk, chunk * 2, ncclFloat, q, c->comm2, s2));
        GH_NCCL_OK(ncclRecv(recv + (size_t)q * chunk, chunk * 2, ncclFloat, q, c->comm2, s2));
      }
      GH_NCCL_OK(ncclGroupEnd());
      for (int k = 0; k < GH_N_COPY_STREAMS; ++k) GH_CUDA_OK(cudaEventRecord(c->ev_sent[fi][k], s2));
      if (c->time_fft_passes) GH_CUDA_OK(cudaEventRecord(c->ev_pass[fi][1], s2));
      return 0;
    }
    if (c->push_transpose) {
      cudaStream_t s2 = c->ce_stream[0];
      GH_CUDA_OK(cudaStreamWaitEvent(s2, c->ev_z[fi], 0));
      GH_CUDA_OK(cudaStreamWaitEvent(s2, c->ev_free[fi], 0));
      if (c->time_fft_passes) GH_CUDA_OK(cudaEventRecord(c->ev_pass[fi][0], s2));
      PushDst pd;
      for (int q = 0; q < GH_MAX_RANKS; ++q)
        pd.p[q] = q < P ? reinterpret_cast<uint4 *>((two && fi == 1) ? c->recv2_peers[q] : c->peers.C[q]) : nullptr;
      transpose_push_kernel<<<c->push_ctas, 256, 0, s2>>>(reinterpret_cast<const uint4 *>(fields[fi]), pd, (long long)(chunk / 2), me, P);
      GH_LAUNCH_CHECK(c);
      for (int k = 0; k < GH_N_COPY_STREAMS; ++k) GH_CUDA_OK(cudaEventRecord(c->ev_sent[fi][k], s2));
      if (c->time_fft_passes) GH_CUDA_OK(cudaEventRecord(c->ev_pass[fi][1], s2));
      return 0;
    }
    // copy streams start after the z pass of this field (ev_z[fi]) and after the barrier that freed the destination
    for (int k = 0; k < GH_N_COPY_STREAMS; ++k) {
      GH_CUDA_OK(cudaStreamWaitEvent(c->ce_stream[k], c->ev_z[fi], 0));
      GH_CUDA_OK(cudaStreamWaitEvent(c->ce_stream[k], c->ev_free[fi], 0));
    }
    if (c->time_fft_passes) GH_CUDA_OK(cudaEventRecord(c->ev_pass[fi][0], c->ce_stream[0]));
    for (int j = 0; j < P; ++j) {
      const int q = (me + 1 + j) % P;  // staggered: no two ranks start on the same destination
      float2 *dst = ((two && fi == 1) ? c->recv2_peers[q] : c->peers.C[q]) + (size_t)me * chunk;
      GH_CUDA_OK(cudaMemcpyAsync(dst, fields[fi] + (size_t)q * chunk, chunk * sizeof(float2), cudaMemcpyDefault,
                                 c->ce_stream[j % c->ce_streams_used]));
    }
    for (int k = 0; k < GH_N_COPY_STREAMS; ++k) GH_CUDA_OK(cudaEventRecord(c->ev_sent[fi][k], c->ce_stream[k]));
    if (c->time_fft_passes) {
      // the last copy stream to finish closes the interval: chain them on stream 0
      for (int k = 1; k < GH_N_COPY_STREAMS; ++k) GH_CUDA_OK(cudaStreamWaitEvent(c->ce_stream[0], c->ev_sent[fi][k], 0));
      GH_CUDA_OK(cudaEventRecord(c->ev_pass[fi][1], c->ce_stream[0]));
    }
    return 0;
  };
  auto landed = [&](int fi) -> int {
    for (int k = 0; k < GH_N_COPY_STREAMS; ++k) GH_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_sent[fi][k], 0));
    return c->nccl_transpose ? 0 : gh_stream_barrier(c);
  };
  // every rank is past its previous use of the receive buffers (last realisation's y passes, velocity, maps)
  if (!c->nccl_transpose && gh_stream_barrier(c)) return 1;
  GH_CUDA_OK(cudaEventRecord(c->ev_free[0], c->stream));
  if (two) GH_CUDA_OK(cudaEventRecord(c->ev_free[1], c->stream));
  if (fft_z_pass<N>(c, c->gridA, false)) return 1;
  GH_CUDA_OK(cudaEventRecord(c->ev_z[0], c->stream));
  if (ship(0)) return 1;
  if (fft_z_pass<N>(c, c->gridB, false)) return 1;
  GH_CUDA_OK(cudaEventRecord(c->ev_z[1], c->stream));
  if (two && ship(1)) return 1;
  if (landed(0)) return 1;
  if (two) {
    if (fft_yx_passes<N>(c, c->gridA, c->gridC)) return 1;
    if (landed(1)) return 1;
    return fft_yx_passes<N>(c, c->gridB, c->recv2);
  }
  // one receive buffer: the y pass of A alone first (it is what reads C), then C is free on every rank
  if (fft_yx_passes<N>(c, c->gridA, c->gridC)) return 1;
  if (!c->nccl_transpose && gh_stream_barrier(c)) return 1;
  GH_CUDA_OK(cudaEventRecord(c->ev_free[1], c->stream));
  if (ship(1)) return 1;
  if (landed(1)) return 1;
  return fft_yx_passes<N>(c, c->gridB, c->gridC);
}

// ---- any other even n_grid: general-length passes (gh_fft_generic.cuh) ----------------------------------------
// The reference's FFTW takes any n_grid (src/fourier.c:85-99); the kernels above exist for powers of two.  Everything
// else runs the same three in-place passes (z, y, x with the normalisation fused) through one runtime-length kernel pair:
// a tile of W adjacent lines in shared memory, mixed-radix Stockham passes between two buffers, natural-order output.
// The CTA bodies live in gh_fft_generic.cuh as __host__ __device__ phase functions so that the CPU tests execute the very
// same code (tests/test_fft_generic_cpu.py).  Measured on a B200 (both fields): 384^3 2.1 ms, 640^3 10.8 ms, 768^3 17.9 ms,
// 1536^3 160 ms, i.e. 23-27 Gcells/s against 86 for the tuned 512^3 kernels (profiles/r2/generic_grid_quickcheck_*.log).
// One rank only (the transposes above are written for power-of-two slabs).
__global__ void __launch_bounds__(256) fft_generic_strided_kernel(float2 *data, const float2 *__restrict__ tw, const __grid_constant__ GfftPlan plan,
                                                                  int W, const __grid_constant__ GfftGeom g, int fast)
{
  extern __shared__ float2 sm[];
  for (int phase = 0; phase < plan.nfact + 2; ++phase) {
    gfft_strided_cta_phase(phase, sm, data, tw, plan, W, g, fast, blockIdx.x, threadIdx.x, blockDim.x);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) fft_generic_rows_kernel(float2 *data, const float2 *__restrict__ tw, const __grid_constant__ GfftPlan plan, int W,
                                                               int pitch,
                                                               long long nrows, int nh, float norm, int fast)
{
  extern __shared__ float2 sm[];
  for (int phase = 0; phase < plan.nfact + 2; ++phase) {
    gfft_rows_cta_phase(phase, sm, data, tw, plan, W, pitch, nrows, nh, norm, fast, blockIdx.x, threadIdx.x, blockDim.x);
    __syncthreads();
  }
}

int fft_field_generic(gh_cuda_ctx *c, float2 *field)
{
  const GhDev &d = c->d;
  if (d.nranks != 1) {
    gh_set_error("n_grid=%d is not a power of two: the general-length FFT runs on one rank only", d.n);
    return 1;
  }
  GfftLaunch L;
  if (!gfft_make_launch(d.n, d.nz_here, &L)) {
    gh_set_error("n_grid=%d: the general-length FFT needs an even n_grid whose lines fit in shared memory twice", d.n);
    return 1;
  }
  // one thread per butterfly for the radices 2, 3, 4, 5, 7, one per output element for any other prime; GH_FFT_GENERIC_SLOW=1:
  // one per output element throughout, the first version (both validated on a B200 against the oracle; 3.2x apart:
  // profiles/r2/generic_grid_*.log)
  const int fast = getenv("GH_FFT_GENERIC_SLOW") == nullptr;
  GH_CUDA_OK(cudaFuncSetAttribute(fft_generic_strided_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_s));
  GH_CUDA_OK(cudaFuncSetAttribute(fft_generic_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_r));
  fft_generic_strided_kernel<<<(unsigned)L.blocks_z, 256, L.smem_s, c->stream>>>(field, c->twiddle, L.pn, L.W, L.gz, fast);
  GH_LAUNCH_CHECK(c);
  fft_generic_strided_kernel<<<(unsigned)L.blocks_y, 256, L.smem_s, c->stream>>>(field, c->twiddle, L.pn, L.W, L.gy, fast);
  GH_LAUNCH_CHECK(c);
  const double normd = pow(sqrt(2.0 * 3.14159265358979323846) / d.l_box, 3.0);  // src/fourier.c:403
  fft_generic_rows_kernel<<<(unsigned)L.blocks_x, 256, L.smem_r, c->stream>>>(field, c->twiddle, L.ph, L.WR, L.pitch, L.nrows, d.nh, (float)normd, fast);
  GH_LAUNCH_CHECK(c);
  return 0;
}

}  // namespace

int gh_fft_supported(int n)
{
  // powers of two 32..4096 have tuned kernels; every other even n_grid takes the general-length passes
  return n >= 8 && n <= 4096 && (n & 1) == 0;
}

// lengths with tuned kernels (and the only ones the multi-rank transposes are written for)
int gh_fft_tuned(int n)
{
  return n == 32 || n == 64 || n == 128 || n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096;
}

int gh_launch_fft_both_fields(gh_cuda_ctx *c)
{
  if (c->d.nranks > 1 && c->have_peers && c->ce_transpose) {
    switch (c->d.n) {
      case 32: return fft_both_fields_ce<32>(c);
      case 64: return fft_both_fields_ce<64>(c);
      case 128: return fft_both_fields_ce<128>(c);
      case 256: return fft_both_fields_ce<256>(c);
      case 512: return fft_both_fields_ce<512>(c);
      case 1024: return fft_both_fields_ce<1024>(c);
      case 2048: return fft_both_fields_ce<2048>(c);
      case 4096: return fft_both_fields_ce<4096>(c);
      default: break;
    }
  }
  if (gh_launch_fft_field(c, c->gridA)) return 1;  // src/fourier.c:391
  return gh_launch_fft_field(c, c->gridB);         // src/fourier.c:392
}

int gh_launch_fft_field(gh_cuda_ctx *c, float2 *field)
{
  switch (c->d.n) {
    case 32: return fft_field<32>(c, field);
    case 64: return fft_field<64>(c, field);
    case 128: return fft_field<128>(c, field);
    case 256: return fft_field<256>(c, field);
    case 512: return fft_field<512>(c, field);
    case 1024: return fft_field<1024>(c, field);
    case 2048: return fft_field<2048>(c, field);
    case 4096: return fft_field<4096>(c, field);
    default: return fft_field_generic(c, field);
  }
}
