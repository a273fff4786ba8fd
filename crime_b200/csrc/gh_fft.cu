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
//                        registers, last pass straight to global (digit-reversed scatter).
//   fft_c2r_rows_kernel: x axis, rows are contiguous.  Half-size complex transform with the Hermitian
//                        pre-twist fused into the load (Im of DC and Nyquist are never read, as FFTW's
//                        c2r), decimation-in-time so the last pass stores coalesced, in place, scaled by
//                        (sqrt(2 pi)/L)^3.
#include "gh_internal.cuh"

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
template <int N> __device__ __forceinline__ int dif_pos_to_freq(int pos)
{
  int f = 0;
#pragma unroll
  for (int s = 0; s < n_steps(N); ++s) {
    const int sub = N / rad_prod(N, s + 1);
    const int q = (pos / sub) % rad_at(N, s);
    f += q * rad_prod(N, s);
  }
  return f;
}
template <int N> __device__ __forceinline__ int freq_to_dif_pos(int f)
{
  int pos = 0;
#pragma unroll
  for (int s = 0; s < n_steps(N); ++s) {
    const int q = (f / rad_prod(N, s)) % rad_at(N, s);
    pos += q * (N / rad_prod(N, s + 1));
  }
  return pos;
}

// ---- strided axis -------------------------------------------------------------------------------
struct StridedGeom {
  int lines_per_group;      // lines sharing one base offset rule (a plane's kx columns, or all columns flat)
  int tiles_per_group;
  long long src_group_stride, dst_group_stride;
  long long src_stride, dst_stride;  // element stride along the transformed axis
  int blk_shift;            // source index n is split as (n >> blk_shift, n & mask): the received
  long long src_chunk;      // all-to-all blocks sit src_chunk apart (blk_shift = 30 disables the split)
};

template <int N, int W, int NT, int S>
struct DifSteps {
  static __device__ __forceinline__ void run(float2 *sm, const float2 *__restrict__ src, float2 *__restrict__ dst,
                                             const float2 *__restrict__ tw, const StridedGeom &g, long long sbase,
                                             long long dbase, int nvalid)
  {
    constexpr int R = rad_at(N, S);
    constexpr int M = N / rad_prod(N, S);  // current block length
    constexpr int SUB = M / R;
    constexpr bool FIRST = (S == 0), LAST = (S == n_steps(N) - 1);
    const int tid = threadIdx.x;
    const int blk_mask = (1 << g.blk_shift) - 1;
#pragma unroll 1
    for (int item = tid; item < W * (N / R); item += NT) {
      const int w = item % W, ii = item / W;
      const int b = ii / SUB, i = ii % SUB;
      const int p0 = b * M + i;
      float2 u[R];
      if constexpr (FIRST) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int n = p0 + r * SUB;
          const long long a = sbase + (long long)(n >> g.blk_shift) * g.src_chunk + (long long)(n & blk_mask) * g.src_stride + w;
          u[r] = (w < nvalid) ? __ldg(src + a) : make_float2(0.f, 0.f);
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = sm[(p0 + r * SUB) * W + w];
      }
      dft<R>(u);
      if constexpr (!LAST) {
        if (SUB > 1) twiddle_powers<R>(u, __ldg(tw + i * (N / M)));
#pragma unroll
        for (int q = 0; q < R; ++q) sm[(p0 + q * SUB) * W + w] = u[q];
      } else {
        // SUB == 1: p0 = b*R; scatter to natural order
        if (w < nvalid) {
          const int f0 = dif_pos_to_freq<N>(p0);
#pragma unroll
          for (int q = 0; q < R; ++q) dst[dbase + (long long)(f0 + q * (N / R)) * g.dst_stride + w] = u[q];
        }
      }
    }
    if constexpr (!LAST) {
      __syncthreads();
      DifSteps<N, W, NT, S + 1>::run(sm, src, dst, tw, g, sbase, dbase, nvalid);
    }
  }
};

template <int N, int W, int NT>
__global__ void __launch_bounds__(NT) fft_strided_kernel(const float2 *__restrict__ src, float2 *__restrict__ dst,
                                                         const float2 *__restrict__ tw, StridedGeom g)
{
  extern __shared__ float2 sm[];
  const int tile = blockIdx.x;
  const int grp = tile / g.tiles_per_group;
  const int l0 = (tile - grp * g.tiles_per_group) * W;
  const long long sbase = (long long)grp * g.src_group_stride + l0;
  const long long dbase = (long long)grp * g.dst_group_stride + l0;
  const int nvalid = min(W, g.lines_per_group - l0);
  DifSteps<N, W, NT, 0>::run(sm, src, dst, tw, g, sbase, dbase, nvalid);
}

// ---- contiguous rows: half-complex -> real, in place ----------------------------------------------
// decimation in time over the half length H = N/2 with the DIF radix list reversed
template <int N, int ROWS, int NT, int T>
struct DitSteps {
  static constexpr int H = N / 2;
  static __device__ __forceinline__ void run(float2 *sm, float2 *__restrict__ out, const float2 *__restrict__ tw,
                                             long long row0, long long nrows, long long row_stride, float norm)
  {
    constexpr int NS = n_steps(H);
    constexpr int S = NS - 1 - T;              // index into the DIF list
    constexpr int R = rad_at(H, S);
    constexpr int SUB = H / rad_prod(H, S + 1);  // product of the radices already applied
    constexpr int M = SUB * R;
    constexpr bool LAST = (T == NS - 1);
    const int tid = threadIdx.x;
#pragma unroll 1
    for (int item = tid; item < ROWS * (H / R); item += NT) {
      const int row = item / (H / R), ii = item % (H / R);
      const int b = ii / SUB, i = ii % SUB;
      const int p0 = b * M + i;
      float2 *s = sm + row * H;
      float2 u[R];
#pragma unroll
      for (int r = 0; r < R; ++r) u[r] = s[p0 + r * SUB];
      if (SUB > 1) twiddle_powers<R>(u, __ldg(tw + i * (N / M)));
      dft<R>(u);
      if constexpr (!LAST) {
#pragma unroll
        for (int q = 0; q < R; ++q) s[p0 + q * SUB] = u[q];
      } else {
        if (row0 + row < nrows) {
          float2 *o = out + (row0 + row) * row_stride;
#pragma unroll
          for (int q = 0; q < R; ++q) o[p0 + q * SUB] = make_float2(u[q].x * norm, u[q].y * norm);
        }
      }
    }
    if constexpr (!LAST) {
      __syncthreads();
      DitSteps<N, ROWS, NT, T + 1>::run(sm, out, tw, row0, nrows, row_stride, norm);
    }
  }
};

template <int N, int ROWS, int NT>
__global__ void __launch_bounds__(NT) fft_c2r_rows_kernel(float2 *__restrict__ data, const float2 *__restrict__ tw,
                                                          long long nrows, float norm)
{
  constexpr int H = N / 2;
  constexpr long long ROW_STRIDE = N / 2 + 1;  // complex elements per row (= 2(N/2+1) floats)
  extern __shared__ float2 sm[];
  const long long row0 = (long long)blockIdx.x * ROWS;
  const int tid = threadIdx.x;
  // stage: Z[k] = (X[k] + conj X[H-k]) + i w^k (X[k] - conj X[H-k]), stored digit-reversed
  for (int idx = tid; idx < ROWS * H; idx += NT) {
    const int row = idx / H, k = idx % H;
    float2 z = make_float2(0.f, 0.f);
    if (row0 + row < nrows) {
      const float2 *x = data + (row0 + row) * ROW_STRIDE;
      float2 a = x[k], bb = x[H - k];
      bb.y = -bb.y;
      if (k == 0) { a.y = 0.f; bb.y = 0.f; }  // Im(DC), Im(Nyquist) are not part of a half-complex spectrum
      const float2 e = cadd(a, bb), d = csub(a, bb);
      const float2 t = cmul(__ldg(tw + k), d);
      z = make_float2(e.x - t.y, e.y + t.x);
    }
    sm[row * H + freq_to_dif_pos<H>(k)] = z;
  }
  __syncthreads();
  DitSteps<N, ROWS, NT, 0>::run(sm, data, tw, row0, nrows, ROW_STRIDE, norm);
}

template <int N> struct FftCfg {
  static constexpr int W = (N <= 512) ? 16 : (N <= 2048 ? 8 : 4);
  static constexpr int NT_S = (W * N / 16) < 64 ? 64 : ((W * N / 16) > 512 ? 512 : (W * N / 16));
  static constexpr int ROWS = (N / 2 >= 2048) ? 2 : ((4096 / (N / 2)) > 16 ? 16 : (4096 / (N / 2)));
  static constexpr int NT_R = (ROWS * (N / 2) / 8) < 64 ? 64 : ((ROWS * (N / 2) / 8) > 512 ? 512 : (ROWS * (N / 2) / 8));
};

template <int N>
int launch_strided(gh_cuda_ctx *c, const float2 *src, float2 *dst, const StridedGeom &g, int ngroups)
{
  using Cfg = FftCfg<N>;
  auto kern = fft_strided_kernel<N, Cfg::W, Cfg::NT_S>;
  const size_t smem = (size_t)N * Cfg::W * sizeof(float2);
  GH_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (long long)g.tiles_per_group * ngroups;
  kern<<<(unsigned)tiles, Cfg::NT_S, smem, c->stream>>>(src, dst, c->twiddle, g);
  GH_LAUNCH_CHECK(c);
  return 0;
}

template <int N>
int launch_rows(gh_cuda_ctx *c, float2 *data, long long nrows, float norm)
{
  using Cfg = FftCfg<N>;
  auto kern = fft_c2r_rows_kernel<N, Cfg::ROWS, Cfg::NT_R>;
  const size_t smem = (size_t)Cfg::ROWS * (N / 2) * sizeof(float2);
  GH_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (nrows + Cfg::ROWS - 1) / Cfg::ROWS;
  kern<<<(unsigned)blocks, Cfg::NT_R, smem, c->stream>>>(data, c->twiddle, nrows, norm);
  GH_LAUNCH_CHECK(c);
  return 0;
}

template <int N>
int fft_field(gh_cuda_ctx *c, float2 *field)
{
  const GhDev &d = c->d;
  using Cfg = FftCfg<N>;
  const int nh = d.nh;
  const double normd = pow(sqrt(2.0 * 3.14159265358979323846) / d.l_box, 3.0);  // src/fourier.c:403
  // (1) z axis, local because k-space is ky-distributed: all nky_here*nh columns as one flat group
  {
    StridedGeom g;
    g.lines_per_group = d.nky_here * nh;
    g.tiles_per_group = (g.lines_per_group + Cfg::W - 1) / Cfg::W;
    g.src_group_stride = g.dst_group_stride = 0;
    g.src_stride = g.dst_stride = (long long)d.nky_here * nh;
    g.blk_shift = 30;
    g.src_chunk = 0;
    if (launch_strided<N>(c, field, field, g, 1)) return 1;
  }
  const float2 *ysrc = field;
  StridedGeom g;
  g.lines_per_group = nh;
  g.tiles_per_group = (nh + Cfg::W - 1) / Cfg::W;
  g.dst_group_stride = (long long)d.n * nh;
  g.dst_stride = nh;
  if (d.nranks > 1) {
    // (2) the one transpose of this field: block q (kz in q's z slab) goes to rank q
    const size_t chunk = (size_t)d.nz_here * d.nky_here * nh;
    GH_NCCL_OK(ncclGroupStart());
    for (int q = 0; q < d.nranks; ++q) {
      GH_NCCL_OK(ncclSend(field + q * chunk, chunk * 2, ncclFloat, q, c->comm, c->stream));
      GH_NCCL_OK(ncclRecv(c->gridC + q * chunk, chunk * 2, ncclFloat, q, c->comm, c->stream));
    }
    GH_NCCL_OK(ncclGroupEnd());
    // received layout [q][z_local][ky_local][kx]; ky = q*nky_here + ky_local
    ysrc = c->gridC;
    g.src_group_stride = (long long)d.nky_here * nh;
    g.src_stride = nh;
    g.blk_shift = ilog2(d.nky_here);
    g.src_chunk = (long long)chunk;
  } else {
    g.src_group_stride = (long long)d.n * nh;
    g.src_stride = nh;
    g.blk_shift = 30;
    g.src_chunk = 0;
  }
  // (3) y axis per z plane, (4) x axis half-complex -> real with the normalisation fused
  if (launch_strided<N>(c, ysrc, field, g, d.nz_here)) return 1;
  return launch_rows<N>(c, field, (long long)d.nz_here * d.n, (float)normd);
}

}  // namespace

int gh_fft_supported(int n)
{
  return n == 32 || n == 64 || n == 128 || n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096;
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
    default:
      gh_set_error("n_grid=%d is not supported by the sm_100a FFT (powers of two 32..4096)", c->d.n);
      return 1;
  }
}
