// Stand-alone probe for the next round's FFT work (not part of libgh_cuda.so): how fast can a strided FFT pass
// stream its [N positions][W adjacent lines] tiles of complex-float through shared memory on B200,
//   (a) with plain coalesced loads (what fft_strided_kernel does today: W*8 B contiguous per position), and
//   (b) with TMA 2-D tile loads (cp.async.bulk.tensor) into a double-buffered shared tile, mbarrier-signalled,
// both writing the tile back with coalesced stores (the traffic of one in-place FFT pass: 8 B read + 8 B written per
// mode).  The y pass of a z-distributed slab is modelled: for every plane, lines run along y (stride nh modes),
// adjacent lines are adjacent kx.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_tile_probe tma_tile_probe.cu
//   ./tma_tile_probe [n=1024] [planes=64]
//
// Written on a box without a GPU at the end of round 1; it compiles, it has not been run.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

// ---- plain version: one CTA per tile, tile staged through shared memory -----------------------------------
template <int W>
__global__ void __launch_bounds__(256) plain_tile_copy(const float2 *__restrict__ src, float2 *__restrict__ dst, int n, int nh,
                                                       int tiles_per_plane, long long n_tiles)
{
  extern __shared__ float2 tile[];
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long plane = t / tiles_per_plane;
    const int kx0 = (int)(t % tiles_per_plane) * W;
    const float2 *s = src + plane * (long long)n * nh + kx0;
    float2 *d = dst + plane * (long long)n * nh + kx0;
    for (int i = threadIdx.x; i < n * W; i += blockDim.x) {
      const int pos = i / W, w = i % W;
      if (kx0 + w < nh) tile[i] = s[(long long)pos * nh + w];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * W; i += blockDim.x) {
      const int pos = i / W, w = i % W;
      if (kx0 + w < nh) d[(long long)pos * nh + w] = tile[i];
    }
    __syncthreads();
  }
}

// ---- TMA version ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned phase)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
      "r"(phase)
      : "memory");
}
// 3-D tensor (x = floats along kx, y = position along the line, z = plane); box = {2W floats, rows, 1}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(map), "r"(x), "r"(y), "r"(z), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

template <int W, int ROWS_PER_BOX>
__global__ void __launch_bounds__(256) tma_tile_copy(const __grid_constant__ CUtensorMap map, float2 *__restrict__ dst, int n, int nh,
                                                     int tiles_per_plane, long long n_tiles)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2 *buf[2] = {reinterpret_cast<float2 *>(smem_raw), reinterpret_cast<float2 *>(smem_raw) + (size_t)n * W};
  __shared__ uint64_t full[2];
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned tile_bytes = (unsigned)(n * W * sizeof(float2));
  auto issue = [&](long long t, int stage) {
    const int plane = (int)(t / tiles_per_plane), kx0 = (int)(t % tiles_per_plane) * W;
    mbar_expect_tx(&full[stage], tile_bytes);
    for (int r = 0; r < n; r += ROWS_PER_BOX) tma_load_3d(buf[stage] + (size_t)r * W, &map, &full[stage], 2 * kx0, r, plane);
  };
  long long t = blockIdx.x;
  if (threadIdx.x == 0 && t < n_tiles) issue(t, 0);
  unsigned phase[2] = {0, 0};
  int stage = 0;
  for (; t < n_tiles; t += gridDim.x, stage ^= 1) {
    const long long tn = t + gridDim.x;
    if (threadIdx.x == 0 && tn < n_tiles) issue(tn, stage ^ 1);  // prefetch the next tile while this one is consumed
    mbar_wait(&full[stage], phase[stage]);
    phase[stage] ^= 1;
    const long long plane = t / tiles_per_plane;
    const int kx0 = (int)(t % tiles_per_plane) * W;
    float2 *d = dst + plane * (long long)n * nh + kx0;
    const float2 *tile = buf[stage];
    for (int i = threadIdx.x; i < n * W; i += blockDim.x) {
      const int pos = i / W, w = i % W;
      if (kx0 + w < nh) d[(long long)pos * nh + w] = tile[i];
    }
    __syncthreads();  // everybody is done with this stage before it is refilled two iterations later
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode()
{
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) {
    fprintf(stderr, "cuTensorMapEncodeTiled not available\n");
    exit(1);
  }
  return (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

template <int W>
static void run(int n, int planes, int n_sm)
{
  const int nh = n / 2 + 1;
  // the row pitch must be a multiple of 16 B for TMA: nh*8 is (n even -> nh odd -> 8 mod 16): pad the pitch by one mode
  const int pitch = nh + (nh & 1);
  const size_t modes = (size_t)planes * n * pitch;
  float2 *a, *b;
  CK(cudaMalloc(&a, modes * sizeof(float2)));
  CK(cudaMalloc(&b, modes * sizeof(float2)));
  CK(cudaMemset(a, 1, modes * sizeof(float2)));
  const int tiles_per_plane = (pitch + W - 1) / W;
  const long long n_tiles = (long long)tiles_per_plane * planes;
  const size_t smem1 = (size_t)n * W * sizeof(float2), smem2 = 2 * smem1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const double gbytes = 2.0 * modes * sizeof(float2) / 1e9;
  float ms;
  if (smem1 <= 227 * 1024) {
    CK(cudaFuncSetAttribute(plain_tile_copy<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, plain_tile_copy<W>, 256, smem1));
    const int grid = n_sm * (per_sm > 0 ? per_sm : 1);
    for (int it = 0; it < 3; ++it) {
      if (it == 1) CK(cudaEventRecord(e0));
      plain_tile_copy<W><<<grid, 256, smem1>>>(a, b, n, pitch, tiles_per_plane, n_tiles);
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("n=%d W=%d plain: %d CTA/SM, %.3f ms per pass, %.0f GB/s\n", n, W, per_sm, ms / 2, gbytes / (ms / 2 * 1e-3));
  }
  if (smem2 <= 227 * 1024) {
    constexpr int ROWS = 256;  // TMA box dimensions are limited to 256
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)2 * pitch, (cuuint64_t)n, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(float2), (cuuint64_t)pitch * n * sizeof(float2)};
    const cuuint32_t box[3] = {2 * W, (cuuint32_t)(n < ROWS ? n : ROWS), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
      exit(1);
    }
    CK(cudaFuncSetAttribute(tma_tile_copy<W, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tma_tile_copy<W, ROWS>, 256, smem2));
    const int grid = n_sm * (per_sm > 0 ? per_sm : 1);
    for (int it = 0; it < 3; ++it) {
      if (it == 1) CK(cudaEventRecord(e0));
      tma_tile_copy<W, ROWS><<<grid, 256, smem2>>>(map, b, n, pitch, tiles_per_plane, n_tiles);
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("n=%d W=%d TMA (2 stages): %d CTA/SM, %.3f ms per pass, %.0f GB/s\n", n, W, per_sm, ms / 2, gbytes / (ms / 2 * 1e-3));
  }
  CK(cudaFree(a));
  CK(cudaFree(b));
}

int main(int argc, char **argv)
{
  const int n = argc > 1 ? atoi(argv[1]) : 1024;
  const int planes = argc > 2 ? atoi(argv[2]) : 64;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s, %d SMs; one pass = %.2f GB\n", prop.name, prop.multiProcessorCount, 2.0 * planes * n * (n / 2 + 2) * 8 / 1e9);
  run<16>(n, planes, prop.multiProcessorCount);
  run<8>(n, planes, prop.multiProcessorCount);
  run<4>(n, planes, prop.multiProcessorCount);
  return 0;
}
