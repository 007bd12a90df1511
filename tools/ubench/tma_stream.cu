// Micro-benchmark: streaming x[B,T,C,hw] (bf16) through a shared-memory ring with TMA tensor boxes
// {W pixels, C channels, F frames}, as a function of the row width W, the CTAs per SM, the number of passes over every
// tile (2 = statistics pass + compute pass, the second one served by L2) and whether the consumers read the data back
// from shared memory (LDS.128 of every byte).  Answers: which tile width can reach the HBM rate, what a second pass
// costs.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/tma_stream tools/ubench/tma_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory"); } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
constexpr int T = 64, C = 64;  // T = 64 so that every chunk is full
constexpr int NTHREADS = 288;  // 8 consumer warps + 1 producer warp

__global__ void __launch_bounds__(NTHREADS) k(const __grid_constant__ CUtensorMap map, int W, int F, int stages, int passes,
                                               int read_back, int tiles_per_b, int n_tiles, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[32];
  const int stage_bytes = W * 2 * C * F;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(s32(&bars[s]), 1); mbar_init(s32(&bars[16 + s]), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int chunks = T / F;
  if (warp == 8) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_b, pix0 = (tile - b * tiles_per_b) * W;
        for (int ps = 0; ps < passes; ++ps)
          for (int c = 0; c < chunks; ++c, ++it) {
            const int s = it % stages;
            if (it >= stages) mbar_wait(s32(&bars[16 + s]), ((it / stages) - 1) & 1);
            mbar_expect_tx(s32(&bars[s]), stage_bytes);
            tma_load_3d(s32(smem) + s * stage_bytes, &map, pix0, 0, b * T + c * F, s32(&bars[s]));
          }
      }
    }
  } else {
    float acc = 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int ps = 0; ps < passes; ++ps)
        for (int c = 0; c < chunks; ++c, ++it) {
          const int s = it % stages;
          mbar_wait(s32(&bars[s]), (it / stages) & 1);
          const uint4* p = reinterpret_cast<const uint4*>(smem + s * stage_bytes);
          if (read_back) {
            for (int i = threadIdx.x; i < stage_bytes / 16; i += 256) { const uint4 v = p[i]; acc += __uint_as_float(v.x ^ v.y ^ v.z ^ v.w); }
          } else {
            acc += __uint_as_float(p[threadIdx.x].x);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(s32(&bars[16 + s]));
        }
    sink[blockIdx.x * 256 + threadIdx.x] = acc;
  }
}

int main() {
  const int B = 16, hw = 16384;
  const size_t n = static_cast<size_t>(B) * T * C * hw;  // 2.1 GB bf16: far beyond L2
  uint16_t* x; cudaMalloc(&x, n * 2); cudaMemset(x, 0, n * 2);
  float* sink; cudaMalloc(&sink, 148 * 4 * 256 * 4);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int W, F, stages, cta_per_sm, passes, read_back, swizzle; };
  const Cfg cfgs[] = {
      {8, 16, 4, 1, 1, 0, 0},  {8, 16, 4, 2, 1, 0, 0},  {8, 16, 4, 2, 1, 1, 0},
      {16, 8, 4, 1, 1, 0, 0},  {16, 8, 4, 2, 1, 0, 0},  {16, 8, 4, 2, 1, 1, 0}, {16, 8, 4, 2, 1, 1, 1}, {16, 8, 4, 2, 2, 1, 0},
      {32, 4, 4, 1, 1, 0, 0},  {32, 4, 4, 2, 1, 0, 0},  {32, 4, 4, 2, 1, 1, 0}, {32, 4, 4, 2, 1, 1, 1}, {32, 4, 4, 2, 2, 1, 1},
      {32, 8, 3, 2, 1, 1, 1},  {32, 8, 3, 2, 2, 1, 1},  {32, 16, 2, 1, 2, 1, 1},
      {64, 2, 4, 2, 1, 1, 1},  {64, 4, 3, 2, 1, 1, 1},  {64, 4, 3, 2, 2, 1, 1}, {64, 4, 4, 1, 2, 1, 1},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)hw, (cuuint64_t)C, (cuuint64_t)B * T};
    const cuuint64_t strides[2] = {(cuuint64_t)hw * 2, (cuuint64_t)C * hw * 2};
    const cuuint32_t box[3] = {(cuuint32_t)c.W, (cuuint32_t)C, (cuuint32_t)c.F};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    if (c.swizzle) sw = c.W == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : c.W == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d (W=%d F=%d)\n", (int)r, c.W, c.F); continue; }
    const int tiles_per_b = hw / c.W, n_tiles = B * tiles_per_b;
    const int smem = c.stages * c.W * 2 * C * c.F + 1024;
    const int grid = 148 * c.cta_per_sm;
    auto launch = [&]() { k<<<grid, NTHREADS, smem>>>(map, c.W, c.F, c.stages, c.passes, c.read_back, tiles_per_b, n_tiles, sink); };
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("W=%2d px (%3d B rows) F=%2d stages=%d (%3d KB) cta/sm=%d passes=%d read_back=%d swizzle=%d: %.3f ms  %.0f GB/s unique  (%s)\n",
           c.W, c.W * 2, c.F, c.stages, smem / 1024, c.cta_per_sm, c.passes, c.read_back, c.swizzle, ms, n * 2 / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
