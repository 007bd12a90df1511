// Micro-benchmark: how fast can one SM pull a [T][C][W pixels] slab out of x[B,T,C,hw] (bf16) as a function of the
// row width W (bytes contiguous per (t, c) row), by TMA tensor boxes, by 1-D bulk copies and by LDG.128?
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory"); } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
constexpr int T = 61, C = 64;
// mode 0: TMA box {W, C, FPB frames}; mode 1: 1-D bulk copy per row (all 128 threads issue); mode 2: LDG.128 -> STS.128
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap map, const uint16_t* x, int hw, int W, int FPB, int tiles_per_b, int n_tiles, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar[2];
  const int slab = T * C * W * 2;
  if (threadIdx.x == 0) { mbar_init(s32(&bar[0]), 1); mbar_init(s32(&bar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  float acc = 0.f;
  int it = 0;
  auto issue = [&](int tile, int buf) {
    const int b = tile / tiles_per_b, pix0 = (tile - b * tiles_per_b) * W;
    const uint32_t dst = s32(smem) + buf * slab, br = s32(&bar[buf]);
    if (MODE == 0) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(br, slab);
        for (int t = 0; t < T; t += FPB) {
          // the last box may run past T within the tensor (b*T+t+FPB <= B*T as long as b < B-1): keep T % FPB == 0 or FPB=1
          tma_load_3d(dst + t * C * W * 2, &map, pix0, 0, b * T + t, br);
        }
      }
    } else if (MODE == 1) {
      if (threadIdx.x == 0) mbar_expect_tx(br, slab);
      __syncthreads();
      for (int r = threadIdx.x; r < T * C; r += blockDim.x)
        bulk_g2s(dst + r * W * 2, x + (static_cast<size_t>(b) * T * C + r) * hw + pix0, W * 2, br);
    }
  };
  if (MODE <= 1) {
    issue(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      if (tile + gridDim.x < n_tiles) issue(tile + gridDim.x, buf ^ 1);
      mbar_wait(s32(&bar[buf]), (it >> 1) & 1);
      acc += reinterpret_cast<const float*>(smem + buf * slab)[threadIdx.x];
      __syncthreads();
    }
  } else {
    const int vpr = W / 8;  // 16-byte vectors per row
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_b, pix0 = (tile - b * tiles_per_b) * W;
      const uint16_t* base = x + static_cast<size_t>(b) * T * C * hw + pix0;
      for (int i0 = threadIdx.x; i0 < T * C * vpr; i0 += blockDim.x * 8) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * blockDim.x;
          v[u] = make_uint4(0, 0, 0, 0);
          if (i < T * C * vpr) {
            const int r = i / vpr, q = i - r * vpr;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(base + static_cast<size_t>(r) * hw + q * 8));
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * blockDim.x;
          if (i < T * C * vpr) reinterpret_cast<uint4*>(smem)[i] = v[u];
        }
      }
      __syncthreads();
      acc += reinterpret_cast<const float*>(smem)[threadIdx.x];
      __syncthreads();
    }
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  const int B = 8, hw = 16384;
  const size_t n = static_cast<size_t>(B) * T * C * hw;
  uint16_t* x; cudaMalloc(&x, n * 2); cudaMemset(x, 0, n * 2);
  float* sink; cudaMalloc(&sink, 148 * 256 * 4);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int W : {8, 16, 32, 64}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int FPB : {1, 61}) {
        if (mode != 0 && FPB != 1) continue;
        if (mode == 0 && FPB == 61 && W * C * 61 * 2 > 200000) continue;
        CUtensorMap map;
        const cuuint64_t dims[3] = {(cuuint64_t)hw, (cuuint64_t)C, (cuuint64_t)B * T};
        const cuuint64_t strides[2] = {(cuuint64_t)hw * 2, (cuuint64_t)C * hw * 2};
        const cuuint32_t box[3] = {(cuuint32_t)W, (cuuint32_t)C, (cuuint32_t)FPB};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int tiles_per_b = hw / W, n_tiles = B * tiles_per_b;
        const int slab = T * C * W * 2;
        const int smem = (mode == 2 ? 1 : 2) * slab;
        if (smem > 220000) { printf("W=%d mode=%d: smem %d too large\n", W, mode, smem); continue; }
        auto launch = [&]() {
          if (mode == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<0><<<148, 256, smem>>>(map, x, hw, W, FPB, tiles_per_b, n_tiles, sink); }
          if (mode == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<1><<<148, 256, smem>>>(map, x, hw, W, FPB, tiles_per_b, n_tiles, sink); }
          if (mode == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<2><<<148, 256, smem>>>(map, x, hw, W, FPB, tiles_per_b, n_tiles, sink); }
        };
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("W=%2d px (%3d B rows) mode=%s FPB=%2d: %.3f ms  %.0f GB/s  (%s)\n", W, W * 2, mode == 0 ? "tma " : mode == 1 ? "bulk" : "ldg ", FPB, ms, n * 2 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
