// Micro-benchmark: per-SM throughput of mma.sync m16n8k16 bf16, ldmatrix.x4, stmatrix.x4, movmatrix on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(sm)) + warp * 2048;
  uint32_t addr = base + (lane >> 3) * 512 + ((lane & 7) ^ 3) * 16;
  float acc[8][4] = {};
  uint32_t a[4] = {0x3f803f80u + lane, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, r[4] = {1u * lane, 2, 3, 4};
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
#pragma unroll
      for (int n = 0; n < 8; ++n) mma_bf16(acc[n], a, r[0], r[1]);
    } else if (MODE == 1) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        uint32_t q[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(addr + (n & 3) * 128));
        r[0] ^= q[0]; r[1] ^= q[1]; r[2] ^= q[2]; r[3] ^= q[3];
      }
    } else if (MODE == 2) {
#pragma unroll
      for (int n = 0; n < 8; ++n)
        asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%4], {%0,%1,%2,%3};" :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3] + n), "r"(addr + (n & 3) * 128) : "memory");
    } else if (MODE == 3) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        uint32_t q;
        asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(q) : "r"(r[n & 3]));
        r[n & 3] = q + 1;
      }
    } else if (MODE == 4) {  // mma + ldmatrix interleaved 4:1
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        uint32_t q[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(addr + n * 128));
        mma_bf16(acc[4 * n], a, q[0], q[1]);
        mma_bf16(acc[4 * n + 1], a, q[2], q[3]);
        mma_bf16(acc[4 * n + 2], a, q[0], q[1]);
        mma_bf16(acc[4 * n + 3], a, q[2], q[3]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int n = 0; n < 8; ++n) for (int e = 0; e < 4; ++e) s += acc[n][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + r[0] + r[1] + r[2] + r[3];
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int threads, int iters) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<MODE><<<148, threads, 65536>>>(out, 10, cyc);
  k<MODE><<<148, threads, 65536>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double ops = double(iters) * 8 * (threads / 32);
  printf("%-28s threads=%4d  cycles/op/SM = %.3f  (err %s)\n", name, threads, c / ops, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  for (int th : {128, 256, 512, 1024}) {
    run<0>("mma.m16n8k16.bf16", th, 2000);
    run<1>("ldmatrix.x4", th, 2000);
    run<2>("stmatrix.x4.trans", th, 2000);
    run<3>("movmatrix.trans", th, 2000);
    run<4>("ldsm.x4.trans + 4 mma", th, 2000);
  }
  return 0;
}
