// Does tcgen05.mma accept a K-major, 128-byte-swizzled A operand whose start address is shifted by whole 128-byte rows
// (not a multiple of the 1024-byte swizzle atom)?  This is what an implicit-GEMM convolution needs: one shared-memory
// copy of an image row [pixel][64 channels] serves the three horizontal taps through three start addresses.
// The tile is written by the threads with the swizzle derived from the ABSOLUTE shared-memory address
// (16-byte chunk index ^ ((address >> 7) & 7)); D[m][n] = sum_k A[m + shift][k] B[n][k] is compared with the host.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/umma_rowshift tools/ubench/umma_rowshift.cu
#include <cuda_bf16.h>
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kRowsA = 160, kK = 64, kN = 64, kM = 128;

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t addr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_offset & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) test_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, float* d, int shift,
                                                      int use_base_offset) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_s;
  const uint32_t base = (s32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - s32(smem_raw));
  unsigned char* sa = sm;                        // [160 rows][128 B]
  unsigned char* sb = sm + kRowsA * 128;         // 160 * 128 = 20480 = 20 * 1024: aligned
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kRowsA * 8; i += 128) {  // 16-byte chunks
    const int r = i >> 3, c = i & 7;
    const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(sa + off) = *reinterpret_cast<const uint4*>(a + r * kK + c * 8);
  }
  for (int i = tid; i < kN * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(sb + off) = *reinterpret_cast<const uint4*>(b + r * kK + c * 8);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s32(&tmem_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy (tensor core) reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = tmem_s;
  if (tid == 0) {
    uint32_t idesc = 0;
    idesc |= 1u << 4, idesc |= 1u << 7, idesc |= 1u << 10;
    idesc |= static_cast<uint32_t>(kN >> 3) << 17;
    idesc |= static_cast<uint32_t>(kM >> 4) << 24;
    const uint32_t a_addr = base + shift * 128;
    const uint64_t da = desc_k_sw128(a_addr, use_base_offset ? ((a_addr >> 7) & 7) : 0);
    const uint64_t db = desc_k_sw128(base + kRowsA * 128, 0);
    for (int k = 0; k < kK / 16; ++k) {
      const uint32_t acc = k != 0;
      asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tacc),
                   "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(s32(&bar))
                 : "memory");
  } while (!done);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < kN; c0 += 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tacc + (static_cast<uint32_t>(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) d[tid * kN + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tacc) : "memory");
}


// throughput: `reps` back-to-back K = 64 products (4 instructions each) from one thread, A start shifted by `shift` rows
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int shift, int reps, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_s;
  const uint32_t base = (s32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (kRowsA * 128 + 256 * 128) / 16; i += 128) reinterpret_cast<uint4*>(smem_raw + (base - s32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s32(&tmem_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = tmem_s;
  if (warp == 0) {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    if (pred) {
      uint32_t idesc = 0;
      idesc |= 1u << 4, idesc |= 1u << 7, idesc |= 1u << 10;
      idesc |= static_cast<uint32_t>(N >> 3) << 17;
      idesc |= static_cast<uint32_t>(kM >> 4) << 24;
      const uint64_t da = desc_k_sw128(base + shift * 128, 0);
      const uint64_t db = desc_k_sw128(base + kRowsA * 128, 0);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tacc),
                       "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(1u)
                       : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      uint32_t done;
      do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(s32(&bar))
                     : "memory");
      } while (!done);
      *cycles = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tacc) : "memory");
}

template <int N>
void rate(int shift) {
  long long* dc;
  cudaMalloc(&dc, 8);
  const int smem = kRowsA * 128 + 256 * 128 + 1024, reps = 2000;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<N><<<1, 128, smem>>>(shift, reps, dc);
  cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("M 128 N %3d K 16, A start + %d rows: %.1f cycles per instruction (floor N/2 = %d) %s\n", N, shift, double(c) / (reps * 4), N / 2,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(dc);
}

// Where do the 64 rows of an M = 64 accumulator live in tensor memory?  D[m][n] = m + 1 for all n; every lane is dumped.
__global__ void __launch_bounds__(128, 1) m64_kernel(float* d) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_s;
  const uint32_t base = (s32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - s32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (64 * 128 + 64 * 128) / 2; i += 128) reinterpret_cast<uint16_t*>(sm)[i] = 0;
  __syncthreads();
  if (tid < 64) {  // A[m][k = 0] = m + 1 (bf16 exact up to 256), B[n][k = 0] = 1; logical chunk 0 sits at chunk (r & 7)
    reinterpret_cast<__nv_bfloat16*>(sm + tid * 128 + ((0 ^ (tid & 7)) << 4))[0] = __float2bfloat16(float(tid + 1));
    reinterpret_cast<__nv_bfloat16*>(sm + 64 * 128 + tid * 128 + ((0 ^ (tid & 7)) << 4))[0] = __float2bfloat16(1.f);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s32(&tmem_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = tmem_s;
  if (tid == 0) {
    uint32_t idesc = 0;
    idesc |= 1u << 4, idesc |= 1u << 7, idesc |= 1u << 10;
    idesc |= static_cast<uint32_t>(64 >> 3) << 17;
    idesc |= static_cast<uint32_t>(64 >> 4) << 24;
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tacc),
                 "l"(desc_k_sw128(base, 0)), "l"(desc_k_sw128(base + 64 * 128, 0)), "r"(idesc), "r"(0u)
                 : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(s32(&bar))
                 : "memory");
  } while (!done);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(tacc + (static_cast<uint32_t>(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  d[tid] = __uint_as_float(r[0]);
  d[128 + tid] = __uint_as_float(r[5]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tacc) : "memory");
}

void m64_layout() {
  float* dd;
  cudaMalloc(&dd, 256 * 4);
  cudaMemset(dd, 0, 256 * 4);
  const int smem = 2 * 64 * 128 + 1024;
  m64_kernel<<<1, 128, smem>>>(dd);
  cudaDeviceSynchronize();
  float h[256];
  cudaMemcpy(h, dd, sizeof(h), cudaMemcpyDeviceToHost);
  printf("M = 64 accumulator: value (= row + 1) seen by TMEM lane 0..127, column 0 [%s]\n", cudaGetErrorString(cudaGetLastError()));
  for (int l = 0; l < 128; ++l) printf("%g%s", h[l], (l & 31) == 31 ? "\n" : " ");
  printf("column 5:\n");
  for (int l = 0; l < 128; ++l) printf("%g%s", h[128 + l], (l & 31) == 31 ? "\n" : " ");
  cudaFree(dd);
}

int main() {
  m64_layout();
  for (int sh = 0; sh <= 2; ++sh) rate<64>(sh), rate<128>(sh), rate<256>(sh);

  std::vector<__nv_bfloat16> ha(kRowsA * kK), hb(kN * kK);
  std::vector<float> fa(kRowsA * kK), fb(kN * kK);
  srand(3);
  for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (rand() % 17 - 8) / 8.f; ha[i] = __float2bfloat16(fa[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (rand() % 13 - 6) / 4.f; hb[i] = __float2bfloat16(fb[i]); }
  __nv_bfloat16 *da, *db;
  float* dd;
  cudaMalloc(&da, ha.size() * 2), cudaMalloc(&db, hb.size() * 2), cudaMalloc(&dd, kM * kN * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = kRowsA * 128 + kN * 128 + 1024;
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hd(kM * kN);
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int shift = 0; shift <= 9; ++shift) {
      cudaMemset(dd, 0, kM * kN * 4);
      test_kernel<<<1, 128, smem>>>(da, db, dd, shift, use_bo);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(hd.data(), dd, kM * kN * 4, cudaMemcpyDeviceToHost);
      double worst = 0;
      for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
          double r = 0;
          for (int k = 0; k < kK; ++k) r += static_cast<double>(fa[(m + shift) * kK + k]) * fb[n * kK + k];
          const double dlt = fabs(r - hd[m * kN + n]);
          if (dlt > worst) worst = dlt;
        }
      printf("base_offset field %s, row shift %d: max |d - ref| = %.4g  (%s)\n", use_bo ? "set " : "zero", shift, worst,
             cudaGetErrorString(e));
    }
  return 0;
}
