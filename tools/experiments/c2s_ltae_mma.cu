// Tensor-core L-TAE forward for the shipped shapes (bf16 I/O, n_head = 16, d_model = 256, C in {64, 128},
// T <= 64): same contract as the general kernel in c2s_ltae.cu, selected by c2s_ltae_forward.
//
// Reference: LTAE.forward / LTAE4WTAE.forward (src/backbones/tae.py:451-504, 589-635).
//
// One CTA = 8 consecutive pixels, one warp per pixel for the attention part.
//   1. staging   x[b, t, c, pix0..pix0+7] (16-byte segments) is read from HBM ONCE and written transposed
//                into shared memory as X2[p][c][t] (bf16, t contiguous), GroupNorm statistics are accumulated
//                on the way (tae.py:461).  Frames known to be zero are never read.
//   2. scores    S^T[h, t] = U'[h, c] X2[c, t]  per pixel on the tensor cores (m16n8k16, M = 16 heads).
//                U' = U * rstd(group, pixel) is rebuilt per pixel in registers and split into bf16 hi + lo,
//                so the product keeps ~16 mantissa bits (x itself is exact in bf16).       tae.py:827-831
//   3. softmax   over T in the accumulator registers (pad -> -1e6, as the reference).       tae.py:836
//   4. values    z[h, c] = a[h, t] X2^T[t, c]: the probability accumulators are re-used as the A operand
//                (hi + lo); two extra n-tiles carry the positional table, so sum_t a PE comes for free.
//   5. epilogue  GroupNorm affine on z, per-head in-projection, MLP, BatchNorm, ReLU, output GroupNorm
//                (tae.py:463, 486-488) with N = 8 pixels on the tensor cores (weights pre-split hi/lo).
// The per-pixel operands make steps 2 and 4 batched 16xK GEMMs (A differs per pixel), which is the shape
// mma.sync m16n8k16 fits exactly; tcgen05's M >= 64 tiles would be 3/4 empty.  See DESIGN.md.
#include <cstdlib>

#include "c2s_ltae_prep.cuh"

namespace c2s {
namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kPix = 8;        // pixels per CTA; two warps share a pixel in the attention part
constexpr int kTP = 64;        // padded frame count
constexpr int kRow = kTP + 8;  // X2 row pitch in elements (144 B: ldmatrix rows hit distinct banks)
constexpr int kH = 16;         // heads
constexpr int kD = 256;        // d_model
constexpr int kAP = 72;            // pitch of the [h][t] fp32 tiles: 64-bit accesses of a half-warp hit 32 distinct banks
constexpr int kAsP = kH * kAP + 4;  // attention staging: floats per pixel ([h][kAP] + pad)
constexpr int kOsRow = kD + 8;     // o_s row pitch (elements)

struct MmaArgs {
  const __nv_bfloat16* x;
  const uint8_t* pad;
  __nv_bfloat16* out;
  float* attn;
  const float* ufrag;     // [C/16][32][8]
  const uint4* wcfrag;    // [16][C/16][32][2]
  const uint4* wmfrag;    // [c_out/16][16][32][2]
  const float* cpos;      // [B, T, 16]
  const float* pe;        // [B, T, 256] or nullptr
  const float* bc;        // [256]
  const float* bm;        // [c_out]
  const float* gamma;
  const float* beta;
  const float* bnf;       // [2, c_out] or nullptr (training)
  const float* on_w;
  const float* on_b;
  float* ypre;
  const uint8_t* attn_keep;  // [16, B, T, hw] dropout keep mask or nullptr
  const uint8_t* mlp_keep;   // [B, c_out, hw] or nullptr
  float attn_keep_scale, mlp_keep_scale;
  __nv_bfloat16* o_hi;       // [B*hw][256] rows for the tcgen05 MLP kernel (nullptr: MLP runs in this kernel)
  __nv_bfloat16* o_lo;
  float* save_o;             // [B*hw][256] fp32 rows for the backward, or nullptr
  int B, T, hw, c_out;
  int attn_only, skip_attn_store, zero_padded;
  float gn_eps;
  int tiles_per_b;
};

template <int C>
struct Smem {
  static constexpr int kX2 = kPix * C * kRow * 2;               // bytes
  static constexpr int kZH = kPix * (C + 8) + 8;                // elements per head of z_hi / z_lo
  static constexpr int kZ = 2 * kH * kZH * 2;                   // bytes (aliases X2)
  static_assert(kZ <= kX2, "z tiles must fit in the X2 region");
  static constexpr int oX2 = 0;
  static constexpr int oAs = oX2 + kX2;                         // float [8][kAsP]
  static constexpr int oCpos = oAs + kPix * kAsP * 4;           // float [16][kAP]
  static constexpr int oPeHi = oCpos + kH * kAP * 4;             // bf16 [16][kRow]
  static constexpr int oPeLo = oPeHi + 16 * kRow * 2;
  static constexpr int oRstd = oPeLo + 16 * kRow * 2;           // float [16][8]
  static constexpr int oMu = oRstd + kH * kPix * 4;             // float [16][8]  mean * rstd
  static constexpr int oSa = oMu + kH * kPix * 4;               // float [16][8]
  static constexpr int oPa = oSa + kH * kPix * 4;               // float [16][16][8]
  static constexpr int oOsHi = oPa + kH * 16 * kPix * 4;        // bf16 [8][kOsRow]
  static constexpr int oOsLo = oOsHi + kPix * kOsRow * 2;
  static constexpr int oYs = oOsLo + kPix * kOsRow * 2;         // float [256][8]
  static constexpr int oUf = oYs + 256 * kPix * 4;              // float [C/16][32][8] score weights (A-fragment order)
  static constexpr int oRed = oUf + (C / 16) * 32 * 8 * 4;      // float [8 pixels][2 warps][max | sum][16] softmax exchange
  static constexpr int kTotal = oRed + kPix * 2 * 2 * kH * 4;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// D += A(16x16, row) * B(16x8, col), bf16 inputs, fp32 accumulate
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// v = hi + lo with hi, lo bf16: ~16 mantissa bits survive
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(v0, v1);
  const float r0 = v0 - __uint_as_float(hi << 16);
  const float r1 = v1 - __uint_as_float(hi & 0xffff0000u);
  lo = pack_bf16(r0, r1);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

template <int C>
__global__ void __launch_bounds__(kThreads, 1) ltae_mma_kernel(const MmaArgs a) {
  using S = Smem<C>;
  constexpr int CPG = C / kH;   // channels per GroupNorm group (8 or 4)
  constexpr int KS = C / 16;    // k-steps over channels
  extern __shared__ __align__(128) unsigned char smem[];
  __nv_bfloat16* X2 = reinterpret_cast<__nv_bfloat16*>(smem + S::oX2);
  float* s_as = reinterpret_cast<float*>(smem + S::oAs);
  float* s_cpos = reinterpret_cast<float*>(smem + S::oCpos);
  __nv_bfloat16* s_pe_hi = reinterpret_cast<__nv_bfloat16*>(smem + S::oPeHi);
  __nv_bfloat16* s_pe_lo = reinterpret_cast<__nv_bfloat16*>(smem + S::oPeLo);
  float* s_rstd = reinterpret_cast<float*>(smem + S::oRstd);
  float* s_mu = reinterpret_cast<float*>(smem + S::oMu);
  float* s_sa = reinterpret_cast<float*>(smem + S::oSa);
  float* s_pa = reinterpret_cast<float*>(smem + S::oPa);
  __nv_bfloat16* s_os_hi = reinterpret_cast<__nv_bfloat16*>(smem + S::oOsHi);
  __nv_bfloat16* s_os_lo = reinterpret_cast<__nv_bfloat16*>(smem + S::oOsLo);
  float* s_ys = reinterpret_cast<float*>(smem + S::oYs);
  float* s_uf = reinterpret_cast<float*>(smem + S::oUf);
  float* s_red = reinterpret_cast<float*>(smem + S::oRed);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / a.tiles_per_b;
  const int pix0 = (blockIdx.x - b * a.tiles_per_b) * kPix;
  const size_t frame_stride = static_cast<size_t>(C) * a.hw;
  const __nv_bfloat16* xb = a.x + static_cast<size_t>(b) * a.T * frame_stride + pix0;

  // ---- phase 0: frame masks (every warp derives them itself: no block barrier before the loads go out) ----
  unsigned long long live_mask = 0, pad_mask = 0;
#pragma unroll
  for (int base = 0; base < kTP; base += 32) {
    const int t = base + lane;
    const bool pd = t < a.T && a.pad != nullptr && __ldg(a.pad + b * a.T + t) != 0;
    const bool lv = t < a.T && !(pd && a.zero_padded);
    live_mask |= static_cast<unsigned long long>(__ballot_sync(0xffffffffu, lv)) << base;
    pad_mask |= static_cast<unsigned long long>(__ballot_sync(0xffffffffu, pd)) << base;
  }
  const int n_live = __popcll(live_mask);

  // ---- phase 1: stage x transposed into X2[p][c][t] and accumulate GroupNorm statistics -------------
  {
    constexpr int NCB = (C / 8 + kWarps - 1) / kWarps;  // 8-channel blocks per warp
    const int cc = lane & 7, tp = lane >> 3;
    // 1a: every global load of this warp's share is issued before anything is consumed (one latency per CTA);
    // the few loads of the per-sample constants go first so that they are not queued behind the features
    static_assert(kH * kTP == 2 * kThreads && 16 * kTP == 2 * kThreads && (C / 16) * 64 <= kThreads, "fill mapping");
    float cpos_r[2], pe_r[2];
    float4 uf_r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = tid + q * kThreads, t = i / kH, h = i - t * kH;
      cpos_r[q] = t < a.T ? __ldg(a.cpos + (static_cast<size_t>(b) * a.T + t) * kMaxHeads + h) : 0.f;
      const int t2 = i / 16, d = i - t2 * 16;
      pe_r[q] = (!a.attn_only && a.pe != nullptr && t2 < a.T) ? __ldg(a.pe + (static_cast<size_t>(b) * a.T + t2) * kD + d) : 0.f;
    }
    if (tid < KS * 64) uf_r = __ldg(reinterpret_cast<const float4*>(a.ufrag) + tid);
    uint4 v[NCB][8][2];
    uint4 pv[NCB];
#pragma unroll
    for (int n = 0; n < NCB; ++n) {
      const int cb = warp + kWarps * n;
      const bool has_work = cb < C / 8;  // warp-uniform
      const int c = cb * 8 + cc;
      const int g = c / CPG;
      const __nv_bfloat16* xc = xb + static_cast<size_t>(c) * a.hw;
      pv[n] = make_uint4(0, 0, 0, 0);
      if (n_live > 0 && has_work) {  // pivot of the shifted sums: first live frame, first channel of the group
        const int t0 = __ffsll(static_cast<long long>(live_mask)) - 1;
        pv[n] = ld_stream_v4(xb + static_cast<size_t>(t0) * frame_stride + static_cast<size_t>(g * CPG) * a.hw);
      }
#pragma unroll
      for (int tb = 0; tb < 8; ++tb) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int t = tb * 8 + tp * 2 + e;
          v[n][tb][e] = make_uint4(0, 0, 0, 0);
          if (has_work && ((live_mask >> t) & 1ull)) v[n][tb][e] = ld_stream_v4(xc + static_cast<size_t>(t) * frame_stride);
        }
      }
    }
    // 1b: per-sample constants (their loads were issued ahead of the feature loads) go to shared memory
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = tid + q * kThreads, t = i / kH, h = i - t * kH;
      s_cpos[h * kAP + t] = cpos_r[q];
    }
    if (tid < KS * 64) {
      const int ks = tid >> 6, lane_e = tid & 63;  // source order [ks][lane][2] -> [ks][2][lane]
      reinterpret_cast<float4*>(s_uf)[ks * 64 + (lane_e & 1) * 32 + (lane_e >> 1)] = uf_r;
    }
    if (!a.attn_only) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = tid + q * kThreads, t = i / 16, d = i - t * 16;
        const __nv_bfloat16 hi = __float2bfloat16_rn(pe_r[q]);
        s_pe_hi[d * kRow + t] = hi;
        s_pe_lo[d * kRow + t] = __float2bfloat16_rn(pe_r[q] - __bfloat162float(hi));
      }
    }
    // 1c: statistics + transposed stores
#pragma unroll
    for (int n = 0; n < NCB; ++n) {
      if (warp + kWarps * n >= C / 8) continue;  // warp-uniform
      const int c = (warp + kWarps * n) * 8 + cc;
      const int g = c / CPG;
      float pivot[8];
      Elem<__nv_bfloat16>::unpack(pv[n], pivot);
      float s1[8], s2[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) s1[p] = 0.f, s2[p] = 0.f;
      __nv_bfloat16* xrow = X2 + static_cast<size_t>(c) * kRow;
#pragma unroll
      for (int tb = 0; tb < 8; ++tb) {
        const int t = tb * 8 + tp * 2;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if ((live_mask >> (t + e)) & 1ull) {
            float f[8];
            Elem<__nv_bfloat16>::unpack(v[n][tb][e], f);
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              const float d = f[p] - pivot[p];
              s1[p] += d;
              s2[p] = fmaf(d, d, s2[p]);
            }
          }
        }
        const uint32_t wa[4] = {v[n][tb][0].x, v[n][tb][0].y, v[n][tb][0].z, v[n][tb][0].w};
        const uint32_t wb[4] = {v[n][tb][1].x, v[n][tb][1].y, v[n][tb][1].z, v[n][tb][1].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // pixel 2i: low halves of (frame t, frame t+1); pixel 2i+1: high halves
          *reinterpret_cast<uint32_t*>(xrow + static_cast<size_t>(2 * i) * C * kRow + t) = __byte_perm(wa[i], wb[i], 0x5410);
          *reinterpret_cast<uint32_t*>(xrow + static_cast<size_t>(2 * i + 1) * C * kRow + t) = __byte_perm(wa[i], wb[i], 0x7632);
        }
      }
      // reduce over the lanes that share a group: all 32 (CPG = 8) or the 16 with the same cc / 4 (CPG = 4)
#pragma unroll
      for (int p = 0; p < 8; ++p) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
          if (CPG == 4 && o == 4) continue;
          s1[p] += __shfl_xor_sync(0xffffffffu, s1[p], o);
          s2[p] += __shfl_xor_sync(0xffffffffu, s2[p], o);
        }
      }
      const int writer = (CPG == 8) ? lane : ((cc & 3) + 4 * tp);  // index within the group's lanes
      if (writer < 8) {
        float t1 = 0.f, t2 = 0.f, pvt = 0.f;
#pragma unroll
        for (int p = 0; p < 8; ++p)
          if (writer == p) t1 = s1[p], t2 = s2[p], pvt = pivot[p];
        const float n_all = static_cast<float>(a.T) * CPG;
        const float n_skip = n_all - static_cast<float>(n_live) * CPG;  // frames known to be zero
        t1 -= n_skip * pvt;
        t2 = fmaf(n_skip * pvt, pvt, t2);
        const float m = t1 / n_all;
        float var = t2 / n_all - m * m;
        var = var < 0.f ? 0.f : var;
        const float rstd = 1.f / sqrtf(var + a.gn_eps);
        s_rstd[g * kPix + writer] = rstd;
        s_mu[g * kPix + writer] = (m + pvt) * rstd;
      }
    }
  }
  __syncthreads();

  // ---- phases 2-4: two warps per pixel ------------------------------------------------------------
  // scores: each warp of the pair takes half of the frames (4 n-tiles); softmax statistics are exchanged
  // through shared memory; values: each warp takes half of the channels (and one positional n-tile).
  const int p = warp >> 1, half = warp & 1;
  const int j = lane & 3, r8 = lane >> 2;  // fragment coordinates: row r8 (and r8 + 8), column pair 2j
  const uint32_t x2p = smem_u32(X2 + static_cast<size_t>(p) * C * kRow);
  const uint32_t pair_bar = 1 + p;         // named barrier of this pixel's two warps
  float sacc[4][4];  // S^T[h, t]: n-tile nt covers t = 8 (4 half + nt) .. + 7
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) sacc[nt][i] = 0.f;
  float mh0 = 0.f, mh1 = 0.f;  // sum_c U'[h, c] mu[g(c)] for rows r8 and r8 + 8
  {
    const int mat = lane >> 3, mr = lane & 7;
#pragma unroll 2
    for (int ks = 0; ks < KS; ++ks) {
      const float4* up = reinterpret_cast<const float4*>(s_uf) + ks * 64 + lane;  // [ks][2][32 lanes]: 16 B lane stride
      const float4 u0 = up[0], u1 = up[32];
      const int c_lo = ks * 16 + 2 * j;
      const int g0 = c_lo / CPG, g1 = (c_lo + 8) / CPG;
      const float r0 = s_rstd[g0 * kPix + p], r1 = s_rstd[g1 * kPix + p];
      const float m0 = s_mu[g0 * kPix + p], m1 = s_mu[g1 * kPix + p];
      uint32_t ahi[4], alo[4];
      split2(u0.x * r0, u0.y * r0, ahi[0], alo[0]);  // (row r8,     k 2j, 2j+1)
      split2(u0.z * r0, u0.w * r0, ahi[1], alo[1]);  // (row r8 + 8, k 2j, 2j+1)
      split2(u1.x * r1, u1.y * r1, ahi[2], alo[2]);  // (row r8,     k 2j+8, 2j+9)
      split2(u1.z * r1, u1.w * r1, ahi[3], alo[3]);  // (row r8 + 8, k 2j+8, 2j+9)
      mh0 += (u0.x + u0.y) * m0 + (u1.x + u1.y) * m1;
      mh1 += (u0.z + u0.w) * m0 + (u1.z + u1.w) * m1;
#pragma unroll
      for (int ntp = 0; ntp < 2; ++ntp) {
        uint32_t bfr[4];
        const int c = (2 * ks + (mat & 1)) * 8 + mr;
        const int t0 = (4 * half + 2 * ntp + (mat >> 1)) * 8;
        ldmatrix_x4_trans(bfr, x2p + static_cast<uint32_t>(c * kRow + t0) * 2u);
        mma_bf16(sacc[2 * ntp], ahi, bfr[0], bfr[1]);
        mma_bf16(sacc[2 * ntp], alo, bfr[0], bfr[1]);
        mma_bf16(sacc[2 * ntp + 1], ahi, bfr[2], bfr[3]);
        mma_bf16(sacc[2 * ntp + 1], alo, bfr[2], bfr[3]);
      }
    }
  }
  mh0 += __shfl_xor_sync(0xffffffffu, mh0, 1);
  mh0 += __shfl_xor_sync(0xffffffffu, mh0, 2);
  mh1 += __shfl_xor_sync(0xffffffffu, mh1, 1);
  mh1 += __shfl_xor_sync(0xffffffffu, mh1, 2);

  // softmax over t for rows h = r8 and r8 + 8                                        tae.py:831-836
  const bool store_attn = a.attn != nullptr && !a.skip_attn_store;
  float* as = s_as + p * kAsP;                              // probabilities [h][66] of this pixel
  float* red = s_red + (p * 2 + half) * 2 * kH;              // [max | sum][h] of this warp
  const float* red_other = s_red + (p * 2 + (half ^ 1)) * 2 * kH;
  {
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int t = (4 * half + nt) * 8 + 2 * j + e;
        float v0 = sacc[nt][e] + s_cpos[r8 * kAP + t] - mh0;
        float v1 = sacc[nt][2 + e] + s_cpos[(r8 + 8) * kAP + t] - mh1;
        if ((pad_mask >> t) & 1ull) v0 = -1e6f, v1 = -1e6f;
        if (t >= a.T) v0 = -INFINITY, v1 = -INFINITY;
        sacc[nt][e] = v0, sacc[nt][2 + e] = v1;
        mx0 = fmaxf(mx0, v0), mx1 = fmaxf(mx1, v1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    if (j == 0) red[r8] = mx0, red[r8 + 8] = mx1;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    mx0 = fmaxf(mx0, red_other[r8]);       // the other half of the frames (T >= 1: at least one side is finite)
    mx1 = fmaxf(mx1, red_other[r8 + 8]);
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sacc[nt][e] = expf(sacc[nt][e] - mx0);
        sacc[nt][2 + e] = expf(sacc[nt][2 + e] - mx1);
        d0 += sacc[nt][e], d1 += sacc[nt][2 + e];
      }
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    if (j == 0) red[kH + r8] = d0, red[kH + r8 + 8] = d1;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    // the sum is formed in the same order by both warps so that they normalise identically
    const float lo0 = half ? red_other[kH + r8] : d0, hi0 = half ? d0 : red_other[kH + r8];
    const float lo1 = half ? red_other[kH + r8 + 8] : d1, hi1 = half ? d1 : red_other[kH + r8 + 8];
    const float inv0 = 1.f / (lo0 + hi0), inv1 = 1.f / (lo1 + hi1);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int t = (4 * half + nt) * 8 + 2 * j;
      float2 q0 = make_float2(sacc[nt][0] * inv0, sacc[nt][1] * inv0);
      float2 q1 = make_float2(sacc[nt][2] * inv1, sacc[nt][3] * inv1);
      if (a.attn_keep != nullptr) {  // training: dropout acts on the attention that is returned (tae.py:837)
        const uint8_t* k0 = a.attn_keep + ((static_cast<size_t>(r8) * a.B + b) * a.T + t) * a.hw + pix0 + p;
        const uint8_t* k1 = a.attn_keep + ((static_cast<size_t>(r8 + 8) * a.B + b) * a.T + t) * a.hw + pix0 + p;
        const float sc = a.attn_keep_scale;
        q0.x = (t < a.T && k0[0]) ? q0.x * sc : 0.f;
        q0.y = (t + 1 < a.T && k0[a.hw]) ? q0.y * sc : 0.f;
        q1.x = (t < a.T && k1[0]) ? q1.x * sc : 0.f;
        q1.y = (t + 1 < a.T && k1[a.hw]) ? q1.y * sc : 0.f;
      }
      *reinterpret_cast<float2*>(as + r8 * kAP + t) = q0;
      *reinterpret_cast<float2*>(as + (r8 + 8) * kAP + t) = q1;
    }
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
  }

  // values: zr[h, c] = sum_t a[h, t] x[t, c]  (+ 16 positional columns)                 tae.py:839
  constexpr int ZT = C / 16;  // channel n-tiles per warp
  float zacc[ZT][4];
  float pacc[4];
  float sa0 = 0.f, sa1 = 0.f;
  if (!a.attn_only) {
#pragma unroll
    for (int nt = 0; nt < ZT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) zacc[nt][i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) pacc[i] = 0.f;
    const int mat = lane >> 3, mr = lane & 7;
    const uint32_t pe_base = smem_u32((mat >> 1) ? s_pe_lo : s_pe_hi);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      // A fragment of the probabilities of ALL frames of this k-step (written by both warps of the pair)
      const float2 p0 = *reinterpret_cast<const float2*>(as + r8 * kAP + ks * 16 + 2 * j);
      const float2 p1 = *reinterpret_cast<const float2*>(as + (r8 + 8) * kAP + ks * 16 + 2 * j);
      const float2 p2 = *reinterpret_cast<const float2*>(as + r8 * kAP + ks * 16 + 8 + 2 * j);
      const float2 p3 = *reinterpret_cast<const float2*>(as + (r8 + 8) * kAP + ks * 16 + 8 + 2 * j);
      sa0 += (p0.x + p0.y) + (p2.x + p2.y);
      sa1 += (p1.x + p1.y) + (p3.x + p3.y);
      uint32_t ahi[4], alo[4];
      split2(p0.x, p0.y, ahi[0], alo[0]);
      split2(p1.x, p1.y, ahi[1], alo[1]);
      split2(p2.x, p2.y, ahi[2], alo[2]);
      split2(p3.x, p3.y, ahi[3], alo[3]);
      const int t0 = (2 * ks + (mat & 1)) * 8;
#pragma unroll
      for (int ntp = 0; ntp < ZT / 2; ++ntp) {
        uint32_t bfr[4];
        const int c = (half * ZT + 2 * ntp + (mat >> 1)) * 8 + mr;
        ldmatrix_x4(bfr, x2p + static_cast<uint32_t>(c * kRow + t0) * 2u);
        mma_bf16(zacc[2 * ntp], ahi, bfr[0], bfr[1]);
        mma_bf16(zacc[2 * ntp], alo, bfr[0], bfr[1]);
        mma_bf16(zacc[2 * ntp + 1], ahi, bfr[2], bfr[3]);
        mma_bf16(zacc[2 * ntp + 1], alo, bfr[2], bfr[3]);
      }
      if (a.pe != nullptr) {  // positional n-tile `half`: matrices (hi, 2ks), (hi, 2ks+1), (lo, 2ks), (lo, 2ks+1)
        uint32_t bp[4];
        ldmatrix_x4(bp, pe_base + static_cast<uint32_t>((half * 8 + mr) * kRow + t0) * 2u);
        mma_bf16(pacc, ahi, bp[0], bp[1]);
        mma_bf16(pacc, alo, bp[0], bp[1]);
        mma_bf16(pacc, ahi, bp[2], bp[3]);
      }
    }
    sa0 += __shfl_xor_sync(0xffffffffu, sa0, 1);
    sa0 += __shfl_xor_sync(0xffffffffu, sa0, 2);
    sa1 += __shfl_xor_sync(0xffffffffu, sa1, 1);
    sa1 += __shfl_xor_sync(0xffffffffu, sa1, 2);
    if (j == 0 && half == 0) s_sa[r8 * kPix + p] = sa0, s_sa[(r8 + 8) * kPix + p] = sa1;
  }
  __syncthreads();  // every warp is done with X2 (z tiles alias it) and the attention staging is complete

  if (store_attn) {  // attn[h, b, t, pix0 .. pix0 + 7]: 32-byte segments                    tae.py:490-493
    const int pp = lane & 7, tq = lane >> 3;
    for (int h = warp; h < kH; h += kThreads / 32) {
      float* dst = a.attn + ((static_cast<size_t>(h) * a.B + b) * a.T + tq) * a.hw + pix0 + pp;
      const float* src = s_as + pp * kAsP + h * kAP + tq;
      const size_t step = static_cast<size_t>(4) * a.hw;
#pragma unroll 4
      for (int t = tq; t < a.T; t += 4, src += 4, dst += step) *dst = *src;
    }
  }
  if (a.attn_only) return;

  // GroupNorm affine on the weighted sums, bf16 hi/lo tiles z[h][p][c] for the in-projection      tae.py:461
  __nv_bfloat16* z_hi = reinterpret_cast<__nv_bfloat16*>(smem + S::oX2);
  __nv_bfloat16* z_lo = z_hi + kH * S::kZH;
  {
#pragma unroll
    for (int nt = 0; nt < ZT; ++nt) {
      const int c = (half * ZT + nt) * 8 + 2 * j;
      const int g = c / CPG;
      const float r = s_rstd[g * kPix + p], m = s_mu[g * kPix + p];
      const float2 gm = __ldg(reinterpret_cast<const float2*>(a.gamma + c));
      const float2 bt = __ldg(reinterpret_cast<const float2*>(a.beta + c));
      // sum_t a (x rstd - mu rstd) gamma + beta sum_t a
      const float z00 = fmaf(gm.x, fmaf(zacc[nt][0], r, -m * sa0), bt.x * sa0);
      const float z01 = fmaf(gm.y, fmaf(zacc[nt][1], r, -m * sa0), bt.y * sa0);
      const float z10 = fmaf(gm.x, fmaf(zacc[nt][2], r, -m * sa1), bt.x * sa1);
      const float z11 = fmaf(gm.y, fmaf(zacc[nt][3], r, -m * sa1), bt.y * sa1);
      uint32_t hi, lo;
      split2(z00, z01, hi, lo);
      const int o0 = r8 * S::kZH + p * (C + 8) + c;
      *reinterpret_cast<uint32_t*>(z_hi + o0) = hi;
      *reinterpret_cast<uint32_t*>(z_lo + o0) = lo;
      split2(z10, z11, hi, lo);
      const int o1 = (r8 + 8) * S::kZH + p * (C + 8) + c;
      *reinterpret_cast<uint32_t*>(z_hi + o1) = hi;
      *reinterpret_cast<uint32_t*>(z_lo + o1) = lo;
    }
    const int i0 = half * 8 + 2 * j;
    s_pa[(r8 * 16 + i0) * kPix + p] = pacc[0];
    s_pa[(r8 * 16 + i0 + 1) * kPix + p] = pacc[1];
    s_pa[((r8 + 8) * 16 + i0) * kPix + p] = pacc[2];
    s_pa[((r8 + 8) * 16 + i0 + 1) * kPix + p] = pacc[3];
  }
  __syncthreads();

  // ---- phase 5a: o[h*16 + i, p] = Wc[h*16 + i, :] . z[h, :, p] + sa[h,p] bc + sum_t a PE ---- tae.py:463,479,839
  for (int h = warp; h < kH; h += kThreads / 32) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const __nv_bfloat16* zh = z_hi + h * S::kZH + r8 * (C + 8) + 2 * j;  // B[k = c][n = pixel r8]
    const __nv_bfloat16* zl = z_lo + h * S::kZH + r8 * (C + 8) + 2 * j;
#pragma unroll 2
    for (int ks = 0; ks < KS; ++ks) {
      const uint4* wp = a.wcfrag + (static_cast<size_t>(h * KS + ks) * 32 + lane) * 2;
      const uint4 wh = __ldg(wp), wl = __ldg(wp + 1);
      const uint32_t ahi[4] = {wh.x, wh.y, wh.z, wh.w};
      const uint32_t alo[4] = {wl.x, wl.y, wl.z, wl.w};
      const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(zh + ks * 16);
      const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(zh + ks * 16 + 8);
      const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(zl + ks * 16);
      const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(zl + ks * 16 + 8);
      mma_bf16(acc, ahi, bh0, bh1);
      mma_bf16(acc, alo, bh0, bh1);
      mma_bf16(acc, ahi, bl0, bl1);
    }
    // accumulator: rows i = r8, r8 + 8; columns pixel 2j, 2j + 1
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = r8 + (e >> 1) * 8, pp = 2 * j + (e & 1);
      const int d = h * 16 + i;
      const float v = acc[e] + s_sa[h * kPix + pp] * __ldg(a.bc + d) + s_pa[(h * 16 + i) * kPix + pp];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      s_os_hi[pp * kOsRow + d] = hi;
      s_os_lo[pp * kOsRow + d] = __float2bfloat16_rn(v - __bfloat162float(hi));
      if (a.save_o != nullptr) a.save_o[(static_cast<size_t>(b) * a.hw + pix0 + pp) * kD + d] = v;
    }
  }
  __syncthreads();

  if (a.o_hi != nullptr) {  // the MLP, BatchNorm, ReLU and output GroupNorm run as a tcgen05 row GEMM (c2s_ltae_mlp_tc.cu)
    const size_t row0 = static_cast<size_t>(b) * a.hw + pix0;
    for (int i = tid; i < 2 * kPix * (kD / 8); i += kThreads) {  // 16-byte pieces of the 8 rows, hi then lo
      const int plane = i / (kPix * (kD / 8)), r = i - plane * (kPix * (kD / 8));
      const int pp = r / (kD / 8), q = r - pp * (kD / 8);
      const __nv_bfloat16* src = (plane ? s_os_lo : s_os_hi) + pp * kOsRow + q * 8;
      __nv_bfloat16* dst = (plane ? a.o_lo : a.o_hi) + (row0 + pp) * kD + q * 8;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
    }
    return;
  }

  // ---- phase 5b: MLP Linear (+ eval BatchNorm + ReLU)                                        tae.py:442-447
  for (int mt = warp; mt < a.c_out / 16; mt += kThreads / 32) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const __nv_bfloat16* oh = s_os_hi + r8 * kOsRow + 2 * j;
    const __nv_bfloat16* ol = s_os_lo + r8 * kOsRow + 2 * j;
#pragma unroll 2
    for (int ks = 0; ks < kD / 16; ++ks) {
      const uint4* wp = a.wmfrag + (static_cast<size_t>(mt * (kD / 16) + ks) * 32 + lane) * 2;
      const uint4 wh = __ldg(wp), wl = __ldg(wp + 1);
      const uint32_t ahi[4] = {wh.x, wh.y, wh.z, wh.w};
      const uint32_t alo[4] = {wl.x, wl.y, wl.z, wl.w};
      const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(oh + ks * 16);
      const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(oh + ks * 16 + 8);
      const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(ol + ks * 16);
      const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(ol + ks * 16 + 8);
      mma_bf16(acc, ahi, bh0, bh1);
      mma_bf16(acc, alo, bh0, bh1);
      mma_bf16(acc, ahi, bl0, bl1);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int jj = mt * 16 + r8 + (e >> 1) * 8, pp = 2 * j + (e & 1);
      float v = acc[e] + __ldg(a.bm + jj);
      if (a.bnf != nullptr) {
        v = fmaxf(fmaf(v, __ldg(a.bnf + jj), __ldg(a.bnf + a.c_out + jj)), 0.f);
        if (a.mlp_keep != nullptr)
          v *= a.mlp_keep[(static_cast<size_t>(b) * a.c_out + jj) * a.hw + pix0 + pp] ? a.mlp_keep_scale : 0.f;
      }
      s_ys[jj * kPix + pp] = v;
    }
  }
  __syncthreads();

  if (a.bnf == nullptr) {  // training: BatchNorm statistics need every row of the batch first
    const size_t row0 = static_cast<size_t>(b) * a.hw + pix0;
    for (int i = tid; i < kPix * a.c_out; i += kThreads) {
      const int pp = i / a.c_out, jj = i - pp * a.c_out;
      a.ypre[(row0 + pp) * a.c_out + jj] = s_ys[jj * kPix + pp];
    }
    return;
  }

  // ---- phase 6: output GroupNorm over c_out / 16 channels per pixel                              tae.py:488
  const int cog = a.c_out / kH;
  for (int i = tid; i < kH * kPix; i += kThreads) {
    const int g = i / kPix, pp = i - g * kPix;
    float* yg = s_ys + g * cog * kPix + pp;
    float m = 0.f;
    for (int k = 0; k < cog; ++k) m += yg[k * kPix];
    m /= static_cast<float>(cog);
    float var = 0.f;
    for (int k = 0; k < cog; ++k) {
      const float dlt = yg[k * kPix] - m;
      var = fmaf(dlt, dlt, var);
    }
    const float rstd = 1.f / sqrtf(var / static_cast<float>(cog) + a.gn_eps);
    for (int k = 0; k < cog; ++k) {
      const int jj = g * cog + k;
      yg[k * kPix] = fmaf((yg[k * kPix] - m) * rstd, __ldg(a.on_w + jj), __ldg(a.on_b + jj));
    }
  }
  __syncthreads();
  __nv_bfloat16* ob = a.out + static_cast<size_t>(b) * a.c_out * a.hw + pix0;
  for (int jj = tid; jj < a.c_out; jj += kThreads) {  // one 16-byte store per (channel, tile)
    float f[8];
    const float4 y0 = *reinterpret_cast<const float4*>(s_ys + jj * kPix);
    const float4 y1 = *reinterpret_cast<const float4*>(s_ys + jj * kPix + 4);
    f[0] = y0.x, f[1] = y0.y, f[2] = y0.z, f[3] = y0.w, f[4] = y1.x, f[5] = y1.y, f[6] = y1.z, f[7] = y1.w;
    st_stream_v4(ob + static_cast<size_t>(jj) * a.hw, Elem<__nv_bfloat16>::pack(f));
  }
}

// ---- fragment-ordered weights (built per call on the device) ----------------------------------------------
// ufrag[ks][lane][8]: fp32 U (in_norm.weight folded) in A-fragment order, rows = heads
__global__ void build_ufrag_kernel(const float* __restrict__ u /*[C][16]*/, float* __restrict__ uf, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (C / 16) * 32 * 8) return;
  const int e = i & 7, lane = (i >> 3) & 31, ks = i >> 8;
  const int j = lane & 3, r8 = lane >> 2;
  const int row = r8 + ((e >> 1) & 1) * 8;
  const int k = 2 * j + (e & 1) + (e >> 2) * 8;
  uf[i] = u[(ks * 16 + k) * kMaxHeads + row];
}

// A-fragment order of a row-major weight W[rows][cols], tiled (rows/16) x (cols/16); hi words then lo words
__global__ void build_wfrag_kernel(const float* __restrict__ w, uint32_t* __restrict__ wf, int rows, int cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one (tile, lane, reg)
  const int n = (rows / 16) * (cols / 16) * 32 * 4;
  if (i >= n) return;
  const int reg = i & 3, lane = (i >> 2) & 31, tile = i >> 7;
  const int ks = tile % (cols / 16), mt = tile / (cols / 16);
  const int j = lane & 3, r8 = lane >> 2;
  const int row = mt * 16 + r8 + (reg & 1) * 8;
  const int k = ks * 16 + 2 * j + (reg >> 1) * 8;
  const float v0 = w[static_cast<size_t>(row) * cols + k], v1 = w[static_cast<size_t>(row) * cols + k + 1];
  uint32_t hi, lo;
  split2(v0, v1, hi, lo);
  wf[(static_cast<size_t>(tile) * 32 + lane) * 8 + reg] = hi;
  wf[(static_cast<size_t>(tile) * 32 + lane) * 8 + 4 + reg] = lo;
}

}  // namespace

bool ltae_mma_eligible(const c2s_ltae_desc& d, const void* x, const void* out) {
  const bool attn_only = (d.flags & C2S_LTAE_ATTN_ONLY) != 0;
  if (d.dtype != C2S_BF16 || d.n_head != kH || d.d_model != kD || !d.has_inconv) return false;
  if (d.C != 64 && d.C != 128) return false;
  if (d.T > kTP || (d.H * d.W) % kPix != 0) return false;
  if (d.pe_mode == C2S_PE_SINUSOID_LINEAR) return false;  // table differs per head chunk
  if (!attn_only && (d.c_out % 16 != 0 || d.c_out > 256)) return false;
  if (reinterpret_cast<uintptr_t>(x) % 16 != 0 || reinterpret_cast<uintptr_t>(out) % 16 != 0) return false;
  return true;
}

size_t ltae_mma_workspace_floats(const c2s_ltae_desc& d) {
  const size_t co = (d.flags & C2S_LTAE_ATTN_ONLY) ? 0 : d.c_out;
  return align64(static_cast<size_t>(d.C / 16) * 32 * 8) + align64(static_cast<size_t>(kD / 16) * (d.C / 16) * 32 * 8) +
         align64((co / 16) * static_cast<size_t>(kD / 16) * 32 * 8);
}

int ltae_mma_forward(const c2s_ltae_desc& d, const c2s_ltae_params& p, const void* x, const uint8_t* pad_mask,
                     void* out, float* attn, float* ws, const LtaeWorkspace& lay, float* frag_ws,
                     cudaStream_t stream) {
  // test hook: keep the MLP inside the attention kernel (mma.sync) instead of the tcgen05 row GEMM
  const bool split_mlp = !(d.flags & C2S_LTAE_ATTN_ONLY) && getenv("C2S_LTAE_NO_TCGEN05") == nullptr;
  const bool attn_only = (d.flags & C2S_LTAE_ATTN_ONLY) != 0;
  const bool train = (d.flags & C2S_LTAE_BN_BATCH_STATS) != 0 && !attn_only;
  const int C = d.C, KS = C / 16;
  float* ufrag = frag_ws;
  uint32_t* wcfrag = reinterpret_cast<uint32_t*>(frag_ws + align64(static_cast<size_t>(KS) * 32 * 8));
  uint32_t* wmfrag = wcfrag + align64(static_cast<size_t>(kD / 16) * KS * 32 * 8);
  build_ufrag_kernel<<<ceil_div(KS * 256, 256), 256, 0, stream>>>(ws + lay.u, ufrag, C);
  C2S_LAUNCH_CHECK("ltae_build_ufrag");
  if (!attn_only) {
    build_wfrag_kernel<<<ceil_div((kD / 16) * KS * 128, 256), 256, 0, stream>>>(p.inconv_weight, wcfrag, kD, C);
    C2S_LAUNCH_CHECK("ltae_build_wcfrag");
    if (!split_mlp) {
      build_wfrag_kernel<<<ceil_div((d.c_out / 16) * (kD / 16) * 128, 256), 256, 0, stream>>>(p.mlp_weight, wmfrag,
                                                                                           d.c_out, kD);
      C2S_LAUNCH_CHECK("ltae_build_wmfrag");
    }
  }
  MmaArgs a{};
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.pad = pad_mask;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.attn = attn;
  a.ufrag = ufrag;
  a.wcfrag = reinterpret_cast<const uint4*>(wcfrag);
  a.wmfrag = reinterpret_cast<const uint4*>(wmfrag);
  a.cpos = ws + lay.cpos;
  a.pe = d.pe_mode != C2S_PE_NONE ? ws + lay.pe : nullptr;
  a.bc = p.inconv_bias, a.bm = p.mlp_bias;
  a.gamma = p.in_norm_weight, a.beta = p.in_norm_bias;
  a.bnf = (attn_only || train) ? nullptr : ws + lay.bnf;
  a.on_w = p.out_norm_weight, a.on_b = p.out_norm_bias;
  a.ypre = train ? ws + lay.ypre : nullptr;
  a.attn_keep = p.attn_keep, a.mlp_keep = p.mlp_keep;
  a.save_o = p.save_o;
  a.attn_keep_scale = d.attn_keep_scale, a.mlp_keep_scale = d.mlp_keep_scale;
  a.B = d.B, a.T = d.T, a.hw = d.H * d.W, a.c_out = attn_only ? 0 : d.c_out;
  a.attn_only = attn_only;
  a.skip_attn_store = (d.flags & C2S_LTAE_SKIP_ATTN_STORE) != 0;
  a.zero_padded = (d.flags & C2S_LTAE_ZERO_PADDED) != 0;
  a.gn_eps = d.gn_eps;
  a.tiles_per_b = a.hw / kPix;
  if (split_mlp) {
    __nv_bfloat16 *w_hi, *w_lo;
    ltae_mlp_tc_buffers(d, ws + lay.tc, &a.o_hi, &a.o_lo, &w_hi, &w_lo);
  }
  const long long n_tiles = static_cast<long long>(d.B) * a.tiles_per_b;
  if (n_tiles > 0x7fffffffll) C2S_UNSUPPORTED("c2s_ltae_forward: too many pixel tiles");
  if (C == 128) {
    C2S_CUDA(cudaFuncSetAttribute(ltae_mma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128>::kTotal));
    ltae_mma_kernel<128><<<static_cast<unsigned>(n_tiles), kThreads, Smem<128>::kTotal, stream>>>(a);
    C2S_LAUNCH_CHECK("ltae_forward<mma,C=128>");
  } else {
    C2S_CUDA(cudaFuncSetAttribute(ltae_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<64>::kTotal));
    ltae_mma_kernel<64><<<static_cast<unsigned>(n_tiles), kThreads, Smem<64>::kTotal, stream>>>(a);
    C2S_LAUNCH_CHECK("ltae_forward<mma,C=64>");
  }
  if (split_mlp) return ltae_mlp_tc_forward(d, p, ws + lay.tc, a.bnf, a.ypre, out, stream);
  return C2S_OK;
}

}  // namespace c2s
