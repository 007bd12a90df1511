// L-TAE attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a: bf16 features, n_head = 16,
// d_model = 256, C in {64, 128}, T <= 61.  Same contract as ltae_mma_kernel (c2s_ltae_mma.cu); reference
// LTAE.forward / LTAE4WTAE.forward, src/backbones/tae.py:451-504, 589-635.
//
// One CTA = 8 consecutive pixels.  The slab x[b, t, :, pix0..pix0+7] is brought by TMA (one 3-D box per live frame)
// into shared memory in its NATURAL layout X[t][c][8 pixels] -- 16-byte rows, no transposition, no register staging.
// That layout is directly a canonical no-swizzle UMMA operand in both roles:
//   scores  S[(t,p), h]   = sum_c X[(t,p), c] U[h, c]      A = X as an MN-major operand (M = 16 frames x 8 pixels)
//   values  Z[(h,p), c]   = sum_(t,p') Ad[(h,p),(t,p')] X[(t,p'), c]   B = X as a K-major operand, Ad block-diagonal
// GroupNorm cannot be applied to x before the score product without rounding it, so the product is kept per
// normalisation group: group g accumulates in its own 16 TMEM columns (B = U restricted to the group's channels,
// bf16 hi + lo) and the epilogue combines S = sum_g rstd[g,p] D_g - mean term + cpos in fp32.
// The value product embeds the per-pixel attention a[h,t,p] as 8x8 diagonal core matrices: M = 16 heads x 8 pixels
// = 128 rows, K = (frame, pixel), 7/8 of the MACs are zeros -- tcgen05 has the throughput to spare (4 k cycles per
// tile) and nothing has to be transposed.
#include <cuda.h>

#include <cstdlib>

#include "c2s_ltae_prep.cuh"

namespace c2s {
namespace {

constexpr int kTcThreads = 128;
constexpr int kPixT = 8;
constexpr int kHeads = 16;
constexpr int kFrames = 62;   // slab frames (61 + one zero frame so that frames pair up)
constexpr float kMask = -1e6f;

struct TcArgs {
  const uint8_t* pad;
  float* attn;
  const uint16_t* ub;     // score weights, UMMA tiles [chunk][group in chunk][hi|lo][512 B]
  const float* ugs;       // [16 groups][16 heads]  sum_{c in g} U[h,c]
  const float* cpos;      // [B, T, 16]
  const float* pe;        // [B, T, 256] or nullptr
  const uint16_t* wct;    // in-projection weights, UMMA tiles [k-step][hi|lo][k_atom 2][n_atom 32][8][8]
  const float* bc;        // inconv.bias [256]
  const float* gamma;     // in_norm.weight [C]
  const float* beta;
  const uint8_t* attn_keep;
  float attn_keep_scale;
  __nv_bfloat16* o_hi;    // [B*hw][256] rows for the tcgen05 MLP kernel
  __nv_bfloat16* o_lo;
  int B, T, hw;
  int attn_only, skip_attn_store, zero_padded;
  float gn_eps;
  int tiles_per_b;
};

template <int C>
struct TcSmem {
  static constexpr int kFrameBytes = C * 16;                       // [C rows][8 pixels] bf16
  static constexpr int oSlab = 0;
  static constexpr int kSlab = kFrames * kFrameBytes;
  static constexpr int oUb = oSlab + kSlab;                        // finite data right behind the slab: the last
  static constexpr int kUb = (C / 16) * (16 / (C / 16)) * 2 * 512;  // m-tile of the scores reads two frames too far
  static constexpr int oS = oUb + kUb;                             // float [62][8][16] scores, then probabilities
  static constexpr int kS = kFrames * kPixT * kHeads * 4;
  static constexpr int oCpos = oS + kS;                            // float [62][16]
  static constexpr int oStat = oCpos + kFrames * kHeads * 4;       // rstd, mean*rstd [16][8], mh [8][16], ugs [16][16]
  static constexpr int oSa = oStat + (2 * kHeads * kPixT + kPixT * kHeads + 16 * kHeads) * 4;  // float [8][16] sum_t a
  static constexpr int oPe = oSa + kPixT * kHeads * 4;             // float [62][16] positional table (first head chunk)
  static constexpr int oAd = (oPe + kFrames * 16 * 4 + 127) & ~127;  // block-diagonal attention tiles: 4 steps x (hi, lo) x 4 KB
  static constexpr int kAd = 4 * 2 * 4096;
  static constexpr int kTotal = oAd + kAd;
  // after the value product the slab region is re-used: A_z (hi, lo) [128 rows][C] + a 3-stage ring of Wc k-step tiles
  static constexpr int kAz = 128 * C * 2;
  static constexpr int oAzHi = oSlab, oAzLo = oSlab + kAz, oWring = oSlab + 2 * kAz;
  static constexpr int kWstage = 2 * 8192;                         // [256 rows][16 k] bf16, hi + lo
  static constexpr int kWStages = (C >= 128) ? 3 : 2;              // the ring may run on into the (dead) score weights
  static_assert(oWring + kWStages * kWstage <= kSlab + kUb, "in-projection operands must fit in the slab + Ub region");
};

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// canonical no-swizzle UMMA operand: 8x(16 B) core matrices of 128 contiguous bytes; lbo / sbo in bytes
__device__ __forceinline__ uint64_t umma_desc_none(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc(int m, int n, bool a_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                                 // c_format = F32
  d |= 1u << 7;                                 // a_format = BF16
  d |= 1u << 10;                                // b_format = BF16
  d |= (a_mn_major ? 1u : 0u) << 15;            // a_major: 1 = MN-major
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int C>
__global__ void __launch_bounds__(kTcThreads, 1)
ltae_tc_kernel(const __grid_constant__ CUtensorMap map_x, const TcArgs a) {
  using S = TcSmem<C>;
  constexpr int CPG = C / kHeads;      // channels per GroupNorm group
  constexpr int GPC = 16 / CPG;        // groups per 16-channel chunk
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[10];  // slab, scores, values, in-projection full[3] / done[3], projection complete
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (s32(smem_raw) + 127u) & ~127u;
  unsigned char* smem = smem_raw + (smem0 - s32(smem_raw));
  float* s_S = reinterpret_cast<float*>(smem + S::oS);
  float* s_cpos = reinterpret_cast<float*>(smem + S::oCpos);
  float* s_rstd = reinterpret_cast<float*>(smem + S::oStat);
  float* s_mur = s_rstd + kHeads * kPixT;
  float* s_mh = s_mur + kHeads * kPixT;
  float* s_ugs = s_mh + kPixT * kHeads;
  float* s_sa = reinterpret_cast<float*>(smem + S::oSa);
  float* s_pe = reinterpret_cast<float*>(smem + S::oPe);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / a.tiles_per_b;
  const int pix0 = (blockIdx.x - b * a.tiles_per_b) * kPixT;

  // frame masks (every warp derives them)
  unsigned long long live_mask = 0, pad_mask = 0;
#pragma unroll
  for (int base = 0; base < 64; base += 32) {
    const int t = base + lane;
    const bool pd = t < a.T && a.pad != nullptr && __ldg(a.pad + b * a.T + t) != 0;
    const bool lv = t < a.T && !(pd && a.zero_padded);
    live_mask |= static_cast<unsigned long long>(__ballot_sync(0xffffffffu, lv)) << base;
    pad_mask |= static_cast<unsigned long long>(__ballot_sync(0xffffffffu, pd)) << base;
  }
  const int n_live = __popcll(live_mask);

  if (tid == 0) {
    for (int i = 0; i < 10; ++i) mbar_init(s32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  // ---- 1. slab + score weights by TMA / bulk copy; frames that are not read are zero-filled -------------
  if (tid == 0) {
    const uint32_t bar = s32(&bars[0]);
    mbar_expect_tx(bar, static_cast<uint32_t>(n_live) * S::kFrameBytes + S::kUb);
    for (int t = 0; t < a.T; ++t)
      if ((live_mask >> t) & 1ull) tma_load_3d(smem0 + S::oSlab + t * S::kFrameBytes, &map_x, pix0, 0, b * a.T + t, bar);
    bulk_g2s(smem0 + S::oUb, a.ub, S::kUb, bar);
  }
  for (int t = 0; t < kFrames; ++t) {
    if (!((live_mask >> t) & 1ull)) {
      uint4* f = reinterpret_cast<uint4*>(smem + S::oSlab + t * S::kFrameBytes);
      for (int i = tid; i < S::kFrameBytes / 16; i += kTcThreads) f[i] = make_uint4(0, 0, 0, 0);
    }
  }
  for (int i = tid; i < kFrames * kHeads; i += kTcThreads) {
    const int t = i / kHeads;
    s_cpos[i] = t < a.T ? __ldg(a.cpos + (static_cast<size_t>(b) * a.T + t) * kMaxHeads + (i - t * kHeads)) : 0.f;
  }
  for (int i = tid; i < 16 * kHeads; i += kTcThreads) s_ugs[i] = __ldg(a.ugs + i);
  if (!a.attn_only) {
    for (int i = tid; i < kFrames * 16; i += kTcThreads) {
      const int t = i >> 4;
      s_pe[i] = (a.pe != nullptr && t < a.T) ? __ldg(a.pe + (static_cast<size_t>(b) * a.T + t) * 256 + (i & 15)) : 0.f;
    }
    uint4* ad = reinterpret_cast<uint4*>(smem + S::oAd);  // off-diagonal elements stay zero for the whole kernel
    for (int i = tid; i < S::kAd / 16; i += kTcThreads) ad[i] = make_uint4(0, 0, 0, 0);
  }
  mbar_wait(s32(&bars[0]), 0);

  // ---- 2. GroupNorm statistics per (pixel, group) over all T frames (padded frames count as zeros) -- tae.py:461
  {
    const int p = tid & 7, g = tid >> 3;
    const __nv_bfloat16* xs = reinterpret_cast<const __nv_bfloat16*>(smem + S::oSlab) + (g * CPG) * 8 + p;
    float pivot = 0.f;
    if (n_live > 0) pivot = __bfloat162float(xs[(__ffsll(static_cast<long long>(live_mask)) - 1) * (C * 8)]);
    float s1 = 0.f, s2 = 0.f;
    for (int t = 0; t < a.T; ++t) {
      if (!((live_mask >> t) & 1ull)) continue;
#pragma unroll
      for (int cc = 0; cc < CPG; ++cc) {
        const float d = __bfloat162float(xs[t * (C * 8) + cc * 8]) - pivot;
        s1 += d;
        s2 = fmaf(d, d, s2);
      }
    }
    const float n_all = static_cast<float>(a.T) * CPG;
    const float n_skip = n_all - static_cast<float>(n_live) * CPG;
    s1 -= n_skip * pivot;
    s2 = fmaf(n_skip * pivot, pivot, s2);
    const float m = s1 / n_all;
    float var = s2 / n_all - m * m;
    var = var < 0.f ? 0.f : var;
    const float rstd = 1.f / sqrtf(var + a.gn_eps);
    s_rstd[g * kPixT + p] = rstd;
    s_mur[g * kPixT + p] = (m + pivot) * rstd;
  }
  __syncthreads();
  {  // mean term of the scores: mh[p][h] = sum_g mean*rstd[g,p] * sum_{c in g} U[h,c]
    const int p = tid >> 4, h = tid & 15;
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < kHeads; ++g) s = fmaf(s_mur[g * kPixT + p], s_ugs[g * kHeads + h], s);
    s_mh[p * kHeads + h] = s;
  }
  // the generic-proxy writes (zero frames) must be visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  // ---- 3. scores: four m-tiles of 16 frames x 8 pixels; group g accumulates in TMEM columns [16 g, 16 g + 16) ----
  const uint32_t idesc_s = umma_idesc(128, 16, /*a_mn_major=*/true);
  for (int mt = 0; mt < 4; ++mt) {
    const uint32_t acc = tmem + static_cast<uint32_t>((mt & 1) * 256);
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ch = 0; ch < C / 16; ++ch) {
        // A: X[(t,p), c] for 16 frames from 16 mt, 16 channels from 16 ch: MN-major, m-groups (frames) SBO apart,
        // k-groups (8 channels) LBO = 128 B apart
        const uint64_t adesc = umma_desc_none(smem0 + S::oSlab + (mt * 16) * S::kFrameBytes + ch * 256, 128, S::kFrameBytes);
#pragma unroll
        for (int gi = 0; gi < GPC; ++gi) {
          const int g = ch * GPC + gi;
          const uint32_t ub = smem0 + S::oUb + ((ch * GPC + gi) * 2) * 512;
          // B: [16 heads][16 channels] K-major: core (n_atom, k_atom) at n_atom * 128 + k_atom * 256
          umma(acc + g * 16, adesc, umma_desc_none(ub, 256, 128), idesc_s, 0);
          umma(acc + g * 16, adesc, umma_desc_none(ub + 512, 256, 128), idesc_s, 1);
        }
      }
      umma_commit(s32(&bars[1]));
    }
    mbar_wait(s32(&bars[1]), mt & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {  // epilogue: thread = TMEM lane = row (t_local, p)
      const int p = tid & 7, t = mt * 16 + (tid >> 3);
      float s[kHeads];
#pragma unroll
      for (int h = 0; h < kHeads; ++h) s[h] = 0.f;
#pragma unroll 4
      for (int g = 0; g < kHeads; ++g) {
        float d[16];
        tmem_ld16(acc + (static_cast<uint32_t>(warp * 32) << 16) + g * 16, d);
        const float r = s_rstd[g * kPixT + p];
#pragma unroll
        for (int h = 0; h < kHeads; ++h) s[h] = fmaf(r, d[h], s[h]);
      }
      if (t < kFrames) {
        const bool padded = (pad_mask >> t) & 1ull;
        float4* dst = reinterpret_cast<float4*>(s_S + (t * kPixT + p) * kHeads);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int h = 4 * q + e;
            v[e] = s[h] + s_cpos[t * kHeads + h] - s_mh[p * kHeads + h];
            if (padded) v[e] = kMask;           // tae.py:831
            if (t >= a.T) v[e] = -INFINITY;
          }
          dst[q] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }

  // ---- 4. softmax over t per (pixel, head) ------------------------------------------------------- tae.py:836
  {
    const int p = tid >> 4, h = tid & 15;
    float* col = s_S + p * kHeads + h;
    constexpr int stride = kPixT * kHeads;
    float mx = -INFINITY;
    for (int t = 0; t < a.T; ++t) mx = fmaxf(mx, col[t * stride]);
    float den = 0.f;
    for (int t = 0; t < a.T; ++t) {
      const float e = expf(col[t * stride] - mx);
      col[t * stride] = e;
      den += e;
    }
    const float inv = 1.f / den;
    float sa = 0.f;
    for (int t = 0; t < a.T; ++t) {
      float v = col[t * stride] * inv;
      if (a.attn_keep != nullptr)  // training: dropout acts on the attention that is returned (tae.py:837)
        v = a.attn_keep[((static_cast<size_t>(h) * a.B + b) * a.T + t) * a.hw + pix0 + p] ? v * a.attn_keep_scale : 0.f;
      col[t * stride] = v;
      sa += v;
    }
    for (int t = a.T; t < kFrames; ++t) col[t * stride] = 0.f;
    s_sa[p * kHeads + h] = sa;
  }
  __syncthreads();
  if (a.attn != nullptr && !a.skip_attn_store) {  // attn[h, b, t, pix0 .. pix0+7]                 tae.py:490-493
    const int pp = lane & 7, tq = lane >> 3;
    for (int h = warp; h < kHeads; h += kTcThreads / 32) {
      float* dst = a.attn + ((static_cast<size_t>(h) * a.B + b) * a.T + tq) * a.hw + pix0 + pp;
      const size_t step = static_cast<size_t>(4) * a.hw;
      for (int t = tq; t < a.T; t += 4, dst += step) *dst = s_S[(t * kPixT + pp) * kHeads + h];
    }
  }

  if (!a.attn_only) {
    const int h = tid >> 3, p = tid & 7;  // thread = TMEM lane = row (head, pixel) of the value / projection tiles
    // ---- 5. values: Z[(h,p), c] = sum_t a[h,t,p] x[t,c,p] as a block-diagonal tcgen05 product -------- tae.py:839
    // A tile of one step (frames 2s, 2s+1): [16 head atoms][2 frame atoms] core matrices diag(a[h, t, 0..7])
    const uint32_t idesc_z = umma_idesc(128, C, /*a_mn_major=*/false);
    constexpr int kSteps = kFrames / 2;
    for (int r0 = 0, round = 0; r0 < kSteps; r0 += 4, ++round) {
      const int n_steps = min(4, kSteps - r0);
      for (int js = 0; js < n_steps; ++js) {
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
          const int t = 2 * (r0 + js) + tl;
          const float v = s_S[(t * kPixT + p) * kHeads + h];
          const __nv_bfloat16 hi = __float2bfloat16_rn(v);
          const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
          unsigned char* tile = smem + S::oAd + js * 8192 + (tl * 16 + h) * 128 + p * 16 + p * 2;
          *reinterpret_cast<__nv_bfloat16*>(tile) = hi;
          *reinterpret_cast<__nv_bfloat16*>(tile + 4096) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int js = 0; js < n_steps; ++js) {
          const int s_ = r0 + js;
          // B: X[(t,p'), c] for frames 2s, 2s+1: K-major, channel atoms SBO = 128 B apart, frame atoms LBO = one frame apart
          const uint64_t bdesc = umma_desc_none(smem0 + S::oSlab + (2 * s_) * S::kFrameBytes, S::kFrameBytes, 128);
          const uint32_t at = smem0 + S::oAd + js * 8192;
          umma(tmem, umma_desc_none(at, 2048, 128), bdesc, idesc_z, s_ != 0);
          umma(tmem, umma_desc_none(at + 4096, 2048, 128), bdesc, idesc_z, 1);
        }
        umma_commit(s32(&bars[2]));
      }
      mbar_wait(s32(&bars[2]), round & 1);  // the tiles may be rewritten (and, after the last round, the slab re-used)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    // ---- 6. GroupNorm affine on the weighted sums -> A operand of the in-projection (bf16 hi + lo) --- tae.py:461
    const float sa = s_sa[p * kHeads + h];
    {
      unsigned char* az_hi = smem + S::oAzHi + h * (C / 8) * 128 + p * 16;
      unsigned char* az_lo = smem + S::oAzLo + h * (C / 8) * 128 + p * 16;
      for (int c0 = 0; c0 < C; c0 += 16) {
        float z[16];
        tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, z);
        uint32_t whi[8], wlo[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float zn[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = c0 + e + q, g = c / CPG;
            // sum_t a (x rstd - mean rstd) gamma + beta sum_t a
            zn[q] = fmaf(__ldg(a.gamma + c), fmaf(z[e + q], s_rstd[g * kPixT + p], -s_mur[g * kPixT + p] * sa),
                         __ldg(a.beta + c) * sa);
          }
          const __nv_bfloat162 hi2 = __floats2bfloat162_rn(zn[0], zn[1]);
          const __nv_bfloat162 lo2 = __floats2bfloat162_rn(zn[0] - __low2float(hi2), zn[1] - __high2float(hi2));
          whi[e / 2] = *reinterpret_cast<const uint32_t*>(&hi2);
          wlo[e / 2] = *reinterpret_cast<const uint32_t*>(&lo2);
        }
        // k-atoms (8 channels) are 128 B apart, head atoms (C/8) * 128 B apart; this thread owns row p of its head atom
        *reinterpret_cast<uint4*>(az_hi + (c0 / 8) * 128) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
        *reinterpret_cast<uint4*>(az_hi + (c0 / 8 + 1) * 128) = make_uint4(whi[4], whi[5], whi[6], whi[7]);
        *reinterpret_cast<uint4*>(az_lo + (c0 / 8) * 128) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
        *reinterpret_cast<uint4*>(az_lo + (c0 / 8 + 1) * 128) = make_uint4(wlo[4], wlo[5], wlo[6], wlo[7]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    // ---- 7. in-projection O[(h,p), d] = Zn[(h,p), :] . Wc[d, :]  (N = 256; row block h keeps columns 16h..16h+15) ----
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc_o = umma_idesc(128, 256, /*a_mn_major=*/false);
      constexpr int KS = C / 16;
      constexpr int NST = S::kWStages;
      auto load = [&](int ks) {
        const int st = ks % NST;
        const uint32_t full = s32(&bars[3 + st]);
        mbar_expect_tx(full, S::kWstage);
        bulk_g2s(smem0 + S::oWring + st * S::kWstage, a.wct + static_cast<size_t>(ks) * (S::kWstage / 2), S::kWstage, full);
      };
      for (int i = 0; i < NST - 1 && i < KS; ++i) load(i);
      for (int ks = 0; ks < KS; ++ks) {
        const int st = ks % NST, nxt = ks + NST - 1;
        if (nxt < KS) {
          if (nxt >= NST) mbar_wait(s32(&bars[6 + nxt % NST]), ((nxt / NST) - 1) & 1);  // its previous products are done
          load(nxt);
        }
        mbar_wait(s32(&bars[3 + st]), (ks / NST) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t a_hi = umma_desc_none(smem0 + S::oAzHi + ks * 256, 128, (C / 8) * 128);
        const uint64_t a_lo = umma_desc_none(smem0 + S::oAzLo + ks * 256, 128, (C / 8) * 128);
        const uint32_t wt = smem0 + S::oWring + st * S::kWstage;
        const uint64_t b_hi = umma_desc_none(wt, 4096, 128), b_lo = umma_desc_none(wt + 8192, 4096, 128);
        umma(tmem + 256, a_hi, b_hi, idesc_o, ks != 0);
        umma(tmem + 256, a_lo, b_hi, idesc_o, 1);
        umma(tmem + 256, a_hi, b_lo, idesc_o, 1);
        umma_commit(s32(&bars[6 + st]));
      }
      umma_commit(s32(&bars[9]));  // single-use barrier: the ring barriers complete several phases and cannot be
    }                              // waited on by threads that did not follow them
    // positional term while the tensor core works: pa[i] = sum_t a[h,t,p] PE[b,t,16h+i] (the table is tiled per head)
    float pa[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pa[i] = 0.f;
    if (a.pe != nullptr) {
      for (int t = 0; t < a.T; ++t) {
        const float av = s_S[(t * kPixT + p) * kHeads + h];
        const float4* pe = reinterpret_cast<const float4*>(s_pe + t * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 e = pe[q];
          pa[4 * q] = fmaf(av, e.x, pa[4 * q]), pa[4 * q + 1] = fmaf(av, e.y, pa[4 * q + 1]);
          pa[4 * q + 2] = fmaf(av, e.z, pa[4 * q + 2]), pa[4 * q + 3] = fmaf(av, e.w, pa[4 * q + 3]);
        }
      }
    }
    mbar_wait(s32(&bars[9]), 0);  // covers every product of the chain
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- 8. o[p, 16h + i] = O + sum_t a * bc + positional term -> bf16 hi/lo rows for the MLP kernel ---- tae.py:463,479
    {
      float o[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // the column offset of tcgen05.ld is warp-uniform: load the warp's 4 heads, keep one
        float v[16];
        tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 256 + (warp * 4 + q) * 16, v);
        if (q == (h & 3)) {
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = v[i];
        }
      }
      uint32_t whi[8], wlo[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float v0 = o[i] + sa * __ldg(a.bc + h * 16 + i) + pa[i];
        const float v1 = o[i + 1] + sa * __ldg(a.bc + h * 16 + i + 1) + pa[i + 1];
        const __nv_bfloat162 hi2 = __floats2bfloat162_rn(v0, v1);
        const __nv_bfloat162 lo2 = __floats2bfloat162_rn(v0 - __low2float(hi2), v1 - __high2float(hi2));
        whi[i / 2] = *reinterpret_cast<const uint32_t*>(&hi2);
        wlo[i / 2] = *reinterpret_cast<const uint32_t*>(&lo2);
      }
      const size_t row = static_cast<size_t>(b) * a.hw + pix0 + p;
      uint4* dh = reinterpret_cast<uint4*>(a.o_hi + row * 256 + h * 16);
      uint4* dl = reinterpret_cast<uint4*>(a.o_lo + row * 256 + h * 16);
      dh[0] = make_uint4(whi[0], whi[1], whi[2], whi[3]), dh[1] = make_uint4(whi[4], whi[5], whi[6], whi[7]);
      dl[0] = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]), dl[1] = make_uint4(wlo[4], wlo[5], wlo[6], wlo[7]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// score weights as UMMA B tiles: [chunk][group in chunk][hi|lo][k_atom 2][n_atom 2][8 heads][8 channels] bf16,
// the channels outside the group are zero; ugs[g][h] = sum_{c in g} U[h,c]
__global__ void build_tc_scores_kernel(const float* __restrict__ u /*[C][16]*/, uint16_t* __restrict__ ub,
                                       float* __restrict__ ugs, int C) {
  const int cpg = C / 16, gpc = 16 / cpg;
  const int n_tiles = (C / 16) * gpc;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one element of one hi tile
  if (i < n_tiles * 256) {
    const int tile = i >> 8, e = i & 255;
    const int ka = e >> 7, na = (e >> 6) & 1, r = (e >> 3) & 7, kk = e & 7;
    const int ch = tile / gpc, gi = tile - ch * gpc;
    const int c_local = ka * 8 + kk, c = ch * 16 + c_local, h = na * 8 + r;
    const bool in_group = (c_local / cpg) == gi;
    const float v = in_group ? u[c * kMaxHeads + h] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    ub[(tile * 2) * 256 + e] = *reinterpret_cast<const uint16_t*>(&hi);
    ub[(tile * 2 + 1) * 256 + e] = *reinterpret_cast<const uint16_t*>(&lo);
  }
  if (i < 256) {
    const int g = i >> 4, h = i & 15;
    float s = 0.f;
    for (int cc = 0; cc < cpg; ++cc) s += u[(g * cpg + cc) * kMaxHeads + h];
    ugs[i] = s;
  }
}

// in-projection weights as UMMA B tiles: [k-step][hi|lo][k_atom 2][n_atom 32][8 rows d][8 channels] bf16
__global__ void build_tc_wc_kernel(const float* __restrict__ wc /*[256][C]*/, uint16_t* __restrict__ out, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one element of one hi tile
  if (i >= (C / 16) * 4096) return;
  const int ks = i >> 12, e = i & 4095;
  const int ka = e >> 11, na = (e >> 6) & 31, r = (e >> 3) & 7, kk = e & 7;
  const float v = wc[static_cast<size_t>(na * 8 + r) * C + ks * 16 + ka * 8 + kk];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  out[static_cast<size_t>(ks) * 8192 + e] = *reinterpret_cast<const uint16_t*>(&hi);
  out[static_cast<size_t>(ks) * 8192 + 4096 + e] = *reinterpret_cast<const uint16_t*>(&lo);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

bool ltae_tc_enabled() { return getenv("C2S_LTAE_TC") != nullptr; }

bool ltae_tc_eligible(const c2s_ltae_desc& d) { return d.T <= 61; }

size_t ltae_tc_workspace_floats(const c2s_ltae_desc& d) {
  // score tiles, group sums, in-projection tiles ([C/16][hi|lo][4096] bf16)
  return align64(static_cast<size_t>(TcSmem<128>::kUb) / 4) + align64(256) + align64(static_cast<size_t>(d.C / 16) * 8192 / 2);
}

int ltae_tc_forward(const c2s_ltae_desc& d, const c2s_ltae_params& p, const void* x, const uint8_t* pad_mask, void* out,
                    float* attn, float* ws, const LtaeWorkspace& lay, float* tc_ws, cudaStream_t stream) {
  const int C = d.C, hw = d.H * d.W;
  const bool attn_only = (d.flags & C2S_LTAE_ATTN_ONLY) != 0;
  const bool train = (d.flags & C2S_LTAE_BN_BATCH_STATS) != 0 && !attn_only;
  uint16_t* ub = reinterpret_cast<uint16_t*>(tc_ws);
  float* ugs = tc_ws + align64(static_cast<size_t>(TcSmem<128>::kUb) / 4);
  uint16_t* wct = reinterpret_cast<uint16_t*>(ugs + align64(256));
  if (!attn_only) {
    build_tc_wc_kernel<<<ceil_div((C / 16) * 4096, 256), 256, 0, stream>>>(p.inconv_weight, wct, C);
    C2S_LAUNCH_CHECK("ltae_build_tc_wc");
  }
  const int n_tiles_b = (C / 16) * (16 / (C / 16));
  build_tc_scores_kernel<<<ceil_div(n_tiles_b * 256, 256), 256, 0, stream>>>(ws + lay.u, ub, ugs, C);
  C2S_LAUNCH_CHECK("ltae_build_tc_scores");

  EncodeTiledFn fn = tc_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return C2S_ERR_CUDA;
  }
  CUtensorMap map;
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(hw), static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(d.B) * d.T};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(hw) * 2, static_cast<cuuint64_t>(C) * hw * 2};
  const cuuint32_t box[3] = {8, static_cast<cuuint32_t>(C), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (features) failed with CUresult %d", static_cast<int>(r));
    return C2S_ERR_CUDA;
  }
  TcArgs a{};
  a.pad = pad_mask, a.attn = attn, a.ub = ub, a.ugs = ugs, a.cpos = ws + lay.cpos;
  a.pe = d.pe_mode != C2S_PE_NONE ? ws + lay.pe : nullptr;
  a.wct = wct, a.bc = p.inconv_bias, a.gamma = p.in_norm_weight, a.beta = p.in_norm_bias;
  a.attn_keep = p.attn_keep, a.attn_keep_scale = d.attn_keep_scale;
  if (!attn_only) {
    __nv_bfloat16 *w_hi, *w_lo;
    ltae_mlp_tc_buffers(d, ws + lay.tc, &a.o_hi, &a.o_lo, &w_hi, &w_lo);
  }
  a.B = d.B, a.T = d.T, a.hw = hw;
  a.attn_only = attn_only;
  a.skip_attn_store = (d.flags & C2S_LTAE_SKIP_ATTN_STORE) != 0;
  a.zero_padded = (d.flags & C2S_LTAE_ZERO_PADDED) != 0;
  a.gn_eps = d.gn_eps;
  a.tiles_per_b = hw / kPixT;
  const long long n_tiles = static_cast<long long>(d.B) * a.tiles_per_b;
  if (C == 128) {
    const int smem = TcSmem<128>::kTotal + 256;
    C2S_CUDA(cudaFuncSetAttribute(ltae_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ltae_tc_kernel<128><<<static_cast<unsigned>(n_tiles), kTcThreads, smem, stream>>>(map, a);
    C2S_LAUNCH_CHECK("ltae_attention<tcgen05,C=128>");
  } else {
    const int smem = TcSmem<64>::kTotal + 256;
    C2S_CUDA(cudaFuncSetAttribute(ltae_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ltae_tc_kernel<64><<<static_cast<unsigned>(n_tiles), kTcThreads, smem, stream>>>(map, a);
    C2S_LAUNCH_CHECK("ltae_attention<tcgen05,C=64>");
  }
  if (!attn_only)
    return ltae_mlp_tc_forward(d, p, ws + lay.tc, train ? nullptr : ws + lay.bnf, train ? ws + lay.ypre : nullptr, out,
                               stream);
  return C2S_OK;
}

}  // namespace c2s
