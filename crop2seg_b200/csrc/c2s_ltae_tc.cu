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
  static constexpr int kTotal = oStat + (2 * kHeads * kPixT + kPixT * kHeads + 16 * kHeads) * 4;
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
  __shared__ __align__(8) unsigned long long bars[3];  // slab landed, score tile done, (stage B) value tile done
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (s32(smem_raw) + 127u) & ~127u;
  unsigned char* smem = smem_raw + (smem0 - s32(smem_raw));
  float* s_S = reinterpret_cast<float*>(smem + S::oS);
  float* s_cpos = reinterpret_cast<float*>(smem + S::oCpos);
  float* s_rstd = reinterpret_cast<float*>(smem + S::oStat);
  float* s_mur = s_rstd + kHeads * kPixT;
  float* s_mh = s_mur + kHeads * kPixT;
  float* s_ugs = s_mh + kPixT * kHeads;

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
    for (int i = 0; i < 3; ++i) mbar_init(s32(&bars[i]), 1);
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
    for (int t = 0; t < a.T; ++t) col[t * stride] *= inv;
    for (int t = a.T; t < kFrames; ++t) col[t * stride] = 0.f;
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

bool ltae_tc_eligible(const c2s_ltae_desc& d) {
  return (d.flags & C2S_LTAE_ATTN_ONLY) != 0 && d.T <= 61 && (d.flags & C2S_LTAE_SKIP_ATTN_STORE) == 0;
}

size_t ltae_tc_workspace_floats(const c2s_ltae_desc& d) {
  return align64(static_cast<size_t>(TcSmem<128>::kUb) / 4) + align64(256);
}

int ltae_tc_forward(const c2s_ltae_desc& d, const c2s_ltae_params& p, const void* x, const uint8_t* pad_mask, float* attn,
                    float* ws, const LtaeWorkspace& lay, float* tc_ws, cudaStream_t stream) {
  const int C = d.C, hw = d.H * d.W;
  uint16_t* ub = reinterpret_cast<uint16_t*>(tc_ws);
  float* ugs = tc_ws + align64(static_cast<size_t>(TcSmem<128>::kUb) / 4);
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
  a.B = d.B, a.T = d.T, a.hw = hw;
  a.attn_only = 1;
  a.skip_attn_store = 0;
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
  return C2S_OK;
}

}  // namespace c2s
