// Fused L-TAE forward for sm_100a: LTAE.forward / LTAE4WTAE.forward (reference src/backbones/tae.py:451-504,
// 589-635) with LightweightMultiHeadAttention (tae.py:760-807) and ScaledDotProductAttention (tae.py:822-847).
//
// One CTA owns a tile of kPT consecutive pixels of one sample and carries the whole pipeline for
// those pixel rows on chip: GroupNorm statistics over (C/h x T) -> per-head scores (collapsed
// in-projection + key projection + query dot, see c2s_ltae_prep.cuh) -> pad-masked softmax over T
// -> attention-weighted temporal sum of the normalised features -> per-head in-projection of the
// sum -> MLP + BatchNorm + ReLU -> output GroupNorm.  The reference's [N,T,D] activations, the
// head-major copies of Q / mask / V and the per-pixel positional tables never exist.
//
// This file holds the general kernel (any C, T, n_head <= 16, fp32 math on the CUDA cores, fp32 or
// bf16 I/O).  x is swept three times by the same CTA (statistics, scores, weighted sum); the second
// and third sweep hit L2 because a tile's slab is a few hundred KB.
#include <cstdlib>
#include <cstring>

#include <mutex>

#include "c2s_ltae_prep.cuh"

namespace c2s {
namespace {

constexpr int kPT = 8;             // pixels per CTA tile (one 16 B bf16 / 32 B fp32 segment per row)
constexpr int kLtaeThreads = 256;
constexpr int kHP = kMaxHeads + 4;  // padded head stride of the score tile (bank-conflict free LDS.128)
constexpr float kMaskFill = -1e6f;  // tae.py:831

struct LtaeArgs {
  const void* x;
  const uint8_t* pad;
  void* out;
  float* attn;
  const float* u;     // [C, kMaxHeads]
  const float* cpos;  // [B, T, kMaxHeads]
  const float* wct;   // [C, D]
  const float* bc;    // [D]
  const float* wmt;   // [D, c_out]
  const float* bm;    // [c_out]
  const float* pe;    // [B, T, D] or nullptr
  const float* gamma; // in_norm.weight [C]
  const float* beta;  // in_norm.bias [C]
  const float* bnf;   // [2, c_out] folded eval BatchNorm (nullptr: leave pre-BN output in ypre)
  const float* on_w;  // out_norm.weight
  const float* on_b;
  float* ypre;        // [N, c_out] pre-BatchNorm rows (training mode)
  float* save_o;      // [N, D] rows entering the MLP (saved for the backward) or nullptr
  const uint8_t* attn_keep;  // [h, B, T, hw] dropout keep mask or nullptr
  const uint8_t* mlp_keep;   // [B, c_out, hw] or nullptr
  float attn_keep_scale, mlp_keep_scale;
  int B, T, C, hw;
  int n_head, cpg, D, dh, c_out, cog;
  int has_inconv, attn_only, skip_attn_store, zero_padded;
  float gn_eps;
  int tiles_per_b;
};

struct LtaeSmem {
  // offsets in floats
  int mu, rstd, sa, u, sc, zs, os, ys, frames, total;
  int zs_head_stride;
};

__host__ __device__ inline LtaeSmem ltae_smem(int T, int C, int D, int c_out, int n_head, bool attn_only) {
  LtaeSmem s{};
  int off = 0;
  auto take = [&](int n) {
    int o = off;
    off += (n + 3) & ~3;
    return o;
  };
  s.mu = take(kMaxHeads * kPT);
  s.rstd = take(kMaxHeads * kPT);
  s.sa = take(kPT * kHP);
  s.u = take(C * kMaxHeads);
  s.sc = take(T * kPT * kHP);
  s.zs_head_stride = C * kPT + 8;
  s.zs = take(attn_only ? 0 : n_head * s.zs_head_stride);
  s.os = take(attn_only ? 0 : D * kPT);
  s.ys = take(attn_only ? 0 : c_out * kPT);
  s.frames = take((T + 1) / 2 + (T + 3) / 4 + 4);  // short[T] frame list + uint8[T] pad flags
  s.total = off;
  return s;
}

template <typename T>
__global__ void __launch_bounds__(kLtaeThreads) ltae_forward_kernel(const LtaeArgs a) {
  extern __shared__ __align__(16) float smem[];
  const LtaeSmem L = ltae_smem(a.T, a.C, a.D, a.c_out, a.n_head, a.attn_only != 0);
  float* s_mu = smem + L.mu;      // [g][p]  mean * rstd
  float* s_rstd = smem + L.rstd;  // [g][p]
  float* s_sa = smem + L.sa;      // [p][kHP]  sum_t attn
  float* s_u = smem + L.u;        // [c][kMaxHeads]
  float* s_sc = smem + L.sc;      // [t][p][kHP]  scores, then attention
  float* s_zs = smem + L.zs;      // [h][c][p] (+8 per head)
  float* s_os = smem + L.os;      // [d][p]
  float* s_ys = smem + L.ys;      // [j][p]
  short* s_frames = reinterpret_cast<short*>(smem + L.frames);
  uint8_t* s_pad = reinterpret_cast<uint8_t*>(s_frames + ((a.T + 1) & ~1));
  __shared__ int s_nframes;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / a.tiles_per_b;
  const int pix0 = (blockIdx.x - b * a.tiles_per_b) * kPT;
  const int n_pix = min(kPT, a.hw - pix0);
  const size_t frame_stride = static_cast<size_t>(a.C) * a.hw;
  const T* xb = static_cast<const T*>(a.x) + static_cast<size_t>(b) * a.T * frame_stride + pix0;

  // ---- phase 0: frame list, pad flags, folded score weights -------------------------------------------
  if (warp == 0) {
    int count = 0;
    for (int base = 0; base < a.T; base += 32) {
      const int t = base + lane;
      const bool padded = t < a.T && a.pad != nullptr && a.pad[b * a.T + t] != 0;
      if (t < a.T) s_pad[t] = padded ? 1 : 0;
      const bool keep = t < a.T && !(padded && a.zero_padded);  // zero frames need not be read
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) s_frames[count + __popc(m & ((1u << lane) - 1u))] = static_cast<short>(t);
      count += __popc(m);
    }
    if (lane == 0) s_nframes = count;
  }
  for (int i = tid; i < a.C * kMaxHeads; i += kLtaeThreads) s_u[i] = a.u[i];
  __syncthreads();
  const int n_frames = s_nframes;

  // ---- phase 1: GroupNorm statistics over (C/h channels x all T frames) per pixel -------- tae.py:461
  {
    const int p = lane % kPT, sub = lane / kPT;
    constexpr int kSub = 32 / kPT;
    const bool live = p < n_pix;
    const int n_read = n_frames * a.cpg;
    const float n_all = static_cast<float>(a.T) * a.cpg;
    for (int g = warp; g < a.n_head; g += kLtaeThreads / 32) {
      const T* xg = xb + static_cast<size_t>(g) * a.cpg * a.hw + p;
      // shifted sums: the pivot removes the cancellation of E[x^2] - E[x]^2
      const float pivot = (live && n_frames > 0) ? Elem<T>::load(xg + static_cast<size_t>(s_frames[0]) * frame_stride) : 0.f;
      float s1 = 0.f, s2 = 0.f;
      if (live) {
#pragma unroll 4
        for (int e = sub; e < n_read; e += kSub) {
          const int fi = e / a.cpg, cc = e - fi * a.cpg;
          const float v = Elem<T>::load(xg + static_cast<size_t>(s_frames[fi]) * frame_stride + static_cast<size_t>(cc) * a.hw) - pivot;
          s1 += v;
          s2 = fmaf(v, v, s2);
        }
      }
#pragma unroll
      for (int o = kPT; o < 32; o <<= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (sub == 0) {
        const float n_skip = n_all - static_cast<float>(n_read);  // frames known to be zero
        s1 -= n_skip * pivot;
        s2 = fmaf(n_skip * pivot, pivot, s2);
        const float m = s1 / n_all;
        float var = s2 / n_all - m * m;
        var = var < 0.f ? 0.f : var;
        const float rstd = 1.f / sqrtf(var + a.gn_eps);
        s_rstd[g * kPT + p] = rstd;
        s_mu[g * kPT + p] = (m + pivot) * rstd;
      }
    }
  }
  __syncthreads();

  // ---- phase 2: scores s[h,t] = U[h,:] . xn[t,:] + cpos[b,h,t]; padded frames get -1e6 --- tae.py:827-831
  for (int item = tid; item < a.T * kPT; item += kLtaeThreads) {
    const int t = item / kPT, p = item - t * kPT;
    float acc[kMaxHeads];
    if (s_pad[t] || p >= n_pix) {
#pragma unroll
      for (int k = 0; k < kMaxHeads; ++k) acc[k] = kMaskFill;
    } else {
      const float4* cp = reinterpret_cast<const float4*>(a.cpos + (static_cast<size_t>(b) * a.T + t) * kMaxHeads);
#pragma unroll
      for (int k4 = 0; k4 < kMaxHeads / 4; ++k4) {
        const float4 c = __ldg(cp + k4);
        acc[4 * k4] = c.x, acc[4 * k4 + 1] = c.y, acc[4 * k4 + 2] = c.z, acc[4 * k4 + 3] = c.w;
      }
      const T* xt = xb + static_cast<size_t>(t) * frame_stride + p;
      for (int g = 0; g < a.n_head; ++g) {
        const float r = s_rstd[g * kPT + p], m = s_mu[g * kPT + p];
#pragma unroll 4
        for (int cc = 0; cc < a.cpg; ++cc) {
          const int c = g * a.cpg + cc;
          const float xn = fmaf(Elem<T>::load(xt + static_cast<size_t>(c) * a.hw), r, -m);
          const float4* up = reinterpret_cast<const float4*>(s_u + c * kMaxHeads);
#pragma unroll
          for (int k4 = 0; k4 < kMaxHeads / 4; ++k4) {
            const float4 w = up[k4];
            acc[4 * k4] = fmaf(w.x, xn, acc[4 * k4]);
            acc[4 * k4 + 1] = fmaf(w.y, xn, acc[4 * k4 + 1]);
            acc[4 * k4 + 2] = fmaf(w.z, xn, acc[4 * k4 + 2]);
            acc[4 * k4 + 3] = fmaf(w.w, xn, acc[4 * k4 + 3]);
          }
        }
      }
    }
    float4* sp = reinterpret_cast<float4*>(s_sc + (t * kPT + p) * kHP);
#pragma unroll
    for (int k4 = 0; k4 < kMaxHeads / 4; ++k4)
      sp[k4] = make_float4(acc[4 * k4], acc[4 * k4 + 1], acc[4 * k4 + 2], acc[4 * k4 + 3]);
  }
  __syncthreads();

  // ---- phase 3: softmax over T per (pixel, head) ------------------------------------------ tae.py:836
  for (int item = tid; item < kPT * kMaxHeads; item += kLtaeThreads) {
    const int p = item / kMaxHeads, hh = item - p * kMaxHeads;
    float* col = s_sc + p * kHP + hh;
    const int stride = kPT * kHP;
    float sum_a = 0.f;
    if (hh < a.n_head) {
      float mx = -INFINITY;
      for (int t = 0; t < a.T; ++t) mx = fmaxf(mx, col[t * stride]);
      float den = 0.f;
      for (int t = 0; t < a.T; ++t) {
        const float e = expf(col[t * stride] - mx);
        col[t * stride] = e;
        den += e;
      }
      for (int t = 0; t < a.T; ++t) {
        float v = col[t * stride] / den;
        if (a.attn_keep != nullptr && p < n_pix)  // dropout acts on the attention that is returned (tae.py:837)
          v *= a.attn_keep[((static_cast<size_t>(hh) * a.B + b) * a.T + t) * a.hw + pix0 + p] ? a.attn_keep_scale : 0.f;
        col[t * stride] = v;
        sum_a += v;
      }
    } else {
      for (int t = 0; t < a.T; ++t) col[t * stride] = 0.f;
    }
    s_sa[p * kHP + hh] = sum_a;
  }
  __syncthreads();

  // attention maps: attn[h, b, t, y, x]                                              tae.py:490-493
  if (a.attn != nullptr && !a.skip_attn_store) {
    const int per_head = a.T * kPT;
    for (int item = tid; item < a.n_head * per_head; item += kLtaeThreads) {
      const int hh = item / per_head, r = item - hh * per_head;
      const int t = r / kPT, p = r - t * kPT;
      if (p < n_pix)
        a.attn[((static_cast<size_t>(hh) * a.B + b) * a.T + t) * a.hw + pix0 + p] = s_sc[(t * kPT + p) * kHP + hh];
    }
  }
  if (a.attn_only) return;

  // ---- phase 4: z[h,c] = sum_t a[h,t] xn[t,c], accumulated on raw x and normalised once --- tae.py:839
  for (int item = tid; item < a.C * kPT; item += kLtaeThreads) {
    const int c = item / kPT, p = item - c * kPT;
    float acc[kMaxHeads];
#pragma unroll
    for (int k = 0; k < kMaxHeads; ++k) acc[k] = 0.f;
    if (p < n_pix) {
      const T* xc = xb + static_cast<size_t>(c) * a.hw + p;
#pragma unroll 2
      for (int fi = 0; fi < n_frames; ++fi) {
        const int t = s_frames[fi];
        const float xv = Elem<T>::load(xc + static_cast<size_t>(t) * frame_stride);
        const float4* ap = reinterpret_cast<const float4*>(s_sc + (t * kPT + p) * kHP);
#pragma unroll
        for (int k4 = 0; k4 < kMaxHeads / 4; ++k4) {
          const float4 w = ap[k4];
          acc[4 * k4] = fmaf(w.x, xv, acc[4 * k4]);
          acc[4 * k4 + 1] = fmaf(w.y, xv, acc[4 * k4 + 1]);
          acc[4 * k4 + 2] = fmaf(w.z, xv, acc[4 * k4 + 2]);
          acc[4 * k4 + 3] = fmaf(w.w, xv, acc[4 * k4 + 3]);
        }
      }
    }
    const int g = c / a.cpg;
    const float r = s_rstd[g * kPT + p], m = s_mu[g * kPT + p];
    const float gm = __ldg(a.gamma + c), bt = __ldg(a.beta + c);
#pragma unroll
    for (int k = 0; k < kMaxHeads; ++k) {
      if (k < a.n_head) {
        const float sa = s_sa[p * kHP + k];
        // sum_t a (x*rstd - mu*rstd) * gamma + beta * sum_t a
        s_zs[k * L.zs_head_stride + c * kPT + p] = fmaf(gm, fmaf(acc[k], r, -m * sa), bt * sa);
      }
    }
  }
  __syncthreads();

  // ---- phase 5: o[d] = Wc[d,:] . z[h(d),:] + sa[h(d)] bc[d] + sum_t a[h(d),t] PE[b,t,d] --- tae.py:463,479,839
  for (int d = tid; d < a.D; d += kLtaeThreads) {
    const int hd = d / a.dh;
    float acc[kPT];
    const float* zh = s_zs + hd * L.zs_head_stride;
    if (a.has_inconv) {
      const float bias = __ldg(a.bc + d);
#pragma unroll
      for (int p = 0; p < kPT; ++p) acc[p] = bias * s_sa[p * kHP + hd];
#pragma unroll 4
      for (int c = 0; c < a.C; ++c) {
        const float w = __ldg(a.wct + static_cast<size_t>(c) * a.D + d);
        const float4 z0 = *reinterpret_cast<const float4*>(zh + c * kPT);
        const float4 z1 = *reinterpret_cast<const float4*>(zh + c * kPT + 4);
        acc[0] = fmaf(w, z0.x, acc[0]), acc[1] = fmaf(w, z0.y, acc[1]);
        acc[2] = fmaf(w, z0.z, acc[2]), acc[3] = fmaf(w, z0.w, acc[3]);
        acc[4] = fmaf(w, z1.x, acc[4]), acc[5] = fmaf(w, z1.y, acc[5]);
        acc[6] = fmaf(w, z1.z, acc[6]), acc[7] = fmaf(w, z1.w, acc[7]);
      }
    } else {
#pragma unroll
      for (int p = 0; p < kPT; ++p) acc[p] = zh[d * kPT + p];
    }
    if (a.pe != nullptr) {
      const float* pe = a.pe + static_cast<size_t>(b) * a.T * a.D + d;
      for (int t = 0; t < a.T; ++t) {
        const float v = __ldg(pe + static_cast<size_t>(t) * a.D);
#pragma unroll
        for (int p = 0; p < kPT; ++p) acc[p] = fmaf(s_sc[(t * kPT + p) * kHP + hd], v, acc[p]);
      }
    }
#pragma unroll
    for (int p = 0; p < kPT; ++p) s_os[d * kPT + p] = acc[p];
    if (a.save_o != nullptr)
      for (int p = 0; p < n_pix; ++p) a.save_o[(static_cast<size_t>(b) * a.hw + pix0 + p) * a.D + d] = acc[p];
  }
  __syncthreads();

  // ---- phase 6: MLP Linear (+ eval BatchNorm + ReLU) -------------------------------------- tae.py:442-447
  for (int j = tid; j < a.c_out; j += kLtaeThreads) {
    float acc[kPT];
    const float bias = __ldg(a.bm + j);
#pragma unroll
    for (int p = 0; p < kPT; ++p) acc[p] = bias;
#pragma unroll 4
    for (int d = 0; d < a.D; ++d) {
      const float w = __ldg(a.wmt + static_cast<size_t>(d) * a.c_out + j);
      const float4 o0 = *reinterpret_cast<const float4*>(s_os + d * kPT);
      const float4 o1 = *reinterpret_cast<const float4*>(s_os + d * kPT + 4);
      acc[0] = fmaf(w, o0.x, acc[0]), acc[1] = fmaf(w, o0.y, acc[1]);
      acc[2] = fmaf(w, o0.z, acc[2]), acc[3] = fmaf(w, o0.w, acc[3]);
      acc[4] = fmaf(w, o1.x, acc[4]), acc[5] = fmaf(w, o1.y, acc[5]);
      acc[6] = fmaf(w, o1.z, acc[6]), acc[7] = fmaf(w, o1.w, acc[7]);
    }
    if (a.bnf != nullptr) {
      const float sc = __ldg(a.bnf + j), sh = __ldg(a.bnf + a.c_out + j);
#pragma unroll
      for (int p = 0; p < kPT; ++p) {
        float v = fmaxf(fmaf(acc[p], sc, sh), 0.f);
        if (a.mlp_keep != nullptr && p < n_pix)
          v *= a.mlp_keep[(static_cast<size_t>(b) * a.c_out + j) * a.hw + pix0 + p] ? a.mlp_keep_scale : 0.f;
        s_ys[j * kPT + p] = v;
      }
    } else {
#pragma unroll
      for (int p = 0; p < kPT; ++p) s_ys[j * kPT + p] = acc[p];
    }
  }
  __syncthreads();

  if (a.bnf == nullptr) {  // training: BatchNorm needs statistics over every row of the batch first
    const size_t row0 = static_cast<size_t>(b) * a.hw + pix0;
    for (int item = tid; item < n_pix * a.c_out; item += kLtaeThreads) {
      const int p = item / a.c_out, j = item - p * a.c_out;
      a.ypre[(row0 + p) * a.c_out + j] = s_ys[j * kPT + p];
    }
    return;
  }

  // ---- phase 7: output GroupNorm over C'/h channels per pixel ------------------------------ tae.py:488
  for (int item = tid; item < a.n_head * kPT; item += kLtaeThreads) {
    const int g = item / kPT, p = item - g * kPT;
    float* yg = s_ys + g * a.cog * kPT + p;
    float m = 0.f;
    for (int k = 0; k < a.cog; ++k) m += yg[k * kPT];
    m /= static_cast<float>(a.cog);
    float var = 0.f;
    for (int k = 0; k < a.cog; ++k) {
      const float dlt = yg[k * kPT] - m;
      var = fmaf(dlt, dlt, var);
    }
    const float rstd = 1.f / sqrtf(var / static_cast<float>(a.cog) + a.gn_eps);
    for (int k = 0; k < a.cog; ++k) {
      const int j = g * a.cog + k;
      yg[k * kPT] = fmaf((yg[k * kPT] - m) * rstd, __ldg(a.on_w + j), __ldg(a.on_b + j));
    }
  }
  __syncthreads();
  T* ob = static_cast<T*>(a.out) + static_cast<size_t>(b) * a.c_out * a.hw + pix0;
  for (int item = tid; item < a.c_out * kPT; item += kLtaeThreads) {
    const int j = item / kPT, p = item - j * kPT;
    if (p < n_pix) Elem<T>::store(ob + static_cast<size_t>(j) * a.hw + p, s_ys[item]);
  }
}

// ---- training-mode epilogue: batch statistics -> BatchNorm -> ReLU -> output GroupNorm ------------------
// partial sums: part[2][c_out][parts], deterministic two-stage reduction (no atomics)
__global__ void bn_partial_kernel(const float* __restrict__ ypre, float* __restrict__ part, size_t n_rows, int c_out,
                                  int parts) {
  // blockIdx.x = part, threads over channels; rows strided by parts
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j >= c_out) return;
  const int pt = blockIdx.x;
  const size_t rows_per = (n_rows + parts - 1) / parts;
  const size_t r0 = pt * rows_per, r1 = min(n_rows, r0 + rows_per);
  // shifted by the first row of the batch for a stable variance
  const float pivot = ypre[j];
  float s1 = 0.f, s2 = 0.f;
  for (size_t r = r0; r < r1; ++r) {
    const float v = ypre[r * c_out + j] - pivot;
    s1 += v;
    s2 = fmaf(v, v, s2);
  }
  part[(0 * c_out + j) * parts + pt] = s1;
  part[(1 * c_out + j) * parts + pt] = s2;
}

__global__ void bn_finish_kernel(const float* __restrict__ ypre, const float* __restrict__ part, size_t n_rows,
                                 int c_out, int parts, float* __restrict__ mean, float* __restrict__ var) {
  // one warp per channel; lanes stride over the partial sums, fixed-order (deterministic) tree reduction in double
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= c_out) return;
  double s1 = 0.0, s2 = 0.0;
  for (int p = lane; p < parts; p += 32) {
    s1 += part[(0 * c_out + j) * parts + p];
    s2 += part[(1 * c_out + j) * parts + p];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane != 0) return;
  const double n = static_cast<double>(n_rows);
  const double m = s1 / n;
  double v = s2 / n - m * m;
  v = v < 0.0 ? 0.0 : v;
  mean[j] = static_cast<float>(m + static_cast<double>(ypre[j]));
  var[j] = static_cast<float>(v);  // biased, as used for normalisation (tae.py:445 in train mode)
}

template <typename T>
__global__ void bn_apply_kernel(const float* __restrict__ ypre, const float* __restrict__ mean,
                                const float* __restrict__ var, const float* __restrict__ bn_w,
                                const float* __restrict__ bn_b, const float* __restrict__ on_w,
                                const float* __restrict__ on_b, T* __restrict__ out, int B, int hw, int c_out,
                                int n_head, float bn_eps, float gn_eps, const uint8_t* __restrict__ keep,
                                float keep_scale) {
  // one thread per (row, group); rows = b*hw + pix
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t n_rows = static_cast<size_t>(B) * hw;
  if (i >= n_rows * n_head) return;
  const int g = static_cast<int>(i / n_rows);
  const size_t row = i - static_cast<size_t>(g) * n_rows;
  const int cog = c_out / n_head;
  const int b = static_cast<int>(row / hw), pix = static_cast<int>(row - static_cast<size_t>(b) * hw);
  auto act = [&](int j) {  // BatchNorm (batch statistics) -> ReLU -> dropout       tae.py:445-448
    float y = fmaxf((ypre[row * c_out + j] - mean[j]) / sqrtf(var[j] + bn_eps) * bn_w[j] + bn_b[j], 0.f);
    if (keep != nullptr) y *= keep[(static_cast<size_t>(b) * c_out + j) * hw + pix] ? keep_scale : 0.f;
    return y;
  };
  float m = 0.f;
  for (int k = 0; k < cog; ++k) m += act(g * cog + k);
  m /= static_cast<float>(cog);
  float v = 0.f;
  for (int k = 0; k < cog; ++k) {
    const float y = act(g * cog + k);
    v = fmaf(y - m, y - m, v);
  }
  const float rstd = 1.f / sqrtf(v / static_cast<float>(cog) + gn_eps);
  for (int k = 0; k < cog; ++k) {
    const int j = g * cog + k;
    const float y = act(j);
    Elem<T>::store(out + (static_cast<size_t>(b) * c_out + j) * hw + pix, fmaf((y - m) * rstd, on_w[j], on_b[j]));
  }
}

int launch_general(const c2s_ltae_desc& d, const c2s_ltae_params& p, const void* x, const uint8_t* pad_mask, void* out,
                   float* attn, float* ws, const LtaeWorkspace& lay, cudaStream_t stream) {
  const bool attn_only = (d.flags & C2S_LTAE_ATTN_ONLY) != 0;
  const bool train = (d.flags & C2S_LTAE_BN_BATCH_STATS) != 0 && !attn_only;
  const bool skip_attn = (d.flags & C2S_LTAE_SKIP_ATTN_STORE) != 0;
  const int hw = d.H * d.W;
  LtaeArgs a{};
  a.x = x, a.pad = pad_mask, a.out = out, a.attn = attn;
  a.u = ws + lay.u, a.cpos = ws + lay.cpos;
  a.wct = d.has_inconv ? ws + lay.wct : nullptr;
  a.bc = p.inconv_bias;
  a.wmt = ws + lay.wmt, a.bm = p.mlp_bias;
  a.pe = d.pe_mode != C2S_PE_NONE ? ws + lay.pe : nullptr;
  a.gamma = p.in_norm_weight, a.beta = p.in_norm_bias;
  a.bnf = (attn_only || train) ? nullptr : ws + lay.bnf;
  a.on_w = p.out_norm_weight, a.on_b = p.out_norm_bias;
  a.ypre = train ? (p.save_y != nullptr ? p.save_y : ws + lay.ypre) : nullptr;
  a.save_o = p.save_o;
  a.attn_keep = p.attn_keep, a.mlp_keep = p.mlp_keep;
  a.attn_keep_scale = d.attn_keep_scale, a.mlp_keep_scale = d.mlp_keep_scale;
  a.B = d.B, a.T = d.T, a.C = d.C, a.hw = hw;
  a.n_head = d.n_head, a.cpg = d.C / d.n_head, a.D = d.d_model, a.dh = d.d_model / d.n_head;
  a.c_out = attn_only ? 0 : d.c_out, a.cog = attn_only ? 0 : d.c_out / d.n_head;
  a.has_inconv = d.has_inconv, a.attn_only = attn_only, a.skip_attn_store = skip_attn;
  a.zero_padded = (d.flags & C2S_LTAE_ZERO_PADDED) != 0;
  a.gn_eps = d.gn_eps;
  a.tiles_per_b = ceil_div(hw, kPT);

  const LtaeSmem L = ltae_smem(d.T, d.C, d.d_model, a.c_out, d.n_head, attn_only);
  const size_t smem_bytes = static_cast<size_t>(L.total) * sizeof(float);
  if (smem_bytes > 227 * 1024)
    C2S_UNSUPPORTED("c2s_ltae_forward: T=%d, C=%d, d_model=%d need %zu B of shared memory per tile (max 232448)",
                    d.T, d.C, d.d_model, smem_bytes);
  const long long n_tiles = static_cast<long long>(d.B) * a.tiles_per_b;
  if (n_tiles > 0x7fffffffll) C2S_UNSUPPORTED("c2s_ltae_forward: too many pixel tiles");
  if (d.dtype == C2S_BF16) {
    C2S_SMEM_ATTR(ltae_forward_kernel<__nv_bfloat16>, 227 * 1024);
    ltae_forward_kernel<__nv_bfloat16><<<static_cast<unsigned>(n_tiles), kLtaeThreads, smem_bytes, stream>>>(a);
  } else {
    C2S_SMEM_ATTR(ltae_forward_kernel<float>, 227 * 1024);
    ltae_forward_kernel<float><<<static_cast<unsigned>(n_tiles), kLtaeThreads, smem_bytes, stream>>>(a);
  }
  C2S_LAUNCH_CHECK("ltae_forward<general>");

  return C2S_OK;
}

// Workspaces whose weight-only preparation was built by the persistent path (most recent 16).  Returns whether
// `ws` was in the set before the call; `mark` adds it, otherwise it is removed.
bool fa_prepared(const void* ws, bool mark) {
  static std::mutex mu;
  static const void* recent[16] = {};
  static int next = 0;
  std::lock_guard<std::mutex> lock(mu);
  int at = -1;
  for (int i = 0; i < 16; ++i)
    if (recent[i] == ws) at = i;
  if (mark && at < 0) {
    recent[next] = ws;
    next = (next + 1) % 16;
  } else if (!mark && at >= 0) {
    recent[at] = nullptr;
  }
  return at >= 0;
}

}  // namespace
}  // namespace c2s

extern "C" {

size_t c2s_ltae_workspace_bytes(const c2s_ltae_desc* d) {
  if (d == nullptr || d->n_head <= 0 || d->d_model <= 0) return 0;
  return c2s::ltae_workspace(*d).total * sizeof(float);
}

int c2s_ltae_forward(const c2s_ltae_desc* dp, const c2s_ltae_params* pp, const void* x, const void* positions,
                     const uint8_t* pad_mask, void* out, float* attn, float* bn_batch_mean, float* bn_batch_var,
                     void* workspace, size_t workspace_bytes, void* stream_ptr) {
  using namespace c2s;
  C2S_CHECK_ARG(dp != nullptr && pp != nullptr, "c2s_ltae_forward: desc/params is NULL");
  c2s_ltae_desc d = *dp;  // local copy: C2S_LTAE_REUSE_FOLDED is dropped when this workspace was not prepared that way
  const c2s_ltae_params& p = *pp;
  const bool attn_only = (d.flags & C2S_LTAE_ATTN_ONLY) != 0;
  const bool train = (d.flags & C2S_LTAE_BN_BATCH_STATS) != 0 && !attn_only;
  const bool skip_attn = (d.flags & C2S_LTAE_SKIP_ATTN_STORE) != 0;
  C2S_CHECK_ARG(x != nullptr, "c2s_ltae_forward: x is NULL");
  C2S_CHECK_ARG(d.B > 0 && d.T > 0 && d.C > 0 && d.H > 0 && d.W > 0,
                "c2s_ltae_forward: non-positive dimension in x[%d,%d,%d,%d,%d]", d.B, d.T, d.C, d.H, d.W);
  C2S_CHECK_ARG(d.dtype == C2S_F32 || d.dtype == C2S_BF16, "c2s_ltae_forward: unknown dtype %d", d.dtype);
  C2S_CHECK_ARG(d.n_head > 0 && d.d_k > 0 && d.d_model > 0, "c2s_ltae_forward: bad n_head/d_k/d_model");
  C2S_CHECK_ARG(d.C % d.n_head == 0, "c2s_ltae_forward: in_channels=%d not divisible by n_head=%d (GroupNorm)", d.C,
                d.n_head);
  C2S_CHECK_ARG(d.d_model % d.n_head == 0, "c2s_ltae_forward: d_model=%d not divisible by n_head=%d", d.d_model,
                d.n_head);
  C2S_CHECK_ARG(d.has_inconv || d.d_model == d.C, "c2s_ltae_forward: without inconv d_model must equal in_channels");
  C2S_CHECK_ARG(p.in_norm_weight && p.in_norm_bias && p.query && p.key_weight && p.key_bias,
                "c2s_ltae_forward: in_norm / attention_head parameters missing");
  C2S_CHECK_ARG(!d.has_inconv || (p.inconv_weight && p.inconv_bias), "c2s_ltae_forward: inconv parameters missing");
  C2S_CHECK_ARG(d.pe_mode >= C2S_PE_NONE && d.pe_mode <= C2S_PE_DOY_TABLE, "c2s_ltae_forward: unknown pe_mode %d",
                d.pe_mode);
  if (d.pe_mode != C2S_PE_NONE) {
    C2S_CHECK_ARG(positions != nullptr, "c2s_ltae_forward: positions is NULL but a positional encoder is configured");
    if (d.pe_mode != C2S_PE_DOY_TABLE) C2S_CHECK_ARG(p.pe_denom != nullptr, "c2s_ltae_forward: pe_denom missing");
    if (d.pe_mode != C2S_PE_SINUSOID)
      C2S_CHECK_ARG(p.pe_fc_weight && p.pe_fc_bias, "c2s_ltae_forward: positional_encoder.fc parameters missing");
    if (d.pe_abs)
      C2S_CHECK_ARG(p.pe_abs_fc_weight && p.pe_abs_fc_bias, "c2s_ltae_forward: positional_encoder_abs parameters missing");
  } else {
    C2S_CHECK_ARG(!d.pe_abs, "c2s_ltae_forward: pe_abs requires a positional encoder");
  }
  if (!attn_only) {
    C2S_CHECK_ARG(out != nullptr, "c2s_ltae_forward: out is NULL");
    C2S_CHECK_ARG(d.c_out > 0 && d.c_out % d.n_head == 0, "c2s_ltae_forward: mlp[-1]=%d not divisible by n_head=%d",
                  d.c_out, d.n_head);
    C2S_CHECK_ARG(p.mlp_weight && p.mlp_bias && p.bn_weight && p.bn_bias && p.out_norm_weight && p.out_norm_bias,
                  "c2s_ltae_forward: mlp / out_norm parameters missing");
    if (train)
      C2S_CHECK_ARG(bn_batch_mean && bn_batch_var, "c2s_ltae_forward: batch statistics outputs missing");
    else
      C2S_CHECK_ARG(p.bn_running_mean && p.bn_running_var, "c2s_ltae_forward: BatchNorm running statistics missing");
  }
  C2S_CHECK_ARG(attn != nullptr || skip_attn, "c2s_ltae_forward: attn is NULL without C2S_LTAE_SKIP_ATTN_STORE");
  C2S_CHECK_ARG(!(attn_only && skip_attn), "c2s_ltae_forward: ATTN_ONLY with SKIP_ATTN_STORE produces nothing");
  if (d.n_head > kMaxHeads) C2S_UNSUPPORTED("c2s_ltae_forward: n_head=%d exceeds the supported %d", d.n_head, kMaxHeads);
  if (d.T > 32767) C2S_UNSUPPORTED("c2s_ltae_forward: T=%d too large", d.T);
  int status = check_device();
  if (status != C2S_OK) return status;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_ptr);

  const LtaeWorkspace lay = ltae_workspace(d);
  C2S_CHECK_ARG(workspace != nullptr && workspace_bytes >= lay.total * sizeof(float),
                "c2s_ltae_forward: workspace of %zu bytes needed, %zu given", lay.total * sizeof(float),
                workspace_bytes);
  float* ws = static_cast<float*>(workspace);
  const int choice = option(C2S_OPT_LTAE_KERNEL);  // parity tests compare the kernels; 0 = automatic
  const bool use_fa = choice != C2S_LTAE_KERNEL_GENERAL && ltae_fa_eligible(d, x, out);
  // the reuse flag is honoured only if the previous call on this very workspace went through the same (persistent) path
  const bool was_prepared = fa_prepared(workspace, /*mark=*/use_fa);  // every call updates the set: another path unmarks
  if (!(use_fa && was_prepared)) d.flags &= ~C2S_LTAE_REUSE_FOLDED;
  status = ltae_prepare(d, p, positions, ws, lay, /*need_transposed=*/!use_fa, stream);
  if (status != C2S_OK) return status;

  const int hw = d.H * d.W;
  float* ypre = train ? (p.save_y != nullptr ? p.save_y : ws + lay.ypre) : nullptr;  // caller-owned when the backward wants it
  if (use_fa) {
    status = ltae_fa_forward(d, p, x, pad_mask, out, attn, ws, lay, ws + lay.fa, stream);
    if (status != C2S_OK) return status;
  } else {
    status = launch_general(d, p, x, pad_mask, out, attn, ws, lay, stream);
    if (status != C2S_OK) return status;
  }

  if (train) {
    const size_t n_rows = static_cast<size_t>(d.B) * hw;
    const int parts = static_cast<int>(n_rows < 1024 ? n_rows : 1024);
    float* part = ws + lay.bnpart;
    dim3 grid(parts, ceil_div(d.c_out, 128));
    bn_partial_kernel<<<grid, 128, 0, stream>>>(ypre, part, n_rows, d.c_out, parts);
    C2S_LAUNCH_CHECK("ltae_bn_partial");
    bn_finish_kernel<<<ceil_div(d.c_out, 8), 256, 0, stream>>>(ypre, part, n_rows, d.c_out, parts, bn_batch_mean,
                                                                 bn_batch_var);
    C2S_LAUNCH_CHECK("ltae_bn_finish");
    const size_t n_items = n_rows * d.n_head;
    if (d.dtype == C2S_BF16) {
      bn_apply_kernel<__nv_bfloat16><<<ceil_div(n_items, 256), 256, 0, stream>>>(
          ypre, bn_batch_mean, bn_batch_var, p.bn_weight, p.bn_bias, p.out_norm_weight, p.out_norm_bias,
          static_cast<__nv_bfloat16*>(out), d.B, hw, d.c_out, d.n_head, d.bn_eps, d.gn_eps, p.mlp_keep,
          d.mlp_keep_scale);
    } else {
      bn_apply_kernel<float><<<ceil_div(n_items, 256), 256, 0, stream>>>(
          ypre, bn_batch_mean, bn_batch_var, p.bn_weight, p.bn_bias, p.out_norm_weight, p.out_norm_bias,
          static_cast<float*>(out), d.B, hw, d.c_out, d.n_head, d.bn_eps, d.gn_eps, p.mlp_keep, d.mlp_keep_scale);
    }
    C2S_LAUNCH_CHECK("ltae_bn_apply");
  }
  return C2S_OK;
}

}  // extern "C"
