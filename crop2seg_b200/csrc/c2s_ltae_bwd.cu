// L-TAE backward for sm_100a: the part of d loss / d (LTAE.forward inputs) that touches the [N, T, C] features
// (reference: autograd through src/backbones/tae.py:451-504, 760-847).
//
// With the folded algebra of c2s_ltae_prep.cuh the forward of one pixel row is
//     xh[t,c] = (x[t,c] - mean_g) rstd_g                                  (GroupNorm without its affine, tae.py:461)
//     s[h,t]  = sum_c U[h,c] xh[t,c] + cpos[b,h,t]      a = softmax_t(mask(s))      at = a * keep * scale  (tae.py:827-837)
//     zr[h,c] = sum_t at[h,t] xh[t,c]      sa[h] = sum_t at[h,t]      zn[h,c] = gamma_c zr[h,c] + beta_c sa[h]
//     o[d]    = sum_c Wc[d,c] zn[h(d),c] + sa[h(d)] bc[d] + sum_t at[h(d),t] PE[b,t,d]                  (tae.py:463,479,839)
// and everything after o (MLP, BatchNorm, ReLU, output GroupNorm: [N, 256] rows) is differentiated by the caller on
// those small rows.  Given grad_o[N, d_model] and grad_attn[h,B,T,H,W] this kernel recomputes the forward up to the
// attention and produces
//     grad_x                                   (GroupNorm backward included; padded frames get their GroupNorm share)
//     grad_U[c,h]   += sum_{n,t} gs[h,t] xh[t,c]          gs = softmax backward of  g_at * keep * scale
//     grad_cpos[b,t,h] += sum_{n in b} gs[h,t]
//     grad_gamma[c] += sum_{n,h} gzn[h,c] zr[h,c]       grad_beta[c] += sum_{n,h} gzn[h,c] sa[h]     (direct terms)
//     zn rows, sa rows   (the caller forms grad_Wc = grad_o^T zn and grad_bc = grad_o^T sa with a library GEMM)
//     grad_pe[b,t,d] += sum_{n in b} at[h(d),t] grad_o[n,d]                       (only when a table is learnable)
// where gzn[h,c] = sum_{d in head h} Wc[d,c] grad_o[d].  The chain from (grad_U, grad_cpos, grad_pe) to the state_dict
// tensors runs over [16, C] / [B, T, 16] sized tensors on the caller's side.
//
// One CTA = 8 consecutive pixels of one sample, fp32 math on the CUDA cores, fp32 or bf16 features (the general
// counterpart of ltae_forward_kernel; the training placements have N = B*256 pixel rows, where this is far from any
// hardware limit).  x is swept five times by the same CTA; all sweeps but the first hit L2.
#include "c2s_ltae_fa.cuh"

namespace c2s {
namespace {

constexpr int kPT = 8;
constexpr int kBwdThreads = 512;
constexpr int kHP = kMaxHeads + 4;
constexpr float kMaskFill = -1e6f;

struct BwdArgs {
  const void* x;
  const uint8_t* pad;
  const float* g_o;     // [N][D] or nullptr (attention only)
  const float* g_attn;  // [h][B][T][hw] or nullptr
  const float* u;       // [C][16]
  const float* cpos;    // [B][T][16]
  const float* wct;     // [C][D]
  const float* bc;      // [D]
  const float* pe;      // [B][T][D] or nullptr
  const float* gamma;
  const float* beta;
  const uint8_t* attn_keep;
  float attn_keep_scale;
  void* g_x;
  float* g_u;
  float* g_cpos;
  float* g_gamma;
  float* g_beta;
  float* zn_rows;
  float* sa_rows;
  float* g_pe;
  int B, T, C, hw, n_head, cpg, D, dh;
  int attn_only, zero_padded;
  float gn_eps;
  int tiles_per_b;
};

struct BwdSmem {
  int mu, rstd, u, sc, ga, gzn, go, gsa, sa, m1, m2, red, frames, total;  // offsets in floats
};

__host__ __device__ inline BwdSmem bwd_smem(int T, int C, int D, bool attn_only) {
  BwdSmem s{};
  int off = 0;
  auto take = [&](int n) {
    int o = off;
    off += (n + 3) & ~3;
    return o;
  };
  s.mu = take(kMaxHeads * kPT);
  s.rstd = take(kMaxHeads * kPT);
  s.u = take(C * kMaxHeads);
  s.sc = take(T * kPT * kHP);
  s.ga = take(T * kPT * kHP);
  s.gzn = take(attn_only ? 0 : C * kPT * kHP);
  s.go = take(attn_only ? 0 : D * kPT);
  s.gsa = take(kPT * kHP);
  s.sa = take(kPT * kHP);
  s.m1 = take(kMaxHeads * kPT);
  s.m2 = take(kMaxHeads * kPT);
  s.red = take(2 * kBwdThreads);
  s.frames = take((T + 3) / 4 + 4);  // uint8[T]: bit 0 = padded, bit 1 = read from memory
  s.total = off;
  return s;
}

constexpr int kSoftSplit = kBwdThreads / (kPT * kMaxHeads);  // frame slices per (pixel, head) column in the softmax passes
constexpr int kBatch = 8;                                     // channels per batch of stream_channels

// Walks over the channels of one (frame, pixel) column, group by group, in batches of up to kBatch channels of one group.
// The loads of the next batch are issued before the current one is handed to ``body(g, c0, n, values)``, so the L2
// latency of the 2-byte column reads overlaps the arithmetic instead of stalling it.
template <typename T, typename Body>
__device__ __forceinline__ void stream_channels(const T* __restrict__ xt, bool rd, int hw, int n_head, int cpg, Body&& body) {
  float cur[kBatch], nxt[kBatch];
  auto fetch = [&](float* dst, int g, int cb) {
    const int c0 = g * cpg + cb, n = min(kBatch, cpg - cb);
#pragma unroll
    for (int u = 0; u < kBatch; ++u) dst[u] = (rd && u < n) ? Elem<T>::load(xt + static_cast<size_t>(c0 + u) * hw) : 0.f;
  };
  int g = 0, cb = 0;
  fetch(cur, 0, 0);
  while (g < n_head) {
    int g2 = g, cb2 = cb + kBatch;
    if (cb2 >= cpg) cb2 = 0, ++g2;
    if (g2 < n_head) fetch(nxt, g2, cb2);
    body(g, g * cpg + cb, min(kBatch, cpg - cb), cur);
#pragma unroll
    for (int u = 0; u < kBatch; ++u) cur[u] = nxt[u];
    g = g2, cb = cb2;
  }
}

template <typename T>
__global__ void __launch_bounds__(kBwdThreads) ltae_backward_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const BwdSmem L = bwd_smem(a.T, a.C, a.D, a.attn_only != 0);
  float* s_mu = smem + L.mu;      // [g][p]  mean * rstd
  float* s_rstd = smem + L.rstd;  // [g][p]
  float* s_u = smem + L.u;        // [c][16]
  float* s_sc = smem + L.sc;      // [t][p][kHP]  scores -> a -> at
  float* s_ga = smem + L.ga;      // [t][p][kHP]  g_a -> gs
  float* s_gzn = smem + L.gzn;    // [c][p][kHP]
  float* s_go = smem + L.go;      // [d][p]
  float* s_gsa = smem + L.gsa;    // [p][kHP]
  float* s_sa = smem + L.sa;      // [p][kHP]
  float* s_m1 = smem + L.m1;      // [g][p]
  float* s_m2 = smem + L.m2;
  float* s_red = smem + L.red;    // [2][kSoftSplit][p][h] partial results of the softmax passes
  uint8_t* s_flag = reinterpret_cast<uint8_t*>(smem + L.frames);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / a.tiles_per_b;
  const int pix0 = (blockIdx.x - b * a.tiles_per_b) * kPT;
  const int n_pix = min(kPT, a.hw - pix0);
  const size_t frame_stride = static_cast<size_t>(a.C) * a.hw;
  const T* xb = static_cast<const T*>(a.x) + static_cast<size_t>(b) * a.T * frame_stride + pix0;
  const size_t row0 = static_cast<size_t>(b) * a.hw + pix0;

  // ---- phase 0: frame flags, folded score weights, this tile's grad_o rows -----------------------------------
  for (int t = tid; t < a.T; t += kBwdThreads) {
    const bool padded = a.pad != nullptr && a.pad[b * a.T + t] != 0;
    s_flag[t] = (padded ? 1 : 0) | ((padded && a.zero_padded) ? 0 : 2);
  }
  for (int i = tid; i < a.C * kMaxHeads; i += kBwdThreads) s_u[i] = a.u[i];
  for (int i = tid; i < kMaxHeads * kPT; i += kBwdThreads) s_m1[i] = 0.f, s_m2[i] = 0.f;
  if (!a.attn_only)
    for (int i = tid; i < a.D * kPT; i += kBwdThreads) {
      const int p = i / a.D, d = i - p * a.D;
      s_go[d * kPT + p] = p < n_pix ? a.g_o[(row0 + p) * a.D + d] : 0.f;
    }
  __syncthreads();
  // x[t, c, p] as the forward sees it: frames that are not read are zeros
  auto xval = [&](int t, int c, int p) -> float {
    return (s_flag[t] & 2) ? Elem<T>::load(xb + static_cast<size_t>(t) * frame_stride + static_cast<size_t>(c) * a.hw + p) : 0.f;
  };
  const bool vec4 = (a.dh & 3) == 0;  // float4 loads of the Wc^T / PE rows
  auto keepf = [&](int hh, int t, int p) -> float {
    if (a.attn_keep == nullptr) return 1.f;
    return a.attn_keep[((static_cast<size_t>(hh) * a.B + b) * a.T + t) * a.hw + pix0 + p] ? a.attn_keep_scale : 0.f;
  };

  // ---- phase 1: GroupNorm statistics (tae.py:461; all T frames, zero frames included) -------------------------
  // One thread = one (frame, pixel); sums are shifted by a pivot (first frame that is read, first channel of the group),
  // added over the 4 frames of a warp by shuffles and over the warps in shared memory.
  {
    int t_first = -1, n_read = 0;
    for (int t = 0; t < a.T; ++t)
      if (s_flag[t] & 2) {
        if (t_first < 0) t_first = t;
        ++n_read;
      }
    for (int it0 = 0; it0 < a.T * kPT; it0 += kBwdThreads) {  // uniform trip count: every lane joins the shuffles
      const int item = it0 + tid;
      const int t = item / kPT, p = item - t * kPT;
      const bool live = item < a.T * kPT && p < n_pix && (s_flag[t] & 2);
      const T* xt = xb + static_cast<size_t>(live ? t : 0) * frame_stride + p;
      const T* xp = xb + static_cast<size_t>(t_first < 0 ? 0 : t_first) * frame_stride + p;
      for (int g = 0, c = 0; g < a.n_head; ++g) {
        float s1 = 0.f, s2 = 0.f;
        if (live) {
          const float pivot = Elem<T>::load(xp + static_cast<size_t>(c) * a.hw);
#pragma unroll 8
          for (int cc = 0; cc < a.cpg; ++cc) {
            const float v = Elem<T>::load(xt + static_cast<size_t>(c + cc) * a.hw) - pivot;
            s1 += v;
            s2 = fmaf(v, v, s2);
          }
        }
        c += a.cpg;
#pragma unroll
        for (int o = kPT; o < 32; o <<= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane < kPT && lane < n_pix) {
          atomicAdd(s_m1 + g * kPT + lane, s1);
          atomicAdd(s_m2 + g * kPT + lane, s2);
        }
      }
    }
    __syncthreads();
    const float n_all = static_cast<float>(a.T) * a.cpg;
    for (int i = tid; i < a.n_head * kPT; i += kBwdThreads) {
      const int g = i / kPT, p = i - g * kPT;
      float s1 = s_m1[i], s2 = s_m2[i];
      s_m1[i] = 0.f, s_m2[i] = 0.f;  // phase 7 accumulates the GroupNorm-backward means here
      if (p >= n_pix) continue;
      const float pivot = t_first < 0 ? 0.f : xval(t_first, g * a.cpg, p);
      const float n_skip = n_all - static_cast<float>(n_read) * a.cpg;
      s1 -= n_skip * pivot;
      s2 = fmaf(n_skip * pivot, pivot, s2);
      const float m = s1 / n_all;
      float var = s2 / n_all - m * m;
      var = var < 0.f ? 0.f : var;
      const float rstd = 1.f / sqrtf(var + a.gn_eps);
      s_rstd[i] = rstd;
      s_mu[i] = (m + pivot) * rstd;
    }
  }
  __syncthreads();

  // ---- phase 2: scores (tae.py:827-831) -----------------------------------------------------------------------
  for (int item = tid; item < a.T * kPT; item += kBwdThreads) {
    const int t = item / kPT, p = item - t * kPT;
    float acc[kMaxHeads];
    if ((s_flag[t] & 1) || p >= n_pix) {
#pragma unroll
      for (int k = 0; k < kMaxHeads; ++k) acc[k] = kMaskFill;
    } else {
      const float4* cp = reinterpret_cast<const float4*>(a.cpos + (static_cast<size_t>(b) * a.T + t) * kMaxHeads);
#pragma unroll
      for (int k4 = 0; k4 < kMaxHeads / 4; ++k4) {
        const float4 c = __ldg(cp + k4);
        acc[4 * k4] = c.x, acc[4 * k4 + 1] = c.y, acc[4 * k4 + 2] = c.z, acc[4 * k4 + 3] = c.w;
      }
      const bool rd = (s_flag[t] & 2) != 0;  // hoisted: the loads of the unrolled loop go out back to back
      const T* xt = xb + static_cast<size_t>(t) * frame_stride + p;
      stream_channels(xt, rd, a.hw, a.n_head, a.cpg, [&](int g, int c0, int n, const float* xv) {
        const float r = s_rstd[g * kPT + p], m = s_mu[g * kPT + p];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (u >= n) break;
          const float xn = fmaf(xv[u], r, -m);
          const float4* up = reinterpret_cast<const float4*>(s_u + (c0 + u) * kMaxHeads);
#pragma unroll
          for (int k4 = 0; k4 < kMaxHeads / 4; ++k4) {
            const float4 w = up[k4];
            acc[4 * k4] = fmaf(w.x, xn, acc[4 * k4]);
            acc[4 * k4 + 1] = fmaf(w.y, xn, acc[4 * k4 + 1]);
            acc[4 * k4 + 2] = fmaf(w.z, xn, acc[4 * k4 + 2]);
            acc[4 * k4 + 3] = fmaf(w.w, xn, acc[4 * k4 + 3]);
          }
        }
      });
    }
    float4* sp = reinterpret_cast<float4*>(s_sc + (t * kPT + p) * kHP);
#pragma unroll
    for (int k4 = 0; k4 < kMaxHeads / 4; ++k4)
      sp[k4] = make_float4(acc[4 * k4], acc[4 * k4 + 1], acc[4 * k4 + 2], acc[4 * k4 + 3]);
  }
  __syncthreads();

  // ---- phase 3: softmax over T (tae.py:836): s_sc = a (before dropout).  Every (pixel, head) column is cut into
  //               kSoftSplit frame slices, one thread each; the slices meet through s_red ------------------------------
  const int col_id = tid % (kPT * kMaxHeads), slice = tid / (kPT * kMaxHeads);
  const int t_per = (a.T + kSoftSplit - 1) / kSoftSplit;
  const int t_lo = min(a.T, slice * t_per), t_hi = min(a.T, t_lo + t_per);
  const int col_p = col_id / kMaxHeads, col_h = col_id - col_p * kMaxHeads;
  const int col_stride = kPT * kHP;
  auto col_sum = [&](const float* part) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < kSoftSplit; ++q) v += part[q * (kPT * kMaxHeads) + col_id];
    return v;
  };
  {
    float* col = s_sc + col_p * kHP + col_h;
    float mx = -INFINITY;
    for (int t = t_lo; t < t_hi; ++t) mx = fmaxf(mx, col[t * col_stride]);
    s_red[tid] = mx;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kSoftSplit; ++q) mx = fmaxf(mx, s_red[q * (kPT * kMaxHeads) + col_id]);
    float den = 0.f;
    if (col_h < a.n_head)
      for (int t = t_lo; t < t_hi; ++t) {
        const float e = expf(col[t * col_stride] - mx);
        col[t * col_stride] = e;
        den += e;
      }
    s_red[kBwdThreads + tid] = den;
    __syncthreads();
    den = col_sum(s_red + kBwdThreads);
    const float inv = 1.f / den;
    for (int t = t_lo; t < t_hi; ++t) col[t * col_stride] = col_h < a.n_head ? col[t * col_stride] * inv : 0.f;
  }

  // ---- phase 4: gzn[h,c] = sum_{d in head h} Wc[d,c] grad_o[d]; g_sa[h] = bc . grad_o + beta . gzn -------------
  if (!a.attn_only) {
    for (int item = tid; item < a.C * kPT; item += kBwdThreads) {
      const int c = item / kPT, p = item - c * kPT;
      const float* w = a.wct + static_cast<size_t>(c) * a.D;
      float* dst = s_gzn + (c * kPT + p) * kHP;
      for (int hh = 0; hh < kMaxHeads; ++hh) {
        float acc = 0.f;
        if (hh < a.n_head) {
          if (vec4) {
            for (int i = 0; i < a.dh; i += 4) {
              const float4 wv = __ldg(reinterpret_cast<const float4*>(w + hh * a.dh + i));
              const float* gq = s_go + (hh * a.dh + i) * kPT + p;
              acc = fmaf(wv.x, gq[0], acc), acc = fmaf(wv.y, gq[kPT], acc);
              acc = fmaf(wv.z, gq[2 * kPT], acc), acc = fmaf(wv.w, gq[3 * kPT], acc);
            }
          } else {
            for (int i = 0; i < a.dh; ++i) acc = fmaf(__ldg(w + hh * a.dh + i), s_go[(hh * a.dh + i) * kPT + p], acc);
          }
        }
        dst[hh] = acc;
      }
    }
    __syncthreads();
    for (int item = tid; item < kPT * kMaxHeads; item += kBwdThreads) {
      const int p = item / kMaxHeads, hh = item - p * kMaxHeads;
      float acc = 0.f;
      if (hh < a.n_head) {
        for (int i = 0; i < a.dh; ++i) acc = fmaf(__ldg(a.bc + hh * a.dh + i), s_go[(hh * a.dh + i) * kPT + p], acc);
        for (int c = 0; c < a.C; ++c) acc = fmaf(__ldg(a.beta + c), s_gzn[(c * kPT + p) * kHP + hh], acc);
      }
      s_gsa[p * kHP + hh] = acc;
    }
  }
  __syncthreads();

  // ---- phase 5: g_at[h,t] = sum_c gamma_c gzn[h,c] xh[t,c] + g_sa[h] + sum_{d in h} grad_o[d] PE[b,t,d] + grad_attn;
  //               g_a = g_at * keep * scale -----------------------------------------------------------------------
  for (int item = tid; item < a.T * kPT; item += kBwdThreads) {
    const int t = item / kPT, p = item - t * kPT;
    float acc[kMaxHeads];
#pragma unroll
    for (int k = 0; k < kMaxHeads; ++k) acc[k] = 0.f;
    if (p < n_pix) {
      if (a.g_attn != nullptr) {
#pragma unroll
        for (int k = 0; k < kMaxHeads; ++k)
          if (k < a.n_head) acc[k] = a.g_attn[((static_cast<size_t>(k) * a.B + b) * a.T + t) * a.hw + pix0 + p];
      }
      if (!a.attn_only) {
#pragma unroll
        for (int k = 0; k < kMaxHeads; ++k) acc[k] += s_gsa[p * kHP + k];
        if (a.pe != nullptr) {
          const float* pe = a.pe + (static_cast<size_t>(b) * a.T + t) * a.D;
#pragma unroll
          for (int k = 0; k < kMaxHeads; ++k) {
            if (k < a.n_head) {
              float s = 0.f;
              if (vec4) {
                for (int i = 0; i < a.dh; i += 4) {
                  const float4 pv = __ldg(reinterpret_cast<const float4*>(pe + k * a.dh + i));
                  const float* gq = s_go + (k * a.dh + i) * kPT + p;
                  s = fmaf(gq[0], pv.x, s), s = fmaf(gq[kPT], pv.y, s);
                  s = fmaf(gq[2 * kPT], pv.z, s), s = fmaf(gq[3 * kPT], pv.w, s);
                }
              } else {
                for (int i = 0; i < a.dh; ++i) s = fmaf(s_go[(k * a.dh + i) * kPT + p], __ldg(pe + k * a.dh + i), s);
              }
              acc[k] += s;
            }
          }
        }
        const bool rd = (s_flag[t] & 2) != 0;
        const T* xt = xb + static_cast<size_t>(t) * frame_stride + p;
        stream_channels(xt, rd, a.hw, a.n_head, a.cpg, [&](int g, int c0, int n, const float* xv) {
          const float r = s_rstd[g * kPT + p], m = s_mu[g * kPT + p];
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            if (u >= n) break;
            const int c = c0 + u;
            const float xn = fmaf(xv[u], r, -m) * __ldg(a.gamma + c);
            const float4* gp = reinterpret_cast<const float4*>(s_gzn + (c * kPT + p) * kHP);
#pragma unroll
            for (int k4 = 0; k4 < kMaxHeads / 4; ++k4) {
              const float4 w = gp[k4];
              acc[4 * k4] = fmaf(w.x, xn, acc[4 * k4]);
              acc[4 * k4 + 1] = fmaf(w.y, xn, acc[4 * k4 + 1]);
              acc[4 * k4 + 2] = fmaf(w.z, xn, acc[4 * k4 + 2]);
              acc[4 * k4 + 3] = fmaf(w.w, xn, acc[4 * k4 + 3]);
            }
          }
        });
      }
#pragma unroll
      for (int k = 0; k < kMaxHeads; ++k) acc[k] = k < a.n_head ? acc[k] * keepf(k, t, p) : 0.f;
    }
    float4* sp = reinterpret_cast<float4*>(s_ga + (t * kPT + p) * kHP);
#pragma unroll
    for (int k4 = 0; k4 < kMaxHeads / 4; ++k4)
      sp[k4] = make_float4(acc[4 * k4], acc[4 * k4 + 1], acc[4 * k4 + 2], acc[4 * k4 + 3]);
  }
  __syncthreads();

  // ---- phase 6: softmax backward gs = a (g_a - sum_t a g_a); s_sc becomes at = a * keep * scale; sa = sum_t at ----
  {
    float* ac = s_sc + col_p * kHP + col_h;
    float* gc = s_ga + col_p * kHP + col_h;
    float dot = 0.f;
    for (int t = t_lo; t < t_hi; ++t) dot = fmaf(ac[t * col_stride], gc[t * col_stride], dot);
    s_red[tid] = dot;
    __syncthreads();
    dot = col_sum(s_red);
    float sa = 0.f, sgs = 0.f;
    const bool live = col_h < a.n_head && col_p < n_pix;
    for (int t = t_lo; t < t_hi; ++t) {
      const float av = ac[t * col_stride];
      const float gsv = av * (gc[t * col_stride] - dot);
      gc[t * col_stride] = gsv;
      sgs += gsv;
      const float at = live ? av * keepf(col_h, t, col_p) : 0.f;
      ac[t * col_stride] = at;
      sa += at;
    }
    __syncthreads();  // every thread has read the dot partials
    s_red[tid] = sa;
    s_red[kBwdThreads + tid] = sgs;
    __syncthreads();
    if (slice == 0) {
      sa = col_sum(s_red);
      s_sa[col_p * kHP + col_h] = sa;
      s_gsa[col_p * kHP + col_h] = col_sum(s_red + kBwdThreads);  // g_sa is dead after phase 5: now sum_t gs (zero up to rounding)
      if (a.sa_rows != nullptr && col_p < n_pix) a.sa_rows[(row0 + col_p) * kMaxHeads + col_h] = sa;
    }
  }
  __syncthreads();
  for (int item = tid; item < a.T * kMaxHeads; item += kBwdThreads) {  // grad_cpos[b,t,h] += sum_p gs
    const int t = item / kMaxHeads, hh = item - t * kMaxHeads;
    float s = 0.f;
    for (int p = 0; p < n_pix; ++p) s += s_ga[(t * kPT + p) * kHP + hh];
    if (hh < a.n_head) atomicAdd(a.g_cpos + (static_cast<size_t>(b) * a.T + t) * kMaxHeads + hh, s);
  }
  if (a.g_pe != nullptr && !a.attn_only) {  // grad_pe[b,t,d] += sum_p at[h(d),t] grad_o[d]
    for (int item = tid; item < a.T * a.D; item += kBwdThreads) {
      const int t = item / a.D, d = item - t * a.D, hd = d / a.dh;
      float s = 0.f;
      for (int p = 0; p < n_pix; ++p) s = fmaf(s_sc[(t * kPT + p) * kHP + hd], s_go[d * kPT + p], s);
      atomicAdd(a.g_pe + (static_cast<size_t>(b) * a.T + t) * a.D + d, s);
    }
  }

  // ---- phase 7: per (c, p): grad_U[c,h] += sum_t gs[h,t] xh[t,c];  zr[h,c] = sum_t at[h,t] xh[t,c] -> zn rows,
  //               direct gamma / beta terms ---------------------------------------------------------------------
  // One thread = one pixel x TWO channels (c and c + C/2): every gs / at row read from shared memory feeds 64 FMAs.
  const int c_half = (a.C + 1) >> 1;
  for (int it0 = 0; it0 < c_half * kPT; it0 += kBwdThreads) {  // uniform trip count: every lane joins the shuffles
    const int item = it0 + tid;
    const bool valid = item < c_half * kPT;
    const int cA = valid ? item / kPT : 0, p = valid ? item - cA * kPT : kPT;
    const int cc[2] = {cA, cA + c_half};
    const bool on[2] = {p < n_pix, p < n_pix && cc[1] < a.C};
    const int gq[2] = {cc[0] / a.cpg, on[1] ? cc[1] / a.cpg : 0};
    float au[2][kMaxHeads], az[2][kMaxHeads];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < kMaxHeads; ++k) au[j][k] = 0.f, az[j][k] = 0.f;
    if (on[0]) {
      float r[2], m[2];
      const T* xc[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        r[j] = s_rstd[gq[j] * kPT + p], m[j] = s_mu[gq[j] * kPT + p];
        xc[j] = xb + static_cast<size_t>(on[j] ? cc[j] : cc[0]) * a.hw + p;
      }
      constexpr int kF = 4;  // frames per batch; the next batch's loads are in flight while this one is consumed
      float cur[2][kF], nxt[2][kF];
      auto fetch = [&](float (*dst)[kF], int t0) {
#pragma unroll
        for (int u = 0; u < kF; ++u) {
          const bool rd = t0 + u < a.T && (s_flag[t0 + u] & 2);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            dst[j][u] = (rd && on[j]) ? Elem<T>::load(xc[j] + static_cast<size_t>(t0 + u) * frame_stride) : 0.f;
        }
      };
      fetch(cur, 0);
      for (int t0 = 0; t0 < a.T; t0 += kF) {
        if (t0 + kF < a.T) fetch(nxt, t0 + kF);
#pragma unroll
        for (int u = 0; u < kF; ++u) {
          const int t = t0 + u;
          if (t >= a.T) break;
          const float xn0 = fmaf(cur[0][u], r[0], -m[0]), xn1 = on[1] ? fmaf(cur[1][u], r[1], -m[1]) : 0.f;
          const float4* gp = reinterpret_cast<const float4*>(s_ga + (t * kPT + p) * kHP);
          const float4* ap = reinterpret_cast<const float4*>(s_sc + (t * kPT + p) * kHP);
#pragma unroll
          for (int k4 = 0; k4 < kMaxHeads / 4; ++k4) {
            const float4 w = gp[k4];
            au[0][4 * k4] = fmaf(w.x, xn0, au[0][4 * k4]), au[1][4 * k4] = fmaf(w.x, xn1, au[1][4 * k4]);
            au[0][4 * k4 + 1] = fmaf(w.y, xn0, au[0][4 * k4 + 1]), au[1][4 * k4 + 1] = fmaf(w.y, xn1, au[1][4 * k4 + 1]);
            au[0][4 * k4 + 2] = fmaf(w.z, xn0, au[0][4 * k4 + 2]), au[1][4 * k4 + 2] = fmaf(w.z, xn1, au[1][4 * k4 + 2]);
            au[0][4 * k4 + 3] = fmaf(w.w, xn0, au[0][4 * k4 + 3]), au[1][4 * k4 + 3] = fmaf(w.w, xn1, au[1][4 * k4 + 3]);
            if (!a.attn_only) {
              const float4 v = ap[k4];
              az[0][4 * k4] = fmaf(v.x, xn0, az[0][4 * k4]), az[1][4 * k4] = fmaf(v.x, xn1, az[1][4 * k4]);
              az[0][4 * k4 + 1] = fmaf(v.y, xn0, az[0][4 * k4 + 1]), az[1][4 * k4 + 1] = fmaf(v.y, xn1, az[1][4 * k4 + 1]);
              az[0][4 * k4 + 2] = fmaf(v.z, xn0, az[0][4 * k4 + 2]), az[1][4 * k4 + 2] = fmaf(v.z, xn1, az[1][4 * k4 + 2]);
              az[0][4 * k4 + 3] = fmaf(v.w, xn0, az[0][4 * k4 + 3]), az[1][4 * k4 + 3] = fmaf(v.w, xn1, az[1][4 * k4 + 3]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kF; ++u) cur[0][u] = nxt[0][u], cur[1][u] = nxt[1][u];
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = cc[j], g = gq[j];
      // GroupNorm-backward means without forming g_xh:  sum_t g_xh[t,c] = U[c,:] . sum_t gs + gamma_c gzn[c,:] . sa  and
      // sum_t g_xh[t,c] xh[t,c] = U[c,:] . au + gamma_c gzn[c,:] . az  (au, az are this pixel's sums, not yet reduced)
      float q1 = 0.f, q2 = 0.f;
      if (on[j]) {
#pragma unroll
        for (int k = 0; k < kMaxHeads; ++k) {
          const float uk = s_u[c * kMaxHeads + k];
          q1 = fmaf(uk, s_gsa[p * kHP + k], q1);
          q2 = fmaf(uk, au[j][k], q2);
        }
      }
      float gg = 0.f, gb = 0.f;
      if (!a.attn_only && on[j]) {
        const float gm = __ldg(a.gamma + c), bt = __ldg(a.beta + c);
#pragma unroll
        for (int k = 0; k < kMaxHeads; ++k) {
          if (k < a.n_head) {
            const float gz = s_gzn[(c * kPT + p) * kHP + k], sa = s_sa[p * kHP + k];
            gg = fmaf(gz, az[j][k], gg);
            gb = fmaf(gz, sa, gb);
            a.zn_rows[((row0 + p) * a.n_head + k) * a.C + c] = fmaf(gm, az[j][k], bt * sa);
          }
        }
        q1 = fmaf(gm, gb, q1);
        q2 = fmaf(gm, gg, q2);
      }
      if (on[j]) {
        atomicAdd(s_m1 + g * kPT + p, q1);
        atomicAdd(s_m2 + g * kPT + p, q2);
      }
      // the 8 lanes of a channel: sum over the pixels, one atomic per (channel, head)
#pragma unroll
      for (int o = 1; o < kPT; o <<= 1) {
        gg += __shfl_xor_sync(0xffffffffu, gg, o);
        gb += __shfl_xor_sync(0xffffffffu, gb, o);
#pragma unroll
        for (int k = 0; k < kMaxHeads; ++k) au[j][k] += __shfl_xor_sync(0xffffffffu, au[j][k], o);
      }
      if (valid && p == 0 && c < a.C) {
#pragma unroll
        for (int k = 0; k < kMaxHeads; ++k)
          if (k < a.n_head) atomicAdd(a.g_u + c * kMaxHeads + k, au[j][k]);
        if (!a.attn_only) {
          atomicAdd(a.g_gamma + c, gg);
          atomicAdd(a.g_beta + c, gb);
        }
      }
    }
  }

  // ---- phase 8: g_xh[t,c] = sum_h gs[h,t] U[h,c] + gamma_c sum_h at[h,t] gzn[h,c]; GroupNorm backward over (T, c in g):
  //               grad_x = rstd (g_xh - mean(g_xh) - xh mean(g_xh xh)) --------------------------------------------
  // One thread = one (frame, pixel): the gs / at rows stay in registers while it walks over the channels.  The two group
  // means come from phase 7, so g_xh is formed, corrected and stored in one sweep; x is loaded four channels ahead.
  T* __restrict__ gx = static_cast<T*>(a.g_x) + static_cast<size_t>(b) * a.T * frame_stride + pix0;
  __syncthreads();
  {
    const float inv_n = 1.f / (static_cast<float>(a.T) * a.cpg);
    for (int item = tid; item < a.T * kPT; item += kBwdThreads) {
      const int t = item / kPT, p = item - t * kPT;
      if (p >= n_pix) continue;
      float4 gs[kMaxHeads / 4], at[kMaxHeads / 4];
#pragma unroll
      for (int k4 = 0; k4 < kMaxHeads / 4; ++k4) {
        gs[k4] = reinterpret_cast<const float4*>(s_ga + (t * kPT + p) * kHP)[k4];
        at[k4] = reinterpret_cast<const float4*>(s_sc + (t * kPT + p) * kHP)[k4];
      }
      const bool rd = (s_flag[t] & 2) != 0;
      const T* __restrict__ xt = xb + static_cast<size_t>(t) * frame_stride + p;
      T* __restrict__ gt = gx + static_cast<size_t>(t) * frame_stride + p;
      stream_channels(xt, rd, a.hw, a.n_head, a.cpg, [&](int g, int c0, int n, const float* xv) {
        const float r = s_rstd[g * kPT + p], m = s_mu[g * kPT + p];
        const float m1 = s_m1[g * kPT + p] * inv_n, m2 = s_m2[g * kPT + p] * inv_n;
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (u >= n) break;
          const int c = c0 + u;
          const float4* up = reinterpret_cast<const float4*>(s_u + c * kMaxHeads);
          float v = 0.f;
#pragma unroll
          for (int k4 = 0; k4 < kMaxHeads / 4; ++k4) {
            const float4 uu = up[k4];
            v = fmaf(gs[k4].x, uu.x, v), v = fmaf(gs[k4].y, uu.y, v), v = fmaf(gs[k4].z, uu.z, v), v = fmaf(gs[k4].w, uu.w, v);
          }
          if (!a.attn_only) {
            const float4* zp = reinterpret_cast<const float4*>(s_gzn + (c * kPT + p) * kHP);
            float z = 0.f;
#pragma unroll
            for (int k4 = 0; k4 < kMaxHeads / 4; ++k4) {
              const float4 zz = zp[k4];
              z = fmaf(at[k4].x, zz.x, z), z = fmaf(at[k4].y, zz.y, z), z = fmaf(at[k4].z, zz.z, z), z = fmaf(at[k4].w, zz.w, z);
            }
            v = fmaf(__ldg(a.gamma + c), z, v);
          }
          Elem<T>::store(gt + static_cast<size_t>(c) * a.hw, r * (v - m1 - fmaf(xv[u], r, -m) * m2));
        }
      });
    }
  }
}

}  // namespace
}  // namespace c2s

extern "C" {

size_t c2s_ltae_backward_workspace_bytes(const c2s_ltae_desc* d) {
  if (d == nullptr) return 0;
  c2s_ltae_desc t = *d;  // the same flag surgery as c2s_ltae_backward
  t.flags &= ~C2S_LTAE_REUSE_FOLDED;
  t.flags |= C2S_LTAE_BN_BATCH_STATS;
  // + the scratch of c2s_ltae_fold_backward (g_qk [h][D], g_ub [16]), which runs on the same workspace
  return c2s_ltae_workspace_bytes(&t) +
         (c2s::align64(static_cast<size_t>(d->n_head) * d->d_model) + c2s::align64(c2s::kMaxHeads)) * sizeof(float);
}

int c2s_ltae_backward(const c2s_ltae_desc* dp, const c2s_ltae_params* pp, const void* x, const void* positions,
                      const uint8_t* pad_mask, const c2s_ltae_bwd_io* iop, void* workspace, size_t workspace_bytes,
                      void* stream_ptr) {
  using namespace c2s;
  C2S_CHECK_ARG(dp != nullptr && pp != nullptr && iop != nullptr, "c2s_ltae_backward: desc/params/io is NULL");
  c2s_ltae_desc d = *dp;
  d.flags &= ~C2S_LTAE_REUSE_FOLDED;
  d.flags |= C2S_LTAE_BN_BATCH_STATS;  // nothing behind o is evaluated here: skip the BatchNorm folding (running stats may be NULL)
  const c2s_ltae_params& p = *pp;
  const c2s_ltae_bwd_io& io = *iop;
  const bool attn_only = (d.flags & C2S_LTAE_ATTN_ONLY) != 0;
  C2S_CHECK_ARG(x != nullptr && io.grad_x != nullptr, "c2s_ltae_backward: x / grad_x is NULL");
  C2S_CHECK_ARG(d.B > 0 && d.T > 0 && d.C > 0 && d.H > 0 && d.W > 0, "c2s_ltae_backward: non-positive dimension");
  C2S_CHECK_ARG(d.dtype == C2S_F32 || d.dtype == C2S_BF16, "c2s_ltae_backward: unknown dtype %d", d.dtype);
  C2S_CHECK_ARG(d.n_head > 0 && d.C % d.n_head == 0 && d.d_model % d.n_head == 0, "c2s_ltae_backward: bad head count");
  C2S_CHECK_ARG(io.grad_u && io.grad_cpos, "c2s_ltae_backward: grad_u / grad_cpos is NULL");
  C2S_CHECK_ARG(attn_only || (io.grad_o && io.grad_gamma && io.grad_beta && io.zn_rows && io.sa_rows),
                "c2s_ltae_backward: grad_o / grad_gamma / grad_beta / zn_rows / sa_rows missing");
  C2S_CHECK_ARG(!attn_only || io.grad_attn != nullptr, "c2s_ltae_backward: ATTN_ONLY needs grad_attn");
  C2S_CHECK_ARG(p.in_norm_weight && p.in_norm_bias && p.query && p.key_weight && p.key_bias,
                "c2s_ltae_backward: in_norm / attention_head parameters missing");
  if (!d.has_inconv) C2S_UNSUPPORTED("c2s_ltae_backward: encoders without inconv (d_model=None) are not supported");
  C2S_CHECK_ARG(p.inconv_weight && p.inconv_bias, "c2s_ltae_backward: inconv parameters missing");
  if (d.n_head > kMaxHeads) C2S_UNSUPPORTED("c2s_ltae_backward: n_head=%d exceeds the supported %d", d.n_head, kMaxHeads);
  C2S_CHECK_ARG(d.pe_mode == C2S_PE_NONE || positions != nullptr, "c2s_ltae_backward: positions is NULL");
  int status = check_device();
  if (status != C2S_OK) return status;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_ptr);

  // the backward always needs the transposed in-projection weights: lay them out as a full (not attention-only) call
  c2s_ltae_desc dl = d;
  const LtaeWorkspace lay = ltae_workspace(dl);
  C2S_CHECK_ARG(workspace != nullptr && workspace_bytes >= lay.total * sizeof(float),
                "c2s_ltae_backward: workspace of %zu bytes needed, %zu given", lay.total * sizeof(float), workspace_bytes);
  float* ws = static_cast<float*>(workspace);
  status = ltae_prepare(d, p, positions, ws, lay, /*need_transposed=*/true, stream);
  if (status != C2S_OK) return status;

  const int hw = d.H * d.W;
  if (option(C2S_OPT_LTAE_BWD_KERNEL) == 0 && ltae_bwd_tc_eligible(d, x, io)) {  // bf16 shipped shapes: tensor-core kernel
    BwdTcArgs t{};
    t.x = static_cast<const __nv_bfloat16*>(x);
    t.g_o = io.grad_o, t.g_attn = io.grad_attn;
    t.u = ws + lay.u, t.cpos = ws + lay.cpos, t.wct = ws + lay.wct, t.wb = ws + lay.wb;
    t.pe = d.pe_mode != C2S_PE_NONE ? ws + lay.pe : nullptr;
    t.gamma = p.in_norm_weight, t.beta = p.in_norm_bias;
    t.pad = pad_mask, t.attn_keep = p.attn_keep, t.attn_keep_scale = d.attn_keep_scale;
    t.g_u = io.grad_u, t.g_cpos = io.grad_cpos, t.g_gamma = io.grad_gamma, t.g_beta = io.grad_beta;
    t.zn_rows = io.zn_rows, t.sa_rows = io.sa_rows;
    t.B = d.B, t.T = d.T, t.hw = hw;
    t.gn_eps = d.gn_eps;
    return ltae_bwd_tc_launch(d, t, io.grad_x, stream);
  }
  BwdArgs a{};
  a.x = x, a.pad = pad_mask, a.g_o = io.grad_o, a.g_attn = io.grad_attn;
  a.u = ws + lay.u, a.cpos = ws + lay.cpos, a.wct = ws + lay.wct, a.bc = p.inconv_bias;
  a.pe = d.pe_mode != C2S_PE_NONE ? ws + lay.pe : nullptr;
  a.gamma = p.in_norm_weight, a.beta = p.in_norm_bias;
  a.attn_keep = p.attn_keep, a.attn_keep_scale = d.attn_keep_scale;
  a.g_x = io.grad_x, a.g_u = io.grad_u, a.g_cpos = io.grad_cpos, a.g_gamma = io.grad_gamma, a.g_beta = io.grad_beta;
  a.zn_rows = io.zn_rows, a.sa_rows = io.sa_rows, a.g_pe = io.grad_pe;
  a.B = d.B, a.T = d.T, a.C = d.C, a.hw = hw;
  a.n_head = d.n_head, a.cpg = d.C / d.n_head, a.D = d.d_model, a.dh = d.d_model / d.n_head;
  a.attn_only = attn_only, a.zero_padded = (d.flags & C2S_LTAE_ZERO_PADDED) != 0;
  a.gn_eps = d.gn_eps;
  a.tiles_per_b = ceil_div(hw, kPT);
  const BwdSmem L = bwd_smem(d.T, d.C, d.d_model, attn_only);
  const size_t smem_bytes = static_cast<size_t>(L.total) * sizeof(float);
  if (smem_bytes > 227 * 1024)
    C2S_UNSUPPORTED("c2s_ltae_backward: T=%d, C=%d, d_model=%d need %zu B of shared memory per tile (max 232448)", d.T, d.C,
                    d.d_model, smem_bytes);
  const long long n_tiles = static_cast<long long>(d.B) * a.tiles_per_b;
  if (n_tiles > 0x7fffffffll) C2S_UNSUPPORTED("c2s_ltae_backward: too many pixel tiles");
  if (d.dtype == C2S_BF16) {
    C2S_SMEM_ATTR(ltae_backward_kernel<__nv_bfloat16>, 227 * 1024);
    ltae_backward_kernel<__nv_bfloat16><<<static_cast<unsigned>(n_tiles), kBwdThreads, smem_bytes, stream>>>(a);
  } else {
    C2S_SMEM_ATTR(ltae_backward_kernel<float>, 227 * 1024);
    ltae_backward_kernel<float><<<static_cast<unsigned>(n_tiles), kBwdThreads, smem_bytes, stream>>>(a);
  }
  C2S_LAUNCH_CHECK("ltae_backward<general>");
  return C2S_OK;
}

}  // extern "C"
