// Weight folding and per-sample positional tables for the fused L-TAE kernels (sm_100a).
//
// The reference computes, per pixel row and frame (tae.py:461-479, 760-778, 822-839):
//     xn = GroupNorm(x);  e = Wc xn + bc + PE(b,t);  k = Wk e + bk;  s[h,t] = q_h . k_h / sqrt(d_k)
//     a = softmax_t(mask(s));  o[h-chunk] = sum_t a[h,t] e[t, h-chunk]
// Everything between xn and s is linear, so the kernels use (probed there
// against the reference):
//     s[h,t]     = U[h,:] . xn[t,:] + cpos[b,h,t]
//     o[h-chunk] = Wc[h-chunk,:] . (sum_t a[h,t] xn[t,:]) + (sum_t a[h,t]) bc[h-chunk] + sum_t a[h,t] PE[b,t,h-chunk]
// with  qk[h,:] = q_h^T Wk[h-block,:] / sqrt(d_k),  U = qk Wc,
//       cpos[b,h,t] = qk[h,:] . (bc + PE[b,t,:]) + q_h . bk[h-block] / sqrt(d_k).
// These small tensors are rebuilt on the device at every call (weights may have been updated by
// an optimizer step); nothing is synchronised with the host.
#pragma once

#include "c2s_common.cuh"

namespace c2s {

constexpr int kMaxHeads = 16;  // accumulators per thread in the attention kernels

// Workspace carve-up (floats).  Every block starts on a 64-float (256 B) boundary.
struct LtaeWorkspace {
  size_t qk;     // [h, D]
  size_t u;      // [C, kMaxHeads]   U transposed, in_norm.weight folded in, zero padded heads
  size_t wb;     // [D]              bc + Wc beta
  size_t ub;     // [kMaxHeads]      sum_c U[h,c] * in_norm.bias[c]  (+ the constant part of cpos)
  size_t wct;    // [C, D]           inconv.weight transposed
  size_t wmt;    // [D, c_out]       mlp.0.weight transposed
  size_t bnf;    // [2, c_out]       eval BatchNorm folded to scale / shift
  size_t pe;     // [B, T, D]        positional table (all encoders summed)
  size_t cpos;   // [B, T, kMaxHeads]
  size_t ypre;   // [B*H*W, c_out]   pre-BatchNorm MLP output (training mode only)
  size_t bnpart; // [2, c_out, parts] partial batch statistics (training mode only)
  size_t tc;     // tcgen05 MLP: o hi/lo rows + mlp weight hi/lo
  size_t fa;     // tensor-core attention kernels: score / in-projection weight fragments, scales, frame masks
  size_t total;  // floats
};

inline size_t align64(size_t n) { return (n + 63) & ~static_cast<size_t>(63); }
size_t ltae_mlp_tc_workspace_floats(const c2s_ltae_desc& d);
size_t ltae_fa_workspace_floats(const c2s_ltae_desc& d);

inline LtaeWorkspace ltae_workspace(const c2s_ltae_desc& d) {
  LtaeWorkspace w{};
  const size_t h = d.n_head, D = d.d_model, C = d.C;
  const bool attn_only = (d.flags & C2S_LTAE_ATTN_ONLY) != 0;
  const size_t co = attn_only ? 0 : d.c_out;
  size_t off = 0;
  auto take = [&](size_t n) {
    size_t o = off;
    off += align64(n);
    return o;
  };
  w.qk = take(h * D);
  w.u = take(C * kMaxHeads);
  w.wb = take(D);
  w.ub = take(kMaxHeads);
  w.wct = take(d.has_inconv ? C * D : 0);
  w.wmt = take(D * co);
  w.bnf = take(2 * co);
  w.pe = take(d.pe_mode != C2S_PE_NONE ? static_cast<size_t>(d.B) * d.T * D : 0);
  w.cpos = take(static_cast<size_t>(d.B) * d.T * kMaxHeads);
  const bool train = (d.flags & C2S_LTAE_BN_BATCH_STATS) != 0 && !attn_only;
  w.ypre = take(train ? static_cast<size_t>(d.B) * d.H * d.W * co : 0);
  w.bnpart = take(train ? 2 * co * 1024 : 0);
  w.tc = take(ltae_mlp_tc_workspace_floats(d));
  w.fa = take(ltae_fa_workspace_floats(d));
  w.total = off;
  return w;
}

// tcgen05 row GEMM for the MLP + BatchNorm + ReLU + output GroupNorm (c2s_ltae_mlp_tc.cu)
void ltae_mlp_tc_buffers(const c2s_ltae_desc& d, float* ws, __nv_bfloat16** o_hi, __nv_bfloat16** o_lo,
                         __nv_bfloat16** w_hi, __nv_bfloat16** w_lo);
int ltae_mlp_tc_forward(const c2s_ltae_desc& d, const c2s_ltae_params& p, float* tc_ws, const float* bnf, float* ypre,
                        void* out, cudaStream_t stream);

// persistent TMA-fed whole-slab kernel (c2s_ltae_fa.cu): bf16, 16 heads, d_model 256, C in {64, 128}, T <= 64, ...
bool ltae_fa_eligible(const c2s_ltae_desc& d, const void* x, const void* out);
int ltae_fa_forward(const c2s_ltae_desc& d, const c2s_ltae_params& p, const void* x, const uint8_t* pad_mask, void* out,
                    float* attn, float* ws, const LtaeWorkspace& lay, float* fa_ws, cudaStream_t stream);

int ltae_prepare(const c2s_ltae_desc& d, const c2s_ltae_params& p, const void* positions, float* ws,
                 const LtaeWorkspace& lay, bool need_transposed, cudaStream_t stream);

}  // namespace c2s
