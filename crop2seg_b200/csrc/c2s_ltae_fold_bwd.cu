// Adjoint of the weight folding of c2s_ltae_prep.cu: the gradients of the folded quantities that c2s_ltae_backward
// returns (grad_U [C][16], grad_cpos [B][T][16]) become state_dict gradients (what autograd derives from
// tae.py:463-479, 760-778, 827-831 in the reference).  With
//     qk[h,:] = q_h^T Wk[h-block,:] / sqrt(dk),  m = qk Wc,  U[c,h] = gamma_c m[h,c],  wb = bc + Wc beta,
//     cpos[b,t,h] = qk[h,:] . (wb + PE[b,t,:]) + q_h . bk_h / sqrt(dk)
// the adjoint is
//     g_m[h,c]  = grad_U[c,h] gamma_c                     g_ub[h] = sum_bt grad_cpos[bt,h]
//     g_qk[h,d] = sum_c g_m[h,c] Wc[d,c] + g_ub[h] wb[d] + sum_bt grad_cpos[bt,h] PE[bt,d]
//     g_wb[d]   = sum_h g_ub[h] qk[h,d]
//     grad in_norm.weight[c] = sum_h grad_U[c,h] m[h,c]   (+ the direct term of c2s_ltae_backward)
//     grad in_norm.bias[c]   = sum_d Wc[d,c] g_wb[d]      (+ direct)
//     grad inconv.weight[d,c]= sum_h qk[h,d] g_m[h,c] + g_wb[d] beta_c   (+ direct, added by c2s_ltae_inconv_grad)
//     grad inconv.bias[d]    = g_wb[d]                                   (+ direct)
//     grad Q[h,j]            = (sum_d g_qk[h,d] Wk[h dk + j, d] + g_ub[h] bk[h dk + j]) / sqrt(dk)
//     grad fc1_k.weight[h dk + j, d] = q[h,j] g_qk[h,d] / sqrt(dk)     grad fc1_k.bias[h dk + j] = g_ub[h] q[h,j] / sqrt(dk)
// Two launches over [16, 256]-sized tensors (the reference-free part of the backward used to be ~30 eager torch ops).
// No atomics: every output element is reduced by one thread group in a fixed order.
#include "c2s_ltae_prep.cuh"

namespace c2s {
namespace {

constexpr int kFoldThreads = 256;

// block = (32 columns d, head h); the 8 warps split the reduction indices, lanes run over d (coalesced Wc^T / PE rows)
__global__ void __launch_bounds__(kFoldThreads) fold_bwd_qk_kernel(const float* __restrict__ g_u, const float* __restrict__ g_cpos,
                                                                    const float* __restrict__ gamma, const float* __restrict__ wct,
                                                                    const float* __restrict__ wb, const float* __restrict__ pe,
                                                                    float* __restrict__ g_qk, float* __restrict__ g_ub, int C, int D,
                                                                    int n_bt) {
  __shared__ float part[8][33];
  __shared__ float ub_part[8];
  const int h = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + lane;
  float s = 0.f, ub = 0.f;
  if (d < D) {
    for (int c = w; c < C; c += 8) s = fmaf(g_u[c * kMaxHeads + h] * gamma[c], wct[static_cast<size_t>(c) * D + d], s);
    for (int bt = w; bt < n_bt; bt += 8) {
      const float g = g_cpos[static_cast<size_t>(bt) * kMaxHeads + h];
      ub += g;
      if (pe != nullptr) s = fmaf(g, pe[static_cast<size_t>(bt) * D + d], s);
    }
  } else {
    for (int bt = w; bt < n_bt; bt += 8) ub += g_cpos[static_cast<size_t>(bt) * kMaxHeads + h];
  }
  part[w][lane] = s;
  if (lane == 0) ub_part[w] = ub;
  __syncthreads();
  if (w == 0) {
    float t = 0.f, u = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][lane], u += ub_part[i];
    if (d < D) g_qk[h * D + d] = fmaf(u, wb[d], t);
    if (blockIdx.x == 0 && lane == 0) g_ub[h] = u;
  }
}

struct FoldOut {
  float *in_norm_weight, *in_norm_bias, *inconv_weight, *inconv_bias, *query, *key_weight, *key_bias;
};

// blocks [0, D): row d of grad inconv.weight and grad inconv.bias[d]
// blocks [D, D + C): channel c of grad in_norm.weight / bias
// blocks [D + C, D + C + h): head h of grad Q, fc1_k.weight, fc1_k.bias
__global__ void __launch_bounds__(kFoldThreads) fold_bwd_out_kernel(
    const float* __restrict__ g_u, const float* __restrict__ g_qk, const float* __restrict__ g_ub, const float* __restrict__ qk,
    const float* __restrict__ wct, const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ q,
    const float* __restrict__ wk, const float* __restrict__ bk, const float* __restrict__ g_gamma_direct,
    const float* __restrict__ g_beta_direct, FoldOut o, int C, int D, int n_head, int dk) {
  __shared__ float red[kFoldThreads / 32];
  __shared__ float sh[kMaxHeads + 1];
  const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float rs = rsqrtf(static_cast<float>(dk));
  if (blk < D) {
    const int d = blk;
    if (tid < n_head) sh[tid] = qk[tid * D + d];
    __syncthreads();
    float gwb = 0.f;
    for (int h = 0; h < n_head; ++h) gwb = fmaf(g_ub[h], sh[h], gwb);
    if (tid == 0 && o.inconv_bias != nullptr) o.inconv_bias[d] = gwb;
    if (o.inconv_weight != nullptr)
      for (int c = tid; c < C; c += kFoldThreads) {
        float s = gwb * beta[c];
        const float gm = gamma[c];
        for (int h = 0; h < n_head; ++h) s = fmaf(sh[h], g_u[c * kMaxHeads + h] * gm, s);
        o.inconv_weight[static_cast<size_t>(d) * C + c] = s;
      }
  } else if (blk < D + C) {
    const int c = blk - D;
    // m[h,c] = sum_d qk[h,d] Wc[d,c]: warp w takes the heads w, w + 8; g_wb[d] = sum_h g_ub[h] qk[h,d] on the way for warp 0
    float gg = 0.f;
    for (int h = w; h < n_head; h += 8) {
      float m = 0.f;
      for (int d = lane; d < D; d += 32) m = fmaf(qk[h * D + d], wct[static_cast<size_t>(c) * D + d], m);
      m = warp_sum(m);
      gg = fmaf(g_u[c * kMaxHeads + h], m, gg);
    }
    if (lane == 0) red[w] = gg;
    float gb = 0.f;
    if (w == 0) {
      for (int d = lane; d < D; d += 32) {
        float gwb = 0.f;
        for (int h = 0; h < n_head; ++h) gwb = fmaf(g_ub[h], qk[h * D + d], gwb);
        gb = fmaf(wct[static_cast<size_t>(c) * D + d], gwb, gb);
      }
      gb = warp_sum(gb);
    }
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int i = 0; i < kFoldThreads / 32; ++i) t += red[i];
      if (o.in_norm_weight != nullptr) o.in_norm_weight[c] = t + (g_gamma_direct != nullptr ? g_gamma_direct[c] : 0.f);
      if (o.in_norm_bias != nullptr) o.in_norm_bias[c] = gb + (g_beta_direct != nullptr ? g_beta_direct[c] : 0.f);
    }
  } else {
    const int h = blk - D - C;
    const float ub = g_ub[h];
    for (int i = tid; i < dk * D; i += kFoldThreads) {
      const int j = i / D, d = i - j * D;
      if (o.key_weight != nullptr) o.key_weight[static_cast<size_t>(h * dk + j) * D + d] = q[h * dk + j] * g_qk[h * D + d] * rs;
    }
    for (int j = w; j < dk; j += 8) {  // one warp per query component
      float s = 0.f;
      for (int d = lane; d < D; d += 32) s = fmaf(g_qk[h * D + d], wk[static_cast<size_t>(h * dk + j) * D + d], s);
      s = warp_sum(s);
      if (lane == 0) {
        if (o.query != nullptr) o.query[h * dk + j] = (s + ub * bk[h * dk + j]) * rs;
        if (o.key_bias != nullptr) o.key_bias[h * dk + j] = ub * q[h * dk + j] * rs;
      }
    }
  }
}

// grad_pe[bt,d] += sum_h grad_cpos[bt,h] qk[h,d]   (the path of the positional table through the scores)
__global__ void fold_bwd_pe_kernel(const float* __restrict__ g_cpos, const float* __restrict__ qk, float* __restrict__ g_pe,
                                   int D, int n_head, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t bt = i / D;
  const int d = static_cast<int>(i - bt * D);
  float s = 0.f;
  for (int h = 0; h < n_head; ++h) s = fmaf(g_cpos[bt * kMaxHeads + h], qk[h * D + d], s);
  g_pe[i] += s;
}

}  // namespace
}  // namespace c2s

extern "C" int c2s_ltae_fold_backward(const c2s_ltae_desc* dp, const c2s_ltae_params* pp, const c2s_ltae_fold_bwd_io* iop,
                                      void* workspace, size_t workspace_bytes, void* stream_ptr) {
  using namespace c2s;
  C2S_CHECK_ARG(dp != nullptr && pp != nullptr && iop != nullptr, "c2s_ltae_fold_backward: desc/params/io is NULL");
  c2s_ltae_desc d = *dp;  // the same flag surgery as c2s_ltae_backward: the workspace layouts must agree
  d.flags &= ~C2S_LTAE_REUSE_FOLDED;
  d.flags |= C2S_LTAE_BN_BATCH_STATS;
  const c2s_ltae_params& p = *pp;
  const c2s_ltae_fold_bwd_io& io = *iop;
  C2S_CHECK_ARG(io.grad_u != nullptr && io.grad_cpos != nullptr, "c2s_ltae_fold_backward: grad_u / grad_cpos is NULL");
  C2S_CHECK_ARG(d.B > 0 && d.T > 0 && d.C > 0 && d.n_head > 0 && d.d_k > 0 && d.d_model > 0, "c2s_ltae_fold_backward: bad desc");
  if (!d.has_inconv) C2S_UNSUPPORTED("c2s_ltae_fold_backward: encoders without inconv (d_model=None) are not supported");
  if (d.n_head > kMaxHeads) C2S_UNSUPPORTED("c2s_ltae_fold_backward: n_head=%d exceeds the supported %d", d.n_head, kMaxHeads);
  C2S_CHECK_ARG(p.in_norm_weight && p.in_norm_bias && p.query && p.key_weight && p.key_bias && p.inconv_weight,
                "c2s_ltae_fold_backward: parameters missing");
  int status = check_device();
  if (status != C2S_OK) return status;
  const LtaeWorkspace lay = ltae_workspace(d);
  const size_t extra = align64(static_cast<size_t>(d.n_head) * d.d_model) + align64(kMaxHeads);
  C2S_CHECK_ARG(workspace != nullptr && workspace_bytes >= (lay.total + extra) * sizeof(float),
                "c2s_ltae_fold_backward: workspace of %zu bytes needed, %zu given", (lay.total + extra) * sizeof(float),
                workspace_bytes);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_ptr);
  float* ws = static_cast<float*>(workspace);
  float* g_qk = ws + lay.total;
  float* g_ub = g_qk + align64(static_cast<size_t>(d.n_head) * d.d_model);
  const int C = d.C, D = d.d_model, h = d.n_head, n_bt = d.B * d.T;
  const float* pe = d.pe_mode != C2S_PE_NONE ? ws + lay.pe : nullptr;
  fold_bwd_qk_kernel<<<dim3(ceil_div(D, 32), h), kFoldThreads, 0, stream>>>(io.grad_u, io.grad_cpos, p.in_norm_weight,
                                                                            ws + lay.wct, ws + lay.wb, pe, g_qk, g_ub, C, D, n_bt);
  C2S_LAUNCH_CHECK("ltae_fold_backward");
  FoldOut o{io.grad_in_norm_weight, io.grad_in_norm_bias, io.grad_inconv_weight, io.grad_inconv_bias,
            io.grad_query,          io.grad_key_weight,   io.grad_key_bias};
  fold_bwd_out_kernel<<<D + C + h, kFoldThreads, 0, stream>>>(io.grad_u, g_qk, g_ub, ws + lay.qk, ws + lay.wct, p.in_norm_weight,
                                                             p.in_norm_bias, p.query, p.key_weight, p.key_bias,
                                                             io.grad_gamma_direct, io.grad_beta_direct, o, C, D, h, d.d_k);
  C2S_LAUNCH_CHECK("ltae_fold_backward");
  if (io.grad_pe != nullptr && pe != nullptr) {
    const size_t n = static_cast<size_t>(n_bt) * D;
    fold_bwd_pe_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(io.grad_cpos, ws + lay.qk, io.grad_pe, D, h, n);
    C2S_LAUNCH_CHECK("ltae_fold_backward");
  }
  return C2S_OK;
}

