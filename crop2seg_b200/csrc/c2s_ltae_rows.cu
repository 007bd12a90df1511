// Rows behind the attention for encoders with SEVERAL learned queries (num_queries > 1), sm_100a.
// Reference: LTAE.forward, src/backbones/tae.py:486-499.  With n queries the attention head returns n rows per pixel
// [B*H*W, n, d_model]; mlp.0 / BatchNorm1d / ReLU act row by row, but out_norm = GroupNorm(n_head, c_out) is applied to
// the TRANSPOSED tensor [B*H*W, c_out, n] (tae.py:488): the statistics of a group run over its c_out / n_head channels
// AND the n queries.  The single-query kernels (c2s_ltae*.cu) produce the rows o of each query (params->save_o); this
// kernel finishes them jointly:
//     y[q][j]   = relu(BN_eval(mlp.0.weight[j,:] . o[q,:] + mlp.0.bias[j]))
//     out[b, q, j, pix] = (y[q][j] - mean_g) * rstd_g * out_norm.weight[j] + out_norm.bias[j],  g = j / (c_out / n_head),
//     mean_g, rstd_g over {y[q][j'] : all q, j' in group g}
// One warp per pixel; lanes run over d_model (coalesced rows of mlp.0.weight, which stays in L1/L2), one butterfly
// reduction per (query, channel).  Eval mode only: in training mode the reference's BatchNorm1d statistics run over the
// rows of all queries (tae.py:444-446) and the Python module refuses num_queries > 1.
#include "c2s_common.cuh"

namespace c2s {
namespace {

constexpr int kRowsWarps = 4;

struct RowsFwdArgs {
  const float* o;        // [n_q][N][D]
  const float* wm;       // [c_out][D]
  const float* bm;       // [c_out]
  const float* bn_w;
  const float* bn_b;
  const float* bn_mean;
  const float* bn_var;
  const float* on_w;
  const float* on_b;
  void* out;             // [B][n_q][c_out][hw]
  long long n_rows;      // B * hw
  int hw, D, c_out, n_head, n_q;
  float bn_eps, gn_eps;
};

template <typename T>
__global__ void __launch_bounds__(kRowsWarps * 32) ltae_rows_forward_kernel(const RowsFwdArgs a) {
  extern __shared__ float rows_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRowsWarps + warp;
  if (row >= a.n_rows) return;  // no CTA-wide barrier below
  float* y = rows_smem + static_cast<size_t>(warp) * (a.n_q * a.c_out + 2 * a.n_head);  // [n_q][c_out], mean[g], rstd[g]
  float* gm = y + a.n_q * a.c_out;
  float* gr = gm + a.n_head;
  for (int q = 0; q < a.n_q; ++q) {
    const float* o = a.o + (static_cast<size_t>(q) * a.n_rows + row) * a.D;
    for (int j = 0; j < a.c_out; ++j) {
      const float* w = a.wm + static_cast<size_t>(j) * a.D;
      float s = 0.f;
      for (int d = lane; d < a.D; d += 32) s = fmaf(__ldg(w + d), __ldg(o + d), s);
      s = warp_sum(s);
      if (lane == 0) {
        const float v = (s + a.bm[j] - a.bn_mean[j]) / sqrtf(a.bn_var[j] + a.bn_eps) * a.bn_w[j] + a.bn_b[j];  // tae.py:445
        y[q * a.c_out + j] = fmaxf(v, 0.f);                                                                      // tae.py:447
      }
    }
  }
  __syncwarp();
  const int cog = a.c_out / a.n_head;
  for (int g = lane; g < a.n_head; g += 32) {  // GroupNorm over (channels of the group) x (queries), tae.py:488
    float m = 0.f;
    for (int q = 0; q < a.n_q; ++q)
      for (int k = 0; k < cog; ++k) m += y[q * a.c_out + g * cog + k];
    m /= static_cast<float>(a.n_q * cog);
    float v = 0.f;
    for (int q = 0; q < a.n_q; ++q)
      for (int k = 0; k < cog; ++k) {
        const float dlt = y[q * a.c_out + g * cog + k] - m;
        v = fmaf(dlt, dlt, v);
      }
    gm[g] = m;
    gr[g] = 1.f / sqrtf(v / static_cast<float>(a.n_q * cog) + a.gn_eps);
  }
  __syncwarp();
  const long long b = row / a.hw;
  const int pix = static_cast<int>(row - b * a.hw);
  T* out = static_cast<T*>(a.out);
  for (int i = lane; i < a.n_q * a.c_out; i += 32) {
    const int q = i / a.c_out, j = i - q * a.c_out, g = j / cog;
    Elem<T>::store(out + ((static_cast<size_t>(b) * a.n_q + q) * a.c_out + j) * a.hw + pix,
                   fmaf((y[i] - gm[g]) * gr[g], a.on_w[j], a.on_b[j]));
  }
}

}  // namespace
}  // namespace c2s

extern "C" int c2s_ltae_rows_forward(const c2s_ltae_desc* desc, const c2s_ltae_params* params, const float* o_rows,
                                     int32_t n_queries, void* out, void* stream_ptr) {
  using namespace c2s;
  C2S_CHECK_ARG(desc != nullptr && params != nullptr && o_rows != nullptr && out != nullptr,
                "c2s_ltae_rows_forward: NULL argument");
  const c2s_ltae_desc& d = *desc;
  const c2s_ltae_params& p = *params;
  C2S_CHECK_ARG(d.B > 0 && d.H > 0 && d.W > 0 && d.d_model > 0 && d.n_head > 0 && n_queries > 0,
                "c2s_ltae_rows_forward: non-positive dimension");
  C2S_CHECK_ARG(d.c_out > 0 && d.c_out % d.n_head == 0, "c2s_ltae_rows_forward: mlp[-1]=%d not divisible by n_head=%d",
                d.c_out, d.n_head);
  C2S_CHECK_ARG(d.dtype == C2S_F32 || d.dtype == C2S_BF16, "c2s_ltae_rows_forward: unknown dtype %d", d.dtype);
  C2S_CHECK_ARG(!(d.flags & C2S_LTAE_BN_BATCH_STATS),
                "c2s_ltae_rows_forward: eval mode only (batch statistics over the rows of all queries are not served)");
  C2S_CHECK_ARG(p.mlp_weight && p.mlp_bias && p.bn_weight && p.bn_bias && p.bn_running_mean && p.bn_running_var &&
                    p.out_norm_weight && p.out_norm_bias,
                "c2s_ltae_rows_forward: mlp / BatchNorm / out_norm parameters missing");
  const size_t smem = static_cast<size_t>(kRowsWarps) * (static_cast<size_t>(n_queries) * d.c_out + 2 * d.n_head) * sizeof(float);
  if (smem > 48 * 1024) C2S_UNSUPPORTED("c2s_ltae_rows_forward: num_queries * mlp[-1] = %d exceeds 3000", n_queries * d.c_out);
  int status = check_device();
  if (status != C2S_OK) return status;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_ptr);
  RowsFwdArgs a{};
  a.o = o_rows, a.wm = p.mlp_weight, a.bm = p.mlp_bias, a.bn_w = p.bn_weight, a.bn_b = p.bn_bias;
  a.bn_mean = p.bn_running_mean, a.bn_var = p.bn_running_var, a.on_w = p.out_norm_weight, a.on_b = p.out_norm_bias;
  a.out = out;
  a.hw = d.H * d.W, a.n_rows = static_cast<long long>(d.B) * a.hw;
  a.D = d.d_model, a.c_out = d.c_out, a.n_head = d.n_head, a.n_q = n_queries;
  a.bn_eps = d.bn_eps, a.gn_eps = d.gn_eps;
  const unsigned grid = static_cast<unsigned>((a.n_rows + kRowsWarps - 1) / kRowsWarps);
  if (d.dtype == C2S_BF16)
    ltae_rows_forward_kernel<__nv_bfloat16><<<grid, kRowsWarps * 32, smem, stream>>>(a);
  else
    ltae_rows_forward_kernel<float><<<grid, kRowsWarps * 32, smem, stream>>>(a);
  C2S_LAUNCH_CHECK("ltae_rows_forward<queries>");
  return C2S_OK;
}
