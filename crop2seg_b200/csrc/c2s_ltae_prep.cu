// Device-side weight folding and positional tables for the L-TAE kernels; see c2s_ltae_prep.cuh.
#include "c2s_ltae_prep.cuh"

namespace c2s {
namespace {

constexpr int kPrepThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (blockDim.x + 31) / 32; ++i) r += scratch[i];
  return r;
}

// qk[h,d] = sum_j Q[h,0,j] * Wk[h*dk + j, d] / sqrt(dk)        (tae.py:768, :827-828)
__global__ void fold_qk_kernel(const float* __restrict__ q, const float* __restrict__ wk, float* __restrict__ qk,
                               int n_head, int dk, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_head * D) return;
  const int h = i / D, d = i - h * D;
  float s = 0.f;
  for (int j = 0; j < dk; ++j) s = fmaf(q[h * dk + j], wk[static_cast<size_t>(h * dk + j) * D + d], s);
  qk[i] = s / sqrtf(static_cast<float>(dk));
}

// u[c, hh] = in_norm.weight[c] * sum_d qk[hh,d] Wc[d,c]   (zero for hh >= n_head)
// block = one head x 32 channels; 8 warps split d, lanes run over channels (coalesced Wc rows)
__global__ void fold_u_kernel(const float* __restrict__ qk, const float* __restrict__ wc,
                              const float* __restrict__ gamma, float* __restrict__ u, int n_head, int C, int D,
                              int has_inconv) {
  __shared__ float part[8][32];
  const int hh = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), dpart = threadIdx.x >> 5;
  float s = 0.f;
  if (hh < n_head && c < C) {
    if (has_inconv) {
      for (int d = dpart; d < D; d += 8) s = fmaf(qk[hh * D + d], wc[static_cast<size_t>(d) * C + c], s);
    } else if (dpart == 0) {
      s = qk[hh * D + c];
    }
  }
  part[dpart][threadIdx.x & 31] = s;
  __syncthreads();
  if (dpart == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    u[c * kMaxHeads + hh] = hh < n_head ? t * gamma[c] : 0.f;
  }
}

// wb[d] = bc[d] + sum_c Wc[d,c] beta[c]  (without inconv: beta[d]); one warp per d
__global__ void fold_wb_kernel(const float* __restrict__ wc, const float* __restrict__ bc,
                               const float* __restrict__ beta, float* __restrict__ wb, int C, int D, int has_inconv) {
  const int d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (d >= D) return;
  if (!has_inconv) {
    if (lane == 0) wb[d] = beta[d];
    return;
  }
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(wc[static_cast<size_t>(d) * C + c], beta[c], s);
  s = warp_sum(s);
  if (lane == 0) wb[d] = s + bc[d];
}

// ub[hh] = sum_d qk[hh,d] * wb[d] + q_h . bk[h-block] / sqrt(dk); one block per head
__global__ void fold_ub_kernel(const float* __restrict__ qk, const float* __restrict__ wb,
                               const float* __restrict__ q, const float* __restrict__ bk, float* __restrict__ ub,
                               int n_head, int dk, int D) {
  __shared__ float scratch[32];
  const int hh = blockIdx.x;
  float s = 0.f;
  if (hh < n_head)
    for (int d = threadIdx.x; d < D; d += blockDim.x) s = fmaf(qk[hh * D + d], wb[d], s);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    float r = 0.f;
    if (hh < n_head) {
      float qb = 0.f;
      for (int j = 0; j < dk; ++j) qb = fmaf(q[hh * dk + j], bk[hh * dk + j], qb);
      r = s + qb / sqrtf(static_cast<float>(dk));
    }
    ub[hh] = r;
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  // out[c, r] = in[r, c]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int c = i / rows, r = i - c * rows;
  out[i] = in[static_cast<size_t>(r) * cols + c];
}

// eval BatchNorm1d folded to y * scale + shift                       (tae.py:445)
__global__ void fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ b,
                               const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                               float* __restrict__ bnf, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float sc = w[i] / sqrtf(rv[i] + eps);
  bnf[i] = sc;
  bnf[n + i] = b[i] - rm[i] * sc;
}

template <typename P>
__device__ __forceinline__ float pos_as_float(const P* pos, size_t i) {
  return static_cast<float>(pos[i]);
}
template <typename P>
__device__ __forceinline__ int pos_as_doy(const P* pos, size_t i) {
  long long v = static_cast<long long>(pos[i]);  // .to(torch.int64) truncates (positional_encoding.py:63)
  v = v < 0 ? 0 : v;
  return static_cast<int>(v > 364 ? 364 : v);
}

constexpr int kPosRows = 4;
// pe[b,t,d] for kPosRows (b,t) rows per block                                 (positional_encoding.py:25-43, 58-73)
template <typename P>
__global__ void pos_table_kernel(const P* __restrict__ pos, int pos_stride, const float* __restrict__ denom,
                                 const float* __restrict__ fc_w, const float* __restrict__ fc_b,
                                 const float* __restrict__ abs_w, const float* __restrict__ abs_b,
                                 float* __restrict__ pe, int D, int dh, int pe_mode, int pe_abs,
                                 const float* __restrict__ qk, const float* __restrict__ ub,
                                 float* __restrict__ cpos, int n_head, size_t n_bt) {
  extern __shared__ float base[];  // [dh] un-tiled sinusoid table (add_linear only), then [D] the finished row
  float* row = base + dh;
  // kPosRows (b, t) rows per block: the qk rows of the cpos products stay in L1 / registers across them
  for (int rr = 0; rr < kPosRows; ++rr) {
    const size_t bt = static_cast<size_t>(blockIdx.x) * kPosRows + rr;
    if (bt >= n_bt) break;  // uniform
    const size_t pi = bt * pos_stride;
    if (pe_mode == C2S_PE_SINUSOID_LINEAR || pe_mode == C2S_PE_SINUSOID) {  // the table has dh distinct columns, tiled h times
      const float p = pos_as_float(pos, pi);
      for (int i = threadIdx.x; i < dh; i += blockDim.x) {
        const float a = p / denom[i];
        base[i] = (i & 1) ? cosf(a) : sinf(a);
      }
      __syncthreads();
    }
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      const int i = d % dh;
      float v = 0.f;
      if (pe_mode == C2S_PE_SINUSOID) {
        v = base[i];
      } else if (pe_mode == C2S_PE_SINUSOID_LINEAR) {
        v = fc_b[d];
        for (int k = 0; k < D; ++k) v = fmaf(fc_w[static_cast<size_t>(d) * D + k], base[k % dh], v);
      } else if (pe_mode == C2S_PE_DOY_TABLE) {
        v = fc_w[static_cast<size_t>(i) * 365 + pos_as_doy(pos, pi)] + fc_b[i];
      }
      if (pe_abs) v += abs_w[static_cast<size_t>(i) * 365 + pos_as_doy(pos, pi + 1)] + abs_b[i];
      pe[bt * D + d] = v;
      row[d] = v;
    }
    __syncthreads();
    // cpos[b,t,hh] = ub[hh] + qk[hh,:] . pe[b,t,:]: 16 lanes per head
    const int hh = threadIdx.x >> 4, part = threadIdx.x & 15;
    float s = 0.f;
    if (hh < n_head)
      for (int d = part; d < D; d += 16) s = fmaf(__ldg(qk + hh * D + d), row[d], s);
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (part == 0 && hh < kMaxHeads) cpos[bt * kMaxHeads + hh] = hh < n_head ? s + ub[hh] : 0.f;
    __syncthreads();  // the next row overwrites base / row
  }
}

// The shipped encoders (plain sinusoid, 16 table columns tiled over 16 heads, d_model 256): one WARP per (b, t) row.  The
// table repeats per head, so cpos[b,t,hh] = ub[hh] + sum_i base[i] * QS[hh][i] with QS[hh][i] = sum_h' qk[hh][16 h' + i]
// (formed once per block): 16 multiply-adds per head instead of 256, no block-wide barrier per row.
template <typename P>
__global__ void __launch_bounds__(256) pos_table_sin16_kernel(const P* __restrict__ pos, const float* __restrict__ denom,
                                                              float* __restrict__ pe, const float* __restrict__ qk,
                                                              const float* __restrict__ ub, float* __restrict__ cpos,
                                                              size_t n_bt) {
  __shared__ float qs[16][17];
  __shared__ float s_den[16], s_ub[16];
  {
    const int hh = threadIdx.x >> 4, i = threadIdx.x & 15;
    float s = 0.f;
#pragma unroll
    for (int h2 = 0; h2 < 16; ++h2) s += qk[hh * 256 + h2 * 16 + i];
    qs[hh][i] = s;
    if (threadIdx.x < 16) s_den[threadIdx.x] = denom[threadIdx.x], s_ub[threadIdx.x] = ub[threadIdx.x];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (size_t bt = static_cast<size_t>(blockIdx.x) * 8 + warp; bt < n_bt; bt += static_cast<size_t>(gridDim.x) * 8) {
    const float p = pos_as_float(pos, bt);
    const int i = lane & 15;
    const float a = p / s_den[i];
    const float base = (i & 1) ? cosf(a) : sinf(a);  // lanes 16-31 hold the same 16 columns
    // pe[bt][d] = base[d % 16]: lane writes d = 8 lane .. 8 lane + 7, i.e. columns (8 lane) % 16 .. + 7
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __shfl_sync(0xffffffffu, base, ((8 * lane) & 15) + k);
    float4* dst = reinterpret_cast<float4*>(pe + bt * 256 + 8 * lane);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s = fmaf(qs[i][k], __shfl_sync(0xffffffffu, base, k), s);
    if (lane < 16) cpos[bt * kMaxHeads + lane] = s + s_ub[lane];
  }
}

// without a positional encoder: cpos[b,t,hh] = ub[hh]
__global__ void fill_cpos_kernel(const float* __restrict__ ub, float* __restrict__ cpos, int n_head, size_t n_bt) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_bt * kMaxHeads) return;
  const int hh = static_cast<int>(i % kMaxHeads);
  cpos[i] = hh < n_head ? ub[hh] : 0.f;
}

}  // namespace

int ltae_prepare(const c2s_ltae_desc& d, const c2s_ltae_params& p, const void* positions, float* ws,
                 const LtaeWorkspace& lay, bool need_transposed, cudaStream_t stream) {
  const int h = d.n_head, D = d.d_model, C = d.C, dk = d.d_k, dh = D / h;
  const bool attn_only = (d.flags & C2S_LTAE_ATTN_ONLY) != 0;
  const bool has_pe = d.pe_mode != C2S_PE_NONE;
  float* qk = ws + lay.qk;
  if (has_pe && (h > kMaxHeads || kPrepThreads < 16 * h)) C2S_UNSUPPORTED("ltae_prepare: n_head=%d too large", h);

  const bool reuse = (d.flags & C2S_LTAE_REUSE_FOLDED) != 0;  // the weight-only blocks of ws are still valid
  if (!reuse) {
  fold_qk_kernel<<<ceil_div(h * D, kPrepThreads), kPrepThreads, 0, stream>>>(p.query, p.key_weight, qk, h, dk, D);
  C2S_LAUNCH_CHECK("ltae_fold_qk");
  fold_u_kernel<<<dim3(ceil_div(C, 32), kMaxHeads), 256, 0, stream>>>(qk, p.inconv_weight, p.in_norm_weight, ws + lay.u,
                                                                   h, C, D, d.has_inconv);
  C2S_LAUNCH_CHECK("ltae_fold_u");
  fold_wb_kernel<<<ceil_div(D, 8), 256, 0, stream>>>(p.inconv_weight, p.inconv_bias, p.in_norm_bias, ws + lay.wb, C, D,
                                                   d.has_inconv);
  C2S_LAUNCH_CHECK("ltae_fold_wb");
  fold_ub_kernel<<<kMaxHeads, kPrepThreads, 0, stream>>>(qk, ws + lay.wb, p.query, p.key_bias, ws + lay.ub, h, dk, D);
  C2S_LAUNCH_CHECK("ltae_fold_ub");
  if (need_transposed && d.has_inconv) {  // attention-only encoders too: c2s_ltae_fold_backward reads Wc^T
    transpose_kernel<<<ceil_div(D * C, kPrepThreads), kPrepThreads, 0, stream>>>(p.inconv_weight, ws + lay.wct, D, C);
    C2S_LAUNCH_CHECK("ltae_transpose_inconv");
  }
  if (!attn_only && need_transposed) {
    transpose_kernel<<<ceil_div(d.c_out * D, kPrepThreads), kPrepThreads, 0, stream>>>(p.mlp_weight, ws + lay.wmt,
                                                                                      d.c_out, D);
    C2S_LAUNCH_CHECK("ltae_transpose_mlp");
  }
  if (!attn_only) {
    if (!(d.flags & C2S_LTAE_BN_BATCH_STATS)) {
      fold_bn_kernel<<<ceil_div(d.c_out, kPrepThreads), kPrepThreads, 0, stream>>>(
          p.bn_weight, p.bn_bias, p.bn_running_mean, p.bn_running_var, d.bn_eps, ws + lay.bnf, d.c_out);
      C2S_LAUNCH_CHECK("ltae_fold_bn");
    }
  }
  }  // !reuse
  const size_t n_bt = static_cast<size_t>(d.B) * d.T;
  if (has_pe && d.pe_mode == C2S_PE_SINUSOID && !d.pe_abs && dh == 16 && D == 256 && h == kMaxHeads) {
    const unsigned blocks = static_cast<unsigned>(ceil_div(n_bt, 8) < 1184 ? ceil_div(n_bt, 8) : 1184);
    if (d.pos_dtype == 0)
      pos_table_sin16_kernel<long long><<<blocks, 256, 0, stream>>>(static_cast<const long long*>(positions), p.pe_denom,
                                                                    ws + lay.pe, qk, ws + lay.ub, ws + lay.cpos, n_bt);
    else
      pos_table_sin16_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(positions), p.pe_denom, ws + lay.pe, qk,
                                                                ws + lay.ub, ws + lay.cpos, n_bt);
    C2S_LAUNCH_CHECK("ltae_pos_table");
  } else if (has_pe) {
    const int stride = d.pe_abs ? 2 : 1;
    const size_t smem = static_cast<size_t>(dh + D) * sizeof(float);
    if (d.pos_dtype == 0) {
      pos_table_kernel<long long><<<static_cast<unsigned>(ceil_div(n_bt, kPosRows)), kPrepThreads, smem, stream>>>(
          static_cast<const long long*>(positions), stride, p.pe_denom, p.pe_fc_weight, p.pe_fc_bias,
          p.pe_abs_fc_weight, p.pe_abs_fc_bias, ws + lay.pe, D, dh, d.pe_mode, d.pe_abs, qk, ws + lay.ub,
          ws + lay.cpos, h, n_bt);
    } else {
      pos_table_kernel<float><<<static_cast<unsigned>(ceil_div(n_bt, kPosRows)), kPrepThreads, smem, stream>>>(
          static_cast<const float*>(positions), stride, p.pe_denom, p.pe_fc_weight, p.pe_fc_bias,
          p.pe_abs_fc_weight, p.pe_abs_fc_bias, ws + lay.pe, D, dh, d.pe_mode, d.pe_abs, qk, ws + lay.ub,
          ws + lay.cpos, h, n_bt);
    }
    C2S_LAUNCH_CHECK("ltae_pos_table");
  }
  if (!has_pe) {
    fill_cpos_kernel<<<ceil_div(n_bt * kMaxHeads, kPrepThreads), kPrepThreads, 0, stream>>>(ws + lay.ub, ws + lay.cpos, h,
                                                                                          n_bt);
    C2S_LAUNCH_CHECK("ltae_fill_cpos");
  }
  return C2S_OK;
}

}  // namespace c2s
