// Backward of the L-TAE rows behind the attention (reference: autograd through tae.py:442-449, 486-488):
//     y = Wm o + bm          (mlp.0, Linear d_model -> c_out)
//     yh = (y - mean) rstd   yb = yh bn_w + bn_b        (mlp.2, BatchNorm1d: batch statistics in training, running in eval)
//     r = relu(yb)           rd = r keep scale          (mlp.4 / mlp.5)
//     out = GroupNorm_16(rd) on_w + on_b                (out_norm)
// on the N = B*H*W pixel rows.  Given grad_out[B, c_out, H, W] and the rows o[N, d_model] saved by the forward it
// produces grad_o[N, d_model] (the input of c2s_ltae_backward) and the gradients of mlp.0, mlp.2 and out_norm.
// These are small dense problems (N x 256 x c_out): three plain shared-memory tiled fp32 GEMMs and two row kernels;
// nothing here is near a hardware limit at the training placements (N = 4096 rows per GPU).
#include "c2s_ltae_prep.cuh"

namespace c2s {
namespace {

constexpr int kTM = 64, kTN = 64, kTK = 16;  // GEMM tile
constexpr int kGemmThreads = 256;

// C[m][n] (+)= sum_k A(m,k) B(k,n), A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]; bias[n] added when given.
// K is split over blockIdx.z (atomic accumulation when gridDim.z > 1 or `accumulate`).
__global__ void __launch_bounds__(kGemmThreads)
sgemm_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B, long long sbk,
             long long sbn, float* __restrict__ Cm, long long ldc, const float* __restrict__ bias, int M, int N, int K,
             int k_per_split, int accumulate) {
  __shared__ __align__(16) float sA[kTK][kTM + 4];
  __shared__ __align__(16) float sB[kTK][kTN + 4];
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4] = {};
  // the next K slice travels from global memory into registers while the current one is multiplied
  constexpr int kPer = kTK * kTM / kGemmThreads;
  float ra[kPer], rb[kPer];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const int i = threadIdx.x + u * kGemmThreads;
      // the faster index follows the contiguous dimension of the operand
      const int ka = (sak == 1) ? i % kTK : i / kTM, mm = (sak == 1) ? i / kTK : i % kTM;
      const int kb = (sbk == 1) ? i % kTK : i / kTN, nn = (sbk == 1) ? i / kTK : i % kTN;
      const int m = m0 + mm, n = n0 + nn;
      ra[u] = (m < M && k0 + ka < k_end) ? A[m * sam + (k0 + ka) * sak] : 0.f;
      rb[u] = (n < N && k0 + kb < k_end) ? B[(k0 + kb) * sbk + n * sbn] : 0.f;
    }
  };
  if (k_begin < k_end) fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += kTK) {
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const int i = threadIdx.x + u * kGemmThreads;
      const int ka = (sak == 1) ? i % kTK : i / kTM, mm = (sak == 1) ? i / kTK : i % kTM;
      const int kb = (sbk == 1) ? i % kTK : i / kTN, nn = (sbk == 1) ? i / kTK : i % kTN;
      sA[ka][mm] = ra[u];
      sB[kb][nn] = rb[u];
    }
    __syncthreads();
    if (k0 + kTK < k_end) fetch(k0 + kTK);
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);  // rows are 272 bytes: 16-byte aligned
      const float4 bv = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr && blockIdx.z == 0) v += bias[n];
      if (accumulate || gridDim.z > 1) atomicAdd(Cm + m * ldc + n, v);
      else Cm[m * ldc + n] = v;
    }
  }
}

struct RowsArgs {
  const float* y;         // [N][c_out] pre-BatchNorm rows
  const void* g_out;      // [B][c_out][hw]
  const float* mean;      // [c_out] batch or running mean
  const float* var;       // [c_out] (biased) batch or running variance
  const float* bn_w;
  const float* bn_b;
  const float* on_w;
  const uint8_t* keep;    // [B][c_out][hw] or nullptr
  float keep_scale, bn_eps, gn_eps;
  float* g_yh;            // out [N][c_out]: d loss / d yh (training) or d loss / d y (eval)
  float* g_bn_w;
  float* g_bn_b;
  float* g_on_w;
  float* g_on_b;
  float* colsum;          // training: [2][c_out] sum_rows g_yh, sum_rows g_yh * yh
  float* g_bm;            // eval: d loss / d mlp.0.bias accumulates here
  int B, hw, c_out, cog, n_groups, train;
};

// one thread per (row, group); the 32 rows of a warp are consecutive pixels of one channel group, so the strided
// reads of grad_out / keep are coalesced and the per-channel sums are reduced in the warp before the atomics
template <typename T>
__global__ void mlp_rows_backward_kernel(const RowsArgs a) {
  const int n_rows = a.B * a.hw;
  const int warps_per_group = (n_rows + 31) / 32;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wid >= warps_per_group * a.n_groups) return;  // warp-uniform
  const int g = wid / warps_per_group;
  const int row = (wid - g * warps_per_group) * 32 + lane;
  const bool live = row < n_rows;
  const int b = live ? row / a.hw : 0, pix = live ? row - b * a.hw : 0;
  const int j0 = g * a.cog;
  constexpr int kMaxCog = 16;  // c_out <= 256
  float yh[kMaxCog], rd[kMaxCog], gz[kMaxCog];
  float m = 0.f;
#pragma unroll
  for (int q = 0; q < kMaxCog; ++q) {
    yh[q] = rd[q] = gz[q] = 0.f;
    if (q < a.cog && live) {
      const int j = j0 + q;
      const float rstd = rsqrtf(a.var[j] + a.bn_eps);
      yh[q] = (a.y[static_cast<size_t>(row) * a.c_out + j] - a.mean[j]) * rstd;
      float v = fmaxf(fmaf(yh[q], a.bn_w[j], a.bn_b[j]), 0.f);
      const size_t oi = (static_cast<size_t>(b) * a.c_out + j) * a.hw + pix;
      if (a.keep != nullptr) v *= a.keep[oi] ? a.keep_scale : 0.f;
      rd[q] = v;
      m += v;
      gz[q] = Elem<T>::load(static_cast<const T*>(a.g_out) + oi);
    }
  }
  m /= static_cast<float>(a.cog);
  float var = 0.f;
#pragma unroll
  for (int q = 0; q < kMaxCog; ++q)
    if (q < a.cog) var = fmaf(rd[q] - m, rd[q] - m, var);
  const float rg = rsqrtf(var / static_cast<float>(a.cog) + a.gn_eps);
  // GroupNorm backward: zh = (rd - m) rg; gh = gz on_w; g_rd = rg (gh - mean(gh) - zh mean(gh zh))
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int q = 0; q < kMaxCog; ++q)
    if (q < a.cog) {
      const float zh = (rd[q] - m) * rg, gh = gz[q] * a.on_w[j0 + q];
      s1 += gh;
      s2 = fmaf(gh, zh, s2);
    }
  s1 /= static_cast<float>(a.cog), s2 /= static_cast<float>(a.cog);
#pragma unroll
  for (int q = 0; q < kMaxCog; ++q) {
    if (q >= a.cog) continue;  // uniform
    const int j = j0 + q;
    const float zh = (rd[q] - m) * rg, gh = gz[q] * a.on_w[j];
    float g_rd = live ? rg * (gh - s1 - zh * s2) : 0.f;
    float p_onw = live ? gz[q] * zh : 0.f, p_onb = live ? gz[q] : 0.f;
    if (a.keep != nullptr && live)
      g_rd *= a.keep[(static_cast<size_t>(b) * a.c_out + j) * a.hw + pix] ? a.keep_scale : 0.f;
    const float yb = fmaf(yh[q], a.bn_w[j], a.bn_b[j]);
    const float g_yb = (live && yb > 0.f) ? g_rd : 0.f;
    float p_bnw = g_yb * yh[q], p_bnb = g_yb;
    const float g_yh = g_yb * a.bn_w[j];
    float out = g_yh, c1 = g_yh, c2 = g_yh * yh[q];
    if (!a.train) out = g_yh * rsqrtf(a.var[j] + a.bn_eps);  // running statistics are constants: this is d loss / d y
    if (live) a.g_yh[static_cast<size_t>(row) * a.c_out + j] = out;
    float p_bm = out;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      p_onw += __shfl_xor_sync(0xffffffffu, p_onw, o);
      p_onb += __shfl_xor_sync(0xffffffffu, p_onb, o);
      p_bnw += __shfl_xor_sync(0xffffffffu, p_bnw, o);
      p_bnb += __shfl_xor_sync(0xffffffffu, p_bnb, o);
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o);
      p_bm += __shfl_xor_sync(0xffffffffu, p_bm, o);
    }
    if (lane == 0) {
      atomicAdd(a.g_on_w + j, p_onw);
      atomicAdd(a.g_on_b + j, p_onb);
      atomicAdd(a.g_bn_w + j, p_bnw);
      atomicAdd(a.g_bn_b + j, p_bnb);
      if (a.train) {
        atomicAdd(a.colsum + j, c1);
        atomicAdd(a.colsum + a.c_out + j, c2);
      } else {
        atomicAdd(a.g_bm + j, p_bm);
      }
    }
  }
}

// training: BatchNorm backward through the batch statistics, in place: g_y = rstd (g_yh - mean_rows(g_yh) - yh mean_rows(g_yh yh));
// also d loss / d mlp.0.bias = sum_rows g_y
__global__ void bn_batch_backward_kernel(float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ mean,
                                         const float* __restrict__ var, const float* __restrict__ colsum, float bn_eps,
                                         float* __restrict__ g_bm, int n_rows, int c_out) {
  // block = 32 rows x c_out channels (threads over channels): coalesced rows, one atomic per (block, channel)
  const int r0 = blockIdx.x * 32;
  for (int j = threadIdx.x; j < c_out; j += blockDim.x) {
    const float rstd = rsqrtf(var[j] + bn_eps), mu = mean[j];
    const float m1 = colsum[j] / static_cast<float>(n_rows), m2 = colsum[c_out + j] / static_cast<float>(n_rows);
    float s = 0.f;
    for (int r = r0; r < min(n_rows, r0 + 32); ++r) {
      const size_t i = static_cast<size_t>(r) * c_out + j;
      const float yh = (y[i] - mu) * rstd;
      const float v = rstd * (g[i] - m1 - yh * m2);
      g[i] = v;
      s += v;
    }
    atomicAdd(g_bm + j, s);
  }
}

// Direct term of the in-projection gradient (stage F of DESIGN.md section 4.6):
//   grad_Wc[h*dh + i][c] += sum_rows grad_o[row][h*dh + i] * zn[row][h][c],   grad_bc[h*dh + i] += sum_rows grad_o[..] * sa[row][h]
// One CTA = one head x one slice of rows; thread = (channel, half of the dh outputs of the head).
constexpr int kHoThreads = 256, kHoRows = 32, kHoMaxDh = 32, kHoMaxC = 256;
template <int PER>  // outputs of the head per thread (dh / parts, rounded up)
__global__ void __launch_bounds__(kHoThreads)
head_outer_kernel(const float* __restrict__ g_o, const float* __restrict__ zn, const float* __restrict__ sa, float* __restrict__ g_wc,
                  float* __restrict__ g_bc, int n_rows, int n_head, int dh, int C, int rows_per_cta) {
  __shared__ float s_g[kHoRows][kHoMaxDh];
  __shared__ float s_sa[kHoRows];
  __shared__ __align__(16) float s_z[kHoRows][kHoMaxC];  // this head's zn rows of the chunk (coalesced loads, reused by every part)
  const int h = blockIdx.x, D = n_head * dh;
  const int r_begin = blockIdx.y * rows_per_cta, r_end = min(n_rows, r_begin + rows_per_cta);
  // thread = (channel slot, part of the head's dh outputs); a slot owns channels c and c + n_c
  const int n_c = (C + 1) / 2;
  const int parts = (dh + PER - 1) / PER, per = PER;
  const int part = threadIdx.x / n_c, c = threadIdx.x - part * n_c;
  const bool busy = part < parts;
  const int i0 = part * per, i1 = min(dh, i0 + per);
  float acc0[PER] = {}, acc1[PER] = {};
  float accb = 0.f;
  const int c0 = c, c1 = c + n_c;
  for (int r0 = r_begin; r0 < r_end; r0 += kHoRows) {
    const int nr = min(kHoRows, r_end - r0);
    for (int i = threadIdx.x; i < nr * kHoMaxDh; i += kHoThreads) {
      const int r = i / kHoMaxDh, j = i - r * kHoMaxDh;
      s_g[r][j] = j < dh ? g_o[static_cast<size_t>(r0 + r) * D + h * dh + j] : 0.f;
    }
    for (int r = threadIdx.x; r < nr; r += kHoThreads) s_sa[r] = sa ? sa[static_cast<size_t>(r0 + r) * 16 + h] : 0.f;
    if ((C & 3) == 0) {  // 16-byte loads, four in flight per thread
      const int c4 = C >> 2, n4 = nr * c4;
      for (int base = 0; base < n4; base += 4 * kHoThreads) {
        float4 v[4];
        int rr[4], cc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = base + u * kHoThreads + threadIdx.x;
          rr[u] = i / c4, cc[u] = i - rr[u] * c4;
          if (i < n4) v[u] = __ldg(reinterpret_cast<const float4*>(zn + (static_cast<size_t>(r0 + rr[u]) * n_head + h) * C) + cc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (base + u * kHoThreads + threadIdx.x < n4) *reinterpret_cast<float4*>(&s_z[rr[u]][cc[u] * 4]) = v[u];
      }
    } else {
      for (int i = threadIdx.x; i < nr * C; i += kHoThreads) {
        const int r = i / C, cc = i - r * C;
        s_z[r][cc] = zn[(static_cast<size_t>(r0 + r) * n_head + h) * C + cc];
      }
    }
    __syncthreads();
    if (busy) {
      for (int r = 0; r < nr; ++r) {
        const float z0 = c0 < C ? s_z[r][c0] : 0.f, z1 = c1 < C ? s_z[r][c1] : 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const float gv = s_g[r][i0 + j];  // rows are kHoMaxDh wide and zero beyond dh
          acc0[j] = fmaf(gv, z0, acc0[j]);
          acc1[j] = fmaf(gv, z1, acc1[j]);
        }
      }
    }
    if (g_bc != nullptr && threadIdx.x < dh)
      for (int r = 0; r < nr; ++r) accb = fmaf(s_g[r][threadIdx.x], s_sa[r], accb);
    __syncthreads();
  }
  if (busy) {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (i0 + j < i1) {
        if (c0 < C) atomicAdd(g_wc + static_cast<size_t>(h * dh + i0 + j) * C + c0, acc0[j]);
        if (c1 < C) atomicAdd(g_wc + static_cast<size_t>(h * dh + i0 + j) * C + c1, acc1[j]);
      }
    }
  }
  if (g_bc != nullptr && threadIdx.x < dh) atomicAdd(g_bc + h * dh + threadIdx.x, accb);
}

int launch_gemm(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* Cm,
                long long ldc, const float* bias, int M, int N, int K, int splits, bool accumulate, cudaStream_t stream,
                const char* name) {
  int k_per = ceil_div(K, splits);
  k_per = ceil_div(k_per, kTK) * kTK;
  dim3 grid(ceil_div(N, kTN), ceil_div(M, kTM), ceil_div(K, k_per));
  sgemm_kernel<<<grid, kGemmThreads, 0, stream>>>(A, sam, sak, B, sbk, sbn, Cm, ldc, bias, M, N, K, k_per, accumulate ? 1 : 0);
  C2S_LAUNCH_CHECK(name);
  return C2S_OK;
}

}  // namespace
}  // namespace c2s

extern "C" {

size_t c2s_ltae_mlp_backward_workspace_bytes(const c2s_ltae_desc* d) {
  if (d == nullptr || d->c_out <= 0) return 0;
  const size_t rows = static_cast<size_t>(d->B) * d->H * d->W;
  // y rows, g rows, column sums
  return (2 * c2s::align64(rows * d->c_out) + c2s::align64(2 * static_cast<size_t>(d->c_out))) * sizeof(float);
}

int c2s_ltae_mlp_backward(const c2s_ltae_desc* dp, const c2s_ltae_params* pp, const c2s_ltae_mlp_bwd_io* iop,
                          void* workspace, size_t workspace_bytes, void* stream_ptr) {
  using namespace c2s;
  C2S_CHECK_ARG(dp != nullptr && pp != nullptr && iop != nullptr, "c2s_ltae_mlp_backward: desc/params/io is NULL");
  const c2s_ltae_desc& d = *dp;
  const c2s_ltae_params& p = *pp;
  const c2s_ltae_mlp_bwd_io& io = *iop;
  C2S_CHECK_ARG(!(d.flags & C2S_LTAE_ATTN_ONLY), "c2s_ltae_mlp_backward: the attention-only encoder has no MLP");
  C2S_CHECK_ARG(d.B > 0 && d.H > 0 && d.W > 0 && d.d_model > 0 && d.c_out > 0 && d.n_head > 0,
                "c2s_ltae_mlp_backward: non-positive dimension");
  C2S_CHECK_ARG(d.dtype == C2S_F32 || d.dtype == C2S_BF16, "c2s_ltae_mlp_backward: unknown dtype %d", d.dtype);
  C2S_CHECK_ARG(d.c_out % d.n_head == 0, "c2s_ltae_mlp_backward: mlp[-1]=%d not divisible by n_head=%d", d.c_out, d.n_head);
  if (d.c_out / d.n_head > 16) C2S_UNSUPPORTED("c2s_ltae_mlp_backward: %d channels per out_norm group exceed the supported 16", d.c_out / d.n_head);
  C2S_CHECK_ARG(io.o_rows && io.grad_out && io.grad_o && io.bn_mean && io.bn_var && io.grad_mlp_weight &&
                    io.grad_mlp_bias && io.grad_bn_weight && io.grad_bn_bias && io.grad_out_norm_weight &&
                    io.grad_out_norm_bias,
                "c2s_ltae_mlp_backward: an io pointer is NULL");
  C2S_CHECK_ARG(p.mlp_weight && p.mlp_bias && p.bn_weight && p.bn_bias && p.out_norm_weight,
                "c2s_ltae_mlp_backward: mlp / out_norm parameters missing");
  C2S_CHECK_ARG(workspace != nullptr && workspace_bytes >= c2s_ltae_mlp_backward_workspace_bytes(dp),
                "c2s_ltae_mlp_backward: workspace too small");
  int status = check_device();
  if (status != C2S_OK) return status;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_ptr);
  const int hw = d.H * d.W, n_rows = d.B * hw, co = d.c_out, D = d.d_model;
  const bool train = (d.flags & C2S_LTAE_BN_BATCH_STATS) != 0;
  float* y = static_cast<float*>(workspace);
  float* g = y + align64(static_cast<size_t>(n_rows) * co);
  float* colsum = g + align64(static_cast<size_t>(n_rows) * co);
  C2S_CUDA(cudaMemsetAsync(colsum, 0, 2 * static_cast<size_t>(co) * sizeof(float), stream));

  // y = o Wm^T + bm: the rows the training-mode forward saved (params->save_y), else recomputed
  const float* y_in = io.y_rows;
  if (y_in == nullptr) {
    status = launch_gemm(io.o_rows, D, 1, p.mlp_weight, 1, D, y, co, p.mlp_bias, n_rows, co, D, 1, false, stream,
                         "ltae_mlp_backward<recompute y>");
    if (status != C2S_OK) return status;
    y_in = y;
  }
  RowsArgs a{};
  a.y = y_in, a.g_out = io.grad_out, a.mean = io.bn_mean, a.var = io.bn_var;
  a.bn_w = p.bn_weight, a.bn_b = p.bn_bias, a.on_w = p.out_norm_weight;
  a.keep = p.mlp_keep, a.keep_scale = d.mlp_keep_scale, a.bn_eps = d.bn_eps, a.gn_eps = d.gn_eps;
  a.g_yh = g, a.g_bn_w = io.grad_bn_weight, a.g_bn_b = io.grad_bn_bias;
  a.g_on_w = io.grad_out_norm_weight, a.g_on_b = io.grad_out_norm_bias;
  a.colsum = colsum, a.g_bm = io.grad_mlp_bias;
  a.B = d.B, a.hw = hw, a.c_out = co, a.cog = co / d.n_head, a.n_groups = d.n_head, a.train = train;
  const int n_warps = ceil_div(n_rows, 32) * d.n_head;
  if (d.dtype == C2S_BF16)
    mlp_rows_backward_kernel<__nv_bfloat16><<<ceil_div(n_warps, 8), 256, 0, stream>>>(a);
  else
    mlp_rows_backward_kernel<float><<<ceil_div(n_warps, 8), 256, 0, stream>>>(a);
  C2S_LAUNCH_CHECK("ltae_mlp_backward<rows>");
  if (train) {
    bn_batch_backward_kernel<<<ceil_div(n_rows, 32), 128, 0, stream>>>(g, y_in, io.bn_mean, io.bn_var, colsum, d.bn_eps,
                                                                     io.grad_mlp_bias, n_rows, co);
    C2S_LAUNCH_CHECK("ltae_mlp_backward<batchnorm>");
  }
  // grad_o = g Wm;  grad_Wm += g^T o (split over the rows)
  status = launch_gemm(g, co, 1, p.mlp_weight, D, 1, io.grad_o, D, nullptr, n_rows, D, co, 1, false, stream,
                       "ltae_mlp_backward<grad_o>");
  if (status != C2S_OK) return status;
  return launch_gemm(g, 1, co, io.o_rows, D, 1, io.grad_mlp_weight, D, nullptr, co, D, n_rows, ceil_div(n_rows, 128), true,
                     stream, "ltae_mlp_backward<grad_weight>");
}


int c2s_ltae_inconv_grad(const float* grad_o, const float* zn_rows, const float* sa_rows, float* grad_inconv_weight,
                         float* grad_inconv_bias, int64_t n_rows, int32_t n_head, int32_t d_model, int32_t C, void* stream_ptr) {
  using namespace c2s;
  C2S_CHECK_ARG(grad_o && zn_rows && grad_inconv_weight, "c2s_ltae_inconv_grad: grad_o / zn_rows / grad_inconv_weight is NULL");
  C2S_CHECK_ARG(grad_inconv_bias == nullptr || sa_rows != nullptr, "c2s_ltae_inconv_grad: the bias gradient needs sa_rows");
  C2S_CHECK_ARG(n_rows > 0 && n_head > 0 && n_head <= 16 && d_model % n_head == 0 && C > 0, "c2s_ltae_inconv_grad: bad shape");
  const int dh = d_model / n_head;
  if (dh > kHoMaxDh || C > kHoMaxC)
    C2S_UNSUPPORTED("c2s_ltae_inconv_grad: d_model / n_head = %d (max %d) or C = %d (max %d) not supported", dh, kHoMaxDh, C,
                    kHoThreads);
  int status = check_device();
  if (status != C2S_OK) return status;
  const int rows_per_cta = 128;
  dim3 grid(n_head, ceil_div(static_cast<int>(n_rows), rows_per_cta));
  // threads = channel slots (C / 2) x parts of the head's dh outputs: the fewest outputs per thread that fit 256 threads
  const int n_c = (C + 1) / 2;
  const int max_parts = kHoThreads / n_c > 0 ? kHoThreads / n_c : 1;
  const int per = ceil_div(dh, max_parts);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_ptr);
  const int rows = static_cast<int>(n_rows);
#define C2S_HO(PER) head_outer_kernel<PER><<<grid, kHoThreads, 0, stream>>>(grad_o, zn_rows, sa_rows, grad_inconv_weight, \
                                                                            grad_inconv_bias, rows, n_head, dh, C, rows_per_cta)
  if (per <= 2) C2S_HO(2);
  else if (per <= 4) C2S_HO(4);
  else if (per <= 8) C2S_HO(8);
  else if (per <= 16) C2S_HO(16);
  else C2S_HO(32);
#undef C2S_HO
  C2S_LAUNCH_CHECK("ltae_inconv_grad");
  return C2S_OK;
}

}  // extern "C"
