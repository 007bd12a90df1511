// Loss side of the training step, sm_100a.
//
//   c2s_boundary_target   y_b = where(get_dilated(y, K, device, 4).sum(1) > 1, 1, 0)   src/learning/utils.py:198-222, 283-285
//                         The reference one-hot encodes the labels (B x K x H x W floats), runs a grouped 3x3 convolution
//                         and sums over the classes; more than one class in the (zero-padded) cross-shaped neighbourhood
//                         means: some in-bounds neighbour differs from the centre.  One pass over the int64 labels.
//   c2s_seg_loss_forward  kind 0: nn.CrossEntropyLoss(weight, label_smoothing) on [B, K, H, W] scores   train.py:462-467
//                         kind 1: FocalCELoss(gamma, size_average, ignore_index, weight)              src/learning/focal_loss.py
//   c2s_seg_loss_backward gradient with respect to the scores (what autograd derives from the two modules).
// One thread per pixel, the K class planes are read coalesced; deterministic two-stage reduction (per-CTA partial sums in
// the workspace, one CTA adds them in a fixed order in double precision).  HBM-bound: scores read once per direction.
#include <cfloat>

#include "c2s_common.cuh"

namespace c2s {
namespace {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 1024;
constexpr int kLossMaxClasses = 64;

struct LossArgs {
  const void* logits;
  const long long* target;
  const float* weight;  // [K] or nullptr
  int K, hw;
  long long n_pix;      // B * H * W
  int kind, ignore_index, size_average;
  float gamma, smoothing;
};

template <typename D>
__device__ __forceinline__ float load_score(const D* p) { return Elem<D>::load(p); }

// log-softmax pieces of one pixel: m = max, lse = log sum exp(z - m)
template <typename D>
__device__ __forceinline__ void pixel_lse(const D* z, int K, int hw, float& m, float& lse) {
  m = -INFINITY;
  for (int k = 0; k < K; ++k) m = fmaxf(m, load_score(z + static_cast<size_t>(k) * hw));
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += expf(load_score(z + static_cast<size_t>(k) * hw) - m);
  lse = logf(s);
}

__device__ __forceinline__ bool target_kept(const LossArgs& a, long long y) {
  // focal_loss.py:24 drops target == ignore_index; nn.CrossEntropyLoss drops its ignore_index (-100, train.py passes none)
  const long long ign = a.kind == 1 ? a.ignore_index : -100;
  return y != ign && y >= 0 && y < a.K;
}

template <typename D>
__global__ void __launch_bounds__(kLossThreads) loss_forward_kernel(const LossArgs a, double* __restrict__ partial) {
  __shared__ double red[3][kLossThreads / 32];
  double num = 0.0, den = 0.0, cnt = 0.0;
  float wtot = 0.f;
  if (a.kind == 0 && a.smoothing > 0.f)
    for (int k = 0; k < a.K; ++k) wtot += a.weight != nullptr ? __ldg(a.weight + k) : 1.f;
  for (long long i = blockIdx.x * static_cast<long long>(kLossThreads) + threadIdx.x; i < a.n_pix;
       i += static_cast<long long>(gridDim.x) * kLossThreads) {
    const long long y = a.target[i];
    if (!target_kept(a, y)) continue;
    const long long b = i / a.hw;
    const D* z = static_cast<const D*>(a.logits) + (static_cast<size_t>(b) * a.K) * a.hw + (i - b * a.hw);
    float m, lse;
    pixel_lse(z, a.K, a.hw, m, lse);
    const float wy = a.weight != nullptr ? __ldg(a.weight + y) : 1.f;
    const float logpt = load_score(z + static_cast<size_t>(y) * a.hw) - m - lse;
    if (a.kind == 0) {
      float l = -wy * logpt;
      if (a.smoothing > 0.f) {  // (1 - eps) w_y nll + eps / K sum_k w_k (-log p_k), both over sum_i w_y (ATen)
        float sm = 0.f;
        for (int k = 0; k < a.K; ++k) {
          const float wk = a.weight != nullptr ? __ldg(a.weight + k) : 1.f;
          sm -= wk * (load_score(z + static_cast<size_t>(k) * a.hw) - m - lse);
        }
        l = (1.f - a.smoothing) * l + a.smoothing / static_cast<float>(a.K) * sm;
      }
      num += l, den += wy;
    } else {
      const float pt = expf(logpt);
      // focal_loss.py:36-38 AS WRITTEN: the gathered class weights keep their [N, 1] shape and broadcast against the
      // [N] focal terms, so a weighted loss is the outer product (sum_i w[y_i]) x (sum_j focal_j) -- the two factors
      // are summed separately here
      num += -powf(1.f - pt, a.gamma) * logpt;
      den += wy, cnt += 1.0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = num, red[1][threadIdx.x >> 5] = den, red[2][threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    double n = 0.0, d = 0.0, c = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) n += red[0][w], d += red[1][w], c += red[2][w];
    partial[blockIdx.x] = n, partial[kLossMaxBlocks + blockIdx.x] = d, partial[2 * kLossMaxBlocks + blockIdx.x] = c;
  }
}

// sums[0] = numerator, sums[1] = weight sum, sums[2] = the factor between the numerator and the loss (the backward
// scales the per-pixel gradients with it), sums[3] = kept pixels
__global__ void loss_finish_kernel(const double* __restrict__ partial, int n_blocks, int kind, int size_average, int weighted,
                                   float* __restrict__ loss, float* __restrict__ sums) {
  if (threadIdx.x != 0) return;
  double n = 0.0, d = 0.0, c = 0.0;
  for (int i = 0; i < n_blocks; ++i) n += partial[i], d += partial[kLossMaxBlocks + i], c += partial[2 * kLossMaxBlocks + i];
  double mult;
  if (kind == C2S_LOSS_CROSS_ENTROPY) mult = 1.0 / d;  // weighted mean; 0 / 0 = nan like torch
  else if (!weighted) mult = size_average ? 1.0 / c : 1.0;
  else mult = size_average ? d / (c * c) : d;           // mean / sum of the N x N outer product
  *loss = static_cast<float>(n * mult);
  sums[0] = static_cast<float>(n), sums[1] = static_cast<float>(d), sums[2] = static_cast<float>(mult), sums[3] = static_cast<float>(c);
}

template <typename D>
__global__ void __launch_bounds__(kLossThreads) loss_backward_kernel(const LossArgs a, const float* __restrict__ sums,
                                                                     const float* __restrict__ grad_loss, D* __restrict__ grad) {
  const float scale = __ldg(grad_loss) * __ldg(sums + 2);
  float wtot = 0.f;
  if (a.kind == 0 && a.smoothing > 0.f)
    for (int k = 0; k < a.K; ++k) wtot += a.weight != nullptr ? __ldg(a.weight + k) : 1.f;
  for (long long i = blockIdx.x * static_cast<long long>(kLossThreads) + threadIdx.x; i < a.n_pix;
       i += static_cast<long long>(gridDim.x) * kLossThreads) {
    const long long y = a.target[i];
    const long long b = i / a.hw;
    const size_t off = (static_cast<size_t>(b) * a.K) * a.hw + (i - b * a.hw);
    D* g = grad + off;
    if (!target_kept(a, y)) {
      for (int k = 0; k < a.K; ++k) Elem<D>::store(g + static_cast<size_t>(k) * a.hw, 0.f);
      continue;
    }
    const D* z = static_cast<const D*>(a.logits) + off;
    float m, lse;
    pixel_lse(z, a.K, a.hw, m, lse);
    const float wy = a.weight != nullptr ? __ldg(a.weight + y) : 1.f;
    float c_soft, c_hot;  // grad_k = scale * (c_soft * p_k - c_hot * [k == y] - c_w * w_k)
    float c_w = 0.f;
    if (a.kind == 0) {
      c_soft = (1.f - a.smoothing) * wy + a.smoothing / static_cast<float>(a.K) * wtot;
      c_hot = (1.f - a.smoothing) * wy;
      c_w = a.smoothing / static_cast<float>(a.K);
    } else {
      const float logpt = load_score(z + static_cast<size_t>(y) * a.hw) - m - lse;
      const float pt = expf(logpt), q = 1.f - pt;
      // d/dlogpt of -(1 - pt)^gamma w logpt = -w [ (1 - pt)^gamma - gamma (1 - pt)^(gamma - 1) pt logpt ]
      const float tail = a.gamma == 0.f ? 0.f : a.gamma * powf(q, a.gamma - 1.f) * pt * logpt;
      const float dl = -(powf(q, a.gamma) - tail);  // the class weights sit in the common factor (see the forward)
      c_soft = -dl, c_hot = -dl;
    }
    for (int k = 0; k < a.K; ++k) {
      const float p = expf(load_score(z + static_cast<size_t>(k) * a.hw) - m - lse);
      float v = c_soft * p - (k == y ? c_hot : 0.f);
      if (c_w != 0.f) v -= c_w * (a.weight != nullptr ? __ldg(a.weight + k) : 1.f);
      Elem<D>::store(g + static_cast<size_t>(k) * a.hw, scale * v);
    }
  }
}

__global__ void __launch_bounds__(kLossThreads) boundary_target_kernel(const long long* __restrict__ y, int H, int W,
                                                                       long long n, int diagonal, long long* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(kLossThreads) + threadIdx.x;
  if (i >= n) return;
  const int x = static_cast<int>(i % W), r = static_cast<int>((i / W) % H);
  const long long c = y[i];
  bool differs = false;
  // zero padding of the one-hot planes: pixels outside the image add no class (F.conv2d(..., padding=(1, 1)))
  if (r > 0) differs |= y[i - W] != c;
  if (r + 1 < H) differs |= y[i + W] != c;
  if (x > 0) differs |= y[i - 1] != c;
  if (x + 1 < W) differs |= y[i + 1] != c;
  if (diagonal) {  // connectivity 8: the full 3 x 3 window
    if (r > 0 && x > 0) differs |= y[i - W - 1] != c;
    if (r > 0 && x + 1 < W) differs |= y[i - W + 1] != c;
    if (r + 1 < H && x > 0) differs |= y[i + W - 1] != c;
    if (r + 1 < H && x + 1 < W) differs |= y[i + W + 1] != c;
  }
  out[i] = differs ? 1 : 0;
}

int check_loss(const c2s_loss_desc* d, const void* logits, const int64_t* target, const char* who) {
  C2S_CHECK_ARG(d != nullptr && logits != nullptr && target != nullptr, "%s: NULL pointer", who);
  C2S_CHECK_ARG(d->B > 0 && d->K > 0 && d->H > 0 && d->W > 0, "%s: non-positive dimension in scores[%d,%d,%d,%d]", who, d->B,
                d->K, d->H, d->W);
  C2S_CHECK_ARG(d->dtype == C2S_F32 || d->dtype == C2S_BF16, "%s: unknown dtype %d", who, d->dtype);
  C2S_CHECK_ARG(d->kind == C2S_LOSS_CROSS_ENTROPY || d->kind == C2S_LOSS_FOCAL, "%s: unknown loss kind %d", who, d->kind);
  C2S_CHECK_ARG(d->label_smoothing >= 0.f && d->label_smoothing <= 1.f, "%s: label_smoothing %g outside [0, 1]", who,
                static_cast<double>(d->label_smoothing));
  if (d->K > kLossMaxClasses) C2S_UNSUPPORTED("%s: %d classes (at most %d)", who, d->K, kLossMaxClasses);
  if (d->kind == C2S_LOSS_FOCAL && d->label_smoothing != 0.f) C2S_UNSUPPORTED("%s: FocalCELoss has no label smoothing", who);
  return check_device();
}

LossArgs loss_args(const c2s_loss_desc* d, const void* logits, const int64_t* target, const float* weight) {
  LossArgs a{};
  a.logits = logits, a.target = reinterpret_cast<const long long*>(target), a.weight = weight;
  a.K = d->K, a.hw = d->H * d->W, a.n_pix = static_cast<long long>(d->B) * d->H * d->W;
  a.kind = d->kind, a.ignore_index = d->ignore_index, a.size_average = d->size_average;
  a.gamma = d->gamma, a.smoothing = d->label_smoothing;
  return a;
}

int loss_grid(long long n_pix) {
  const long long blocks = (n_pix + kLossThreads - 1) / kLossThreads;
  return static_cast<int>(blocks < kLossMaxBlocks ? blocks : kLossMaxBlocks);
}

}  // namespace
}  // namespace c2s

extern "C" size_t c2s_seg_loss_workspace_bytes(void) { return 3 * c2s::kLossMaxBlocks * sizeof(double) + 4 * sizeof(float); }

extern "C" int c2s_seg_loss_forward(const c2s_loss_desc* d, const void* logits, const int64_t* target, const float* weight,
                                    float* loss, void* workspace, size_t workspace_bytes, void* stream_ptr) {
  using namespace c2s;
  int status = check_loss(d, logits, target, "c2s_seg_loss_forward");
  if (status != C2S_OK) return status;
  C2S_CHECK_ARG(loss != nullptr, "c2s_seg_loss_forward: loss is NULL");
  C2S_CHECK_ARG(workspace != nullptr && workspace_bytes >= c2s_seg_loss_workspace_bytes() &&
                    reinterpret_cast<uintptr_t>(workspace) % 8 == 0,
                "c2s_seg_loss_forward: needs %zu bytes of 8-byte aligned workspace", c2s_seg_loss_workspace_bytes());
  cudaStream_t stream = static_cast<cudaStream_t>(stream_ptr);
  const LossArgs a = loss_args(d, logits, target, weight);
  double* partial = static_cast<double*>(workspace);
  float* sums = reinterpret_cast<float*>(partial + 3 * kLossMaxBlocks);
  const int grid = loss_grid(a.n_pix);
  if (d->dtype == C2S_BF16) loss_forward_kernel<__nv_bfloat16><<<grid, kLossThreads, 0, stream>>>(a, partial);
  else loss_forward_kernel<float><<<grid, kLossThreads, 0, stream>>>(a, partial);
  C2S_LAUNCH_CHECK("seg_loss_forward");
  loss_finish_kernel<<<1, 32, 0, stream>>>(partial, grid, d->kind, d->size_average, weight != nullptr, loss, sums);
  C2S_LAUNCH_CHECK("seg_loss_finish");
  return C2S_OK;
}

extern "C" int c2s_seg_loss_backward(const c2s_loss_desc* d, const void* logits, const int64_t* target, const float* weight,
                                     const void* workspace, const float* grad_loss, void* grad_logits, void* stream_ptr) {
  using namespace c2s;
  int status = check_loss(d, logits, target, "c2s_seg_loss_backward");
  if (status != C2S_OK) return status;
  C2S_CHECK_ARG(workspace != nullptr && grad_loss != nullptr && grad_logits != nullptr, "c2s_seg_loss_backward: NULL pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_ptr);
  const LossArgs a = loss_args(d, logits, target, weight);
  const float* sums = reinterpret_cast<const float*>(static_cast<const double*>(workspace) + 3 * kLossMaxBlocks);
  const long long blocks = (a.n_pix + kLossThreads - 1) / kLossThreads;
  const int grid = static_cast<int>(blocks < 148 * 16 ? blocks : 148 * 16);
  if (d->dtype == C2S_BF16)
    loss_backward_kernel<__nv_bfloat16><<<grid, kLossThreads, 0, stream>>>(a, sums, grad_loss, static_cast<__nv_bfloat16*>(grad_logits));
  else
    loss_backward_kernel<float><<<grid, kLossThreads, 0, stream>>>(a, sums, grad_loss, static_cast<float*>(grad_logits));
  C2S_LAUNCH_CHECK("seg_loss_backward");
  return C2S_OK;
}

extern "C" int c2s_boundary_target(const int64_t* target, int32_t B, int32_t H, int32_t W, int32_t connectivity,
                                   int64_t* boundary, void* stream_ptr) {
  using namespace c2s;
  C2S_CHECK_ARG(target != nullptr && boundary != nullptr, "c2s_boundary_target: NULL pointer");
  C2S_CHECK_ARG(B > 0 && H > 0 && W > 0, "c2s_boundary_target: non-positive size [%d,%d,%d]", B, H, W);
  int status = check_device();
  if (status != C2S_OK) return status;
  const long long n = static_cast<long long>(B) * H * W;
  const long long blocks = (n + kLossThreads - 1) / kLossThreads;
  if (blocks > 0x7fffffffll) C2S_UNSUPPORTED("c2s_boundary_target: too many pixels");
  boundary_target_kernel<<<static_cast<unsigned>(blocks), kLossThreads, 0, static_cast<cudaStream_t>(stream_ptr)>>>(
      reinterpret_cast<const long long*>(target), H, W, n, connectivity == 8 ? 1 : 0, reinterpret_cast<long long*>(boundary));
  C2S_LAUNCH_CHECK("boundary_target");
  return C2S_OK;
}
