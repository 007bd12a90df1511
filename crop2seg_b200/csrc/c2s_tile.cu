// Edges of the tile inference pipeline, sm_100a.
//
//   c2s_tile_patchify   raw tile [T, C, H, W] (int16 / uint16 / float32 reflectances) -> model input patches
//                       [P, T_pad, C, 128, 128] (float32 or bfloat16): zero-pad the RAW tile to whole patches
//                       (src/helpers/dataset_creator.py:385-388), reorder the channels (src/datasets/s2_ts_cz_crop.py:248,
//                       374), normalise (d - mean[c]) / std[c] in fp32 (s2_ts_cz_crop.py:393-398), pad_value on the frames
//                       behind T (pad_collate, src/utils.py:14-33).  The reference does this on the host, patch by patch,
//                       and ships 4-byte floats over PCIe; here the raw 2-byte tile crosses once.
//   c2s_tile_classmap   logits [P, K, 128, 128] -> class map [H, W] uint8 (+ probabilities [K, H, W] float32): softmax
//                       over the classes, FIRST maximum, patches put back row-major and cropped to the tile
//                       (src/webapp/prediction.py:316-333, which copies every patch to the host first).
// Both are HBM-bound element-wise kernels: coalesced rows of 128 pixels, one pass, nothing staged.
#include <cfloat>

#include "c2s_common.cuh"

namespace c2s {
namespace {

constexpr int kTileThreads = 256;
constexpr int kPatchMax = 128;

struct TileArgs {
  const void* tile;
  const int32_t* order;
  const float* mean;
  const float* stdv;
  void* patches;
  int T, T_pad, C, H, W, patch, grid_w, patch_begin, patch_count;
  float pad_value;
};

// 8 consecutive raw values as floats: 8-byte loads when the address allows it (a 10980-pixel int16 row starts on an
// 8-byte boundary, a 16-byte one only every other row), scalar loads at ragged edges
template <typename S>
__device__ __forceinline__ void load_raw8(const S* src, int n_inside, float (&raw)[8]) {
  if (n_inside == 8 && (reinterpret_cast<uintptr_t>(src) & 7u) == 0) {
    constexpr int CH = sizeof(S);  // uint2 chunks: 2 (16-bit) or 4 (float)
    uint2 w[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) w[k] = __ldg(reinterpret_cast<const uint2*>(src) + k);
    const S* v = reinterpret_cast<const S*>(w);
#pragma unroll
    for (int k = 0; k < 8; ++k) raw[k] = static_cast<float>(v[k]);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) raw[k] = k < n_inside ? static_cast<float>(__ldg(src + k)) : 0.f;  // np.pad(..., 'constant')
  }
}

// one CTA = one (patch, frame, channel) plane of patch x patch pixels; one thread = 8 consecutive pixels of a row
template <typename S, typename D>
__global__ void __launch_bounds__(kTileThreads) tile_patchify_kernel(const TileArgs a) {
  const int plane = blockIdx.x;
  const int c = plane % a.C;
  const int t = (plane / a.C) % a.T_pad;
  const int p = plane / (a.C * a.T_pad);
  const int vec_per_row = a.patch / 8;
  const int n_vec = a.patch * vec_per_row;
  D* out = static_cast<D*>(a.patches) + static_cast<size_t>(plane) * a.patch * a.patch;
  const bool behind = t >= a.T;  // pad_collate: frames behind the series hold pad_value (after the normalisation)
  const int pid = a.patch_begin + p;
  const int Y0 = (pid / a.grid_w) * a.patch, X00 = (pid % a.grid_w) * a.patch;
  const float m = behind ? 0.f : __ldg(a.mean + c), s = behind ? 1.f : __ldg(a.stdv + c);
  const S* plane_src = behind ? nullptr
                              : static_cast<const S*>(a.tile) + (static_cast<size_t>(t) * a.C + __ldg(a.order + c)) * a.H * a.W;
  for (int i = threadIdx.x; i < n_vec; i += kTileThreads) {
    const int y = i / vec_per_row, xv = i - y * vec_per_row;
    float v[8];
    if (behind) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = a.pad_value;
    } else {
      const int Y = Y0 + y, X0 = X00 + xv * 8;
      const int inside = Y < a.H ? min(8, max(0, a.W - X0)) : 0;
      float raw[8];
      if (inside > 0) {
        load_raw8(plane_src + static_cast<size_t>(Y) * a.W + X0, inside, raw);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) raw[k] = 0.f;  // zero-padding of the RAW tile (dataset_creator.py:385-388)
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __fdiv_rn(__fsub_rn(raw[k], m), s);  // exactly (d - mean) / std in fp32
    }
    D* dst = out + static_cast<size_t>(i) * 8;
    if constexpr (sizeof(D) == 2) {
      st_stream_v4(dst, Elem<__nv_bfloat16>::pack(v));
    } else {
      const float lo[4] = {v[0], v[1], v[2], v[3]}, hi[4] = {v[4], v[5], v[6], v[7]};
      st_stream_v4(dst, Elem<float>::pack(lo));
      st_stream_v4(dst + 4, Elem<float>::pack(hi));
    }
  }
}

struct ClassArgs {
  const void* logits;
  uint8_t* classmap;
  float* proba;
  int K, H, W, patch, grid_w, patch_begin, patch_count;
};

// one thread = one pixel of a patch (consecutive threads = consecutive pixels: every class plane is read coalesced)
template <typename D, int KMAX>
__global__ void __launch_bounds__(kTileThreads) tile_classmap_kernel(const ClassArgs a) {
  const int pp = a.patch * a.patch;
  const D* logits = static_cast<const D*>(a.logits);
  const int p = blockIdx.y;  // one patch per grid row
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < pp; q += gridDim.x * blockDim.x) {
    const int pid = a.patch_begin + p;
    const int Y = (pid / a.grid_w) * a.patch + q / a.patch, X = (pid % a.grid_w) * a.patch + q % a.patch;
    if (Y >= a.H || X >= a.W) continue;  // cropped away (prediction.py:332-333)
    const D* src = logits + static_cast<size_t>(p) * a.K * pp + q;
    float e[KMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < a.K) {
        e[k] = Elem<D>::load(src + static_cast<size_t>(k) * pp);
        mx = fmaxf(mx, e[k]);
      }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < a.K) {
        e[k] = expf(__fsub_rn(e[k], mx));  // torch.nn.Softmax(dim=1): exp(x - max) / sum  (prediction.py:318)
        sum = __fadd_rn(sum, e[k]);
      }
    int best = 0;
    float pbest = -1.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < a.K) {
        const float pr = __fdiv_rn(e[k], sum);
        if (pr > pbest) pbest = pr, best = k;  // strict: the FIRST maximum wins, like pred_.max(dim=1)[1] (prediction.py:320)
        if (a.proba != nullptr) a.proba[(static_cast<size_t>(k) * a.H + Y) * a.W + X] = pr;
      }
    a.classmap[static_cast<size_t>(Y) * a.W + X] = static_cast<uint8_t>(best);
  }
}

int check_tile_desc(const c2s_tile_desc* d, const char* who) {
  C2S_CHECK_ARG(d != nullptr, "%s: desc is NULL", who);
  C2S_CHECK_ARG(d->H > 0 && d->W > 0 && d->patch > 0 && d->patch % 8 == 0 && d->patch <= 1024,
                "%s: bad tile %d x %d / patch %d (patch must be a multiple of 8)", who, d->H, d->W, d->patch);
  const int gh = d->grid_h > 0 ? d->grid_h : ceil_div(d->H, d->patch), gw = d->grid_w > 0 ? d->grid_w : ceil_div(d->W, d->patch);
  C2S_CHECK_ARG(static_cast<long long>(gh) * d->patch >= d->H && static_cast<long long>(gw) * d->patch >= d->W,
                "%s: the %d x %d patch grid does not cover the tile", who, gh, gw);
  C2S_CHECK_ARG(d->patch_begin >= 0 && d->patch_count > 0 &&
                    static_cast<long long>(d->patch_begin) + d->patch_count <= static_cast<long long>(gh) * gw,
                "%s: patches [%d, %d) outside the %d x %d grid", who, d->patch_begin, d->patch_begin + d->patch_count, gh, gw);
  return C2S_OK;
}

}  // namespace
}  // namespace c2s

extern "C" int c2s_tile_patchify(const c2s_tile_desc* d, const void* tile, const int32_t* channels_order, const float* mean,
                                 const float* std, void* patches, void* stream_ptr) {
  using namespace c2s;
  int status = check_tile_desc(d, "c2s_tile_patchify");
  if (status != C2S_OK) return status;
  C2S_CHECK_ARG(tile && channels_order && mean && std && patches, "c2s_tile_patchify: NULL pointer");
  C2S_CHECK_ARG(d->T > 0 && d->T_pad >= d->T && d->C > 0, "c2s_tile_patchify: bad T=%d / T_pad=%d / C=%d", d->T, d->T_pad, d->C);
  C2S_CHECK_ARG(d->src_dtype >= C2S_RAW_I16 && d->src_dtype <= C2S_RAW_F32, "c2s_tile_patchify: unknown src_dtype %d", d->src_dtype);
  C2S_CHECK_ARG(d->dst_dtype == C2S_F32 || d->dst_dtype == C2S_BF16, "c2s_tile_patchify: unknown dst_dtype %d", d->dst_dtype);
  C2S_CHECK_ARG(reinterpret_cast<uintptr_t>(patches) % 16 == 0, "c2s_tile_patchify: patches must be 16-byte aligned");
  status = check_device();
  if (status != C2S_OK) return status;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_ptr);
  TileArgs a{};
  a.tile = tile, a.order = channels_order, a.mean = mean, a.stdv = std, a.patches = patches;
  a.T = d->T, a.T_pad = d->T_pad, a.C = d->C, a.H = d->H, a.W = d->W, a.patch = d->patch;
  a.grid_w = d->grid_w > 0 ? d->grid_w : ceil_div(d->W, d->patch);
  a.patch_begin = d->patch_begin, a.patch_count = d->patch_count, a.pad_value = d->pad_value;
  const long long planes = static_cast<long long>(a.patch_count) * a.T_pad * a.C;
  if (planes > 0x7fffffffll) C2S_UNSUPPORTED("c2s_tile_patchify: more than 2^31 - 1 (patch, frame, channel) planes in one call");
  const unsigned grid = static_cast<unsigned>(planes);
  const bool bf = d->dst_dtype == C2S_BF16;
#define C2S_PATCHIFY(S)                                                                              \
  do {                                                                                               \
    if (bf) tile_patchify_kernel<S, __nv_bfloat16><<<grid, kTileThreads, 0, stream>>>(a);            \
    else tile_patchify_kernel<S, float><<<grid, kTileThreads, 0, stream>>>(a);                       \
  } while (0)
  if (d->src_dtype == C2S_RAW_I16) C2S_PATCHIFY(int16_t);
  else if (d->src_dtype == C2S_RAW_U16) C2S_PATCHIFY(uint16_t);
  else C2S_PATCHIFY(float);
#undef C2S_PATCHIFY
  C2S_LAUNCH_CHECK("tile_patchify");
  return C2S_OK;
}

extern "C" int c2s_tile_classmap(const c2s_tile_desc* d, const void* logits, int32_t n_classes, uint8_t* classmap,
                                 float* proba, void* stream_ptr) {
  using namespace c2s;
  int status = check_tile_desc(d, "c2s_tile_classmap");
  if (status != C2S_OK) return status;
  C2S_CHECK_ARG(logits && classmap, "c2s_tile_classmap: NULL pointer");
  C2S_CHECK_ARG(n_classes > 0 && n_classes <= 256, "c2s_tile_classmap: %d classes (the class map is uint8)", n_classes);
  C2S_CHECK_ARG(d->dst_dtype == C2S_F32 || d->dst_dtype == C2S_BF16, "c2s_tile_classmap: unknown logits dtype %d", d->dst_dtype);
  if (n_classes > 32) C2S_UNSUPPORTED("c2s_tile_classmap: more than 32 classes (%d)", n_classes);
  status = check_device();
  if (status != C2S_OK) return status;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_ptr);
  ClassArgs a{};
  a.logits = logits, a.classmap = classmap, a.proba = proba;
  a.K = n_classes, a.H = d->H, a.W = d->W, a.patch = d->patch;
  a.grid_w = d->grid_w > 0 ? d->grid_w : ceil_div(d->W, d->patch);
  a.patch_begin = d->patch_begin, a.patch_count = d->patch_count;
  if (a.patch_count > 65535) C2S_UNSUPPORTED("c2s_tile_classmap: more than 65535 patches in one call");
  const dim3 grid(static_cast<unsigned>(ceil_div(a.patch * a.patch, kTileThreads)), static_cast<unsigned>(a.patch_count));
  const bool bf = d->dst_dtype == C2S_BF16;
  if (n_classes <= 16) {
    if (bf) tile_classmap_kernel<__nv_bfloat16, 16><<<grid, kTileThreads, 0, stream>>>(a);
    else tile_classmap_kernel<float, 16><<<grid, kTileThreads, 0, stream>>>(a);
  } else {
    if (bf) tile_classmap_kernel<__nv_bfloat16, 32><<<grid, kTileThreads, 0, stream>>>(a);
    else tile_classmap_kernel<float, 32><<<grid, kTileThreads, 0, stream>>>(a);
  }
  C2S_LAUNCH_CHECK("tile_classmap");
  return C2S_OK;
}
