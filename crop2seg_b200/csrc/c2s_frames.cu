// Lengths-aware frame packing for blocks that are shared across the sequence:
// TemporallySharedBlock.smart_forward, src/backbones/temp_shared_block.py:18-47.  The reference boolean-indexes the
// [B*T, C, H, W] batch (`out[~pad_mask]`, a nonzero() with a host synchronisation) and scatters the block's output
// into a pad_value-filled tensor (`temp[~pad_mask] = ...`).  Here:
//   c2s_frame_index    pad_mask[n] -> slot[n] (position of frame f among the valid frames, -1 for padded ones) and the
//                      number of valid frames, one CTA, exclusive scan on the device;
//   c2s_frames_gather  packed[slot[f]] = frames[f] for every valid frame;
//   c2s_frames_scatter out[f] = slot[f] >= 0 ? packed[slot[f]] : pad_value   (every output byte written exactly once:
//                      no fill pass followed by an indexed overwrite).
// Pure byte movement: HBM-bound, 16-byte vectors when the frames allow it.
#include "c2s_common.cuh"

namespace c2s {
namespace {

constexpr int kScanThreads = 1024;
constexpr int kCopyThreads = 256;

__global__ void __launch_bounds__(kScanThreads) frame_index_kernel(const uint8_t* __restrict__ pad, int n,
                                                                   int32_t* __restrict__ slot, int32_t* __restrict__ count) {
  __shared__ int warp_sums[kScanThreads / 32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += kScanThreads) {
    const int i = base + threadIdx.x;
    const int valid = (i < n && pad[i] == 0) ? 1 : 0;
    int incl = valid;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      warp_sums[lane] = w;  // inclusive over the warps
    }
    __syncthreads();
    const int before = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + incl - valid;
    if (i < n) slot[i] = valid ? before : -1;
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_sums[kScanThreads / 32 - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

// blockIdx.y = frame, blockIdx.x = piece of the frame; VEC: 16-byte or 1-byte units
template <typename V, bool SCATTER>
__global__ void __launch_bounds__(kCopyThreads) frames_move_kernel(const V* __restrict__ src, V* __restrict__ dst,
                                                                   const int32_t* __restrict__ slot, long long units,
                                                                   V fill) {
  const int f = blockIdx.y;
  const int s = __ldg(slot + f);
  if (!SCATTER && s < 0) return;  // gather: padded frames are not read
  const V* from = SCATTER ? (s >= 0 ? src + static_cast<size_t>(s) * units : nullptr) : src + static_cast<size_t>(f) * units;
  V* to = SCATTER ? dst + static_cast<size_t>(f) * units : dst + static_cast<size_t>(s) * units;
  for (long long i = blockIdx.x * static_cast<long long>(kCopyThreads) + threadIdx.x; i < units;
       i += static_cast<long long>(gridDim.x) * kCopyThreads)
    to[i] = from != nullptr ? from[i] : fill;
}

int move_frames(bool scatter, const void* src, void* dst, const int32_t* slot, int64_t n_frames, int64_t frame_bytes,
                int32_t elem_bytes, uint32_t fill_bits, cudaStream_t stream, const char* name) {
  const bool vec = frame_bytes % 16 == 0 && reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0;
  const long long units = vec ? frame_bytes / 16 : frame_bytes / elem_bytes;
  long long bx = (units + kCopyThreads - 1) / kCopyThreads;
  if (bx > 64) bx = 64;  // a frame is walked by up to 64 CTAs; the grid's second dimension is the frame
  dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(n_frames));
  if (vec) {
    const uint32_t w = elem_bytes == 2 ? (fill_bits & 0xffffu) * 0x10001u : fill_bits;
    const uint4 fill = make_uint4(w, w, w, w);
    if (scatter) frames_move_kernel<uint4, true><<<grid, kCopyThreads, 0, stream>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), slot, units, fill);
    else frames_move_kernel<uint4, false><<<grid, kCopyThreads, 0, stream>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), slot, units, fill);
  } else if (elem_bytes == 4) {
    if (scatter) frames_move_kernel<uint32_t, true><<<grid, kCopyThreads, 0, stream>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst), slot, units, fill_bits);
    else frames_move_kernel<uint32_t, false><<<grid, kCopyThreads, 0, stream>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst), slot, units, fill_bits);
  } else {
    const uint16_t f16 = static_cast<uint16_t>(fill_bits & 0xffffu);
    if (scatter) frames_move_kernel<uint16_t, true><<<grid, kCopyThreads, 0, stream>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), slot, units, f16);
    else frames_move_kernel<uint16_t, false><<<grid, kCopyThreads, 0, stream>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), slot, units, f16);
  }
  C2S_LAUNCH_CHECK(name);
  return C2S_OK;
}

int check_frames(const void* a, const void* b, const int32_t* slot, int64_t n_frames, int64_t frame_elems, int32_t dtype,
                 const char* who) {
  C2S_CHECK_ARG(a != nullptr && b != nullptr && slot != nullptr, "%s: NULL pointer", who);
  C2S_CHECK_ARG(n_frames > 0 && frame_elems > 0, "%s: non-positive size (%lld frames of %lld elements)", who,
                static_cast<long long>(n_frames), static_cast<long long>(frame_elems));
  C2S_CHECK_ARG(dtype == C2S_F32 || dtype == C2S_BF16, "%s: unknown dtype %d", who, dtype);
  if (n_frames > 65535) C2S_UNSUPPORTED("%s: more than 65535 frames in one call", who);
  return check_device();
}

}  // namespace
}  // namespace c2s

extern "C" int c2s_frame_index(const uint8_t* pad_mask, int64_t n_frames, int32_t* slot, int32_t* n_valid, void* stream_ptr) {
  using namespace c2s;
  C2S_CHECK_ARG(pad_mask != nullptr && slot != nullptr && n_valid != nullptr, "c2s_frame_index: NULL pointer");
  C2S_CHECK_ARG(n_frames > 0, "c2s_frame_index: %lld frames", static_cast<long long>(n_frames));
  if (n_frames > 0x7fffffffll) C2S_UNSUPPORTED("c2s_frame_index: more than 2^31 - 1 frames");
  int status = check_device();
  if (status != C2S_OK) return status;
  frame_index_kernel<<<1, kScanThreads, 0, static_cast<cudaStream_t>(stream_ptr)>>>(pad_mask, static_cast<int>(n_frames), slot, n_valid);
  C2S_LAUNCH_CHECK("frame_index");
  return C2S_OK;
}

extern "C" int c2s_frames_gather(const void* frames, const int32_t* slot, void* packed, int64_t n_frames, int64_t frame_elems,
                                 int32_t dtype, void* stream_ptr) {
  using namespace c2s;
  int status = check_frames(frames, packed, slot, n_frames, frame_elems, dtype, "c2s_frames_gather");
  if (status != C2S_OK) return status;
  const int es = dtype == C2S_BF16 ? 2 : 4;
  return move_frames(false, frames, packed, slot, n_frames, frame_elems * es, es, 0u, static_cast<cudaStream_t>(stream_ptr), "frames_gather");
}

extern "C" int c2s_frames_scatter(const void* packed, const int32_t* slot, void* out, int64_t n_frames, int64_t frame_elems,
                                  int32_t dtype, float pad_value, void* stream_ptr) {
  using namespace c2s;
  int status = check_frames(packed, out, slot, n_frames, frame_elems, dtype, "c2s_frames_scatter");
  if (status != C2S_OK) return status;
  const int es = dtype == C2S_BF16 ? 2 : 4;
  uint32_t bits;
  if (dtype == C2S_BF16) {
    const __nv_bfloat16 h = __float2bfloat16_rn(pad_value);
    bits = *reinterpret_cast<const uint16_t*>(&h);
  } else {
    bits = *reinterpret_cast<const uint32_t*>(&pad_value);
  }
  return move_frames(true, packed, out, slot, n_frames, frame_elems * es, es, bits, static_cast<cudaStream_t>(stream_ptr), "frames_scatter");
}
