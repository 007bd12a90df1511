// pad_mask[b,t] = (input[b,t] == pad_value).all()  --  the first statement of every model forward (utae.py:201-203,
// wtae.py:221-223, timeunet.py:170-172).  The reference materialises the
// full-size boolean comparison and reduces it three times; here one CTA owns one frame and leaves at the first
// vector that holds a value other than pad_value, so a valid frame costs a few kilobytes of reads and only padded
// frames are scanned to the end (the answer for them needs every element).  NaN != pad_value, as in torch.
#include "c2s_common.cuh"

namespace c2s {
namespace {

constexpr int kPadThreads = 256;

template <typename T>
__device__ __forceinline__ bool vec_differs(const uint4& v, float pad);
template <>
__device__ __forceinline__ bool vec_differs<float>(const uint4& v, float pad) {
  return __uint_as_float(v.x) != pad || __uint_as_float(v.y) != pad || __uint_as_float(v.z) != pad || __uint_as_float(v.w) != pad;
}
template <>
__device__ __forceinline__ bool vec_differs<__nv_bfloat16>(const uint4& v, float pad) {
  float f[8];
  Elem<__nv_bfloat16>::unpack(v, f);
  bool d = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) d = d || (f[i] != pad);
  return d;
}

template <typename T>
__global__ void __launch_bounds__(kPadThreads) pad_mask_kernel(const T* __restrict__ x, long long frame_elems, float pad,
                                                               uint8_t* __restrict__ mask, int vectorised) {
  constexpr int VEC = Elem<T>::kVec;
  const T* f = x + static_cast<size_t>(blockIdx.x) * frame_elems;
  bool differs = false;
  if (vectorised) {
    const long long n_vec = frame_elems / VEC;
    const uint4* v = reinterpret_cast<const uint4*>(f);
    // rounds of 4 vectors per thread (16 KB per CTA in flight); the CTA votes after every round and leaves early
    for (long long base = 0; base < n_vec; base += 4 * kPadThreads) {
      uint4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = base + u * kPadThreads + threadIdx.x;
        r[u] = i < n_vec ? ld_stream_v4(v + i) : make_uint4(0, 0, 0, 0);  // out of range: not examined below
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = base + u * kPadThreads + threadIdx.x;
        if (i < n_vec) differs = differs || vec_differs<T>(r[u], pad);
      }
      if (__syncthreads_or(differs)) {
        if (threadIdx.x == 0) mask[blockIdx.x] = 0;
        return;
      }
    }
    for (long long i = n_vec * VEC + threadIdx.x; i < frame_elems; i += kPadThreads) differs = differs || (Elem<T>::load(f + i) != pad);
  } else {
    for (long long base = 0; base < frame_elems; base += 8 * kPadThreads) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long i = base + u * kPadThreads + threadIdx.x;
        if (i < frame_elems) differs = differs || (Elem<T>::load(f + i) != pad);
      }
      if (__syncthreads_or(differs)) {
        if (threadIdx.x == 0) mask[blockIdx.x] = 0;
        return;
      }
    }
  }
  const int any = __syncthreads_or(differs);
  if (threadIdx.x == 0) mask[blockIdx.x] = any ? 0 : 1;
}

}  // namespace
}  // namespace c2s

extern "C" int c2s_pad_mask(const void* x, int32_t dtype, int64_t n_frames, int64_t frame_elems, float pad_value,
                            uint8_t* mask, void* stream_ptr) {
  using namespace c2s;
  C2S_CHECK_ARG(x != nullptr && mask != nullptr, "c2s_pad_mask: x/mask is NULL");
  C2S_CHECK_ARG(n_frames > 0 && frame_elems > 0, "c2s_pad_mask: non-positive size (%lld frames of %lld elements)",
                static_cast<long long>(n_frames), static_cast<long long>(frame_elems));
  C2S_CHECK_ARG(dtype == C2S_F32 || dtype == C2S_BF16, "c2s_pad_mask: unknown dtype %d", dtype);
  if (n_frames > 0x7fffffffll) C2S_UNSUPPORTED("c2s_pad_mask: more than 2^31 - 1 frames");
  int status = check_device();
  if (status != C2S_OK) return status;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_ptr);
  const size_t esize = dtype == C2S_BF16 ? 2 : 4;
  // whole 16-byte vectors need every frame to start on a 16-byte boundary
  const int vectorised = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && ((static_cast<size_t>(frame_elems) * esize) % 16 == 0);
  // torch compares a bf16 tensor with the scalar rounded to bf16 (`input == pad_value`, utae.py:201): do the same
  if (dtype == C2S_BF16) pad_value = __bfloat162float(__float2bfloat16_rn(pad_value));
  if (dtype == C2S_BF16)
    pad_mask_kernel<__nv_bfloat16><<<static_cast<unsigned>(n_frames), kPadThreads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x), frame_elems, pad_value, mask, vectorised);
  else
    pad_mask_kernel<float><<<static_cast<unsigned>(n_frames), kPadThreads, 0, stream>>>(static_cast<const float*>(x), frame_elems,
                                                                                         pad_value, mask, vectorised);
  C2S_LAUNCH_CHECK("pad_mask");
  return C2S_OK;
}
