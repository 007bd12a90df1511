// Library services of the crop2seg_b200 C ABI: error text, launch accounting, ABI version.
#include <atomic>
#include <cstring>

#include "c2s_common.cuh"

namespace c2s {

namespace {
thread_local char g_error[512] = "";
// process-wide (autograd runs backward kernels on its own thread); diagnostics only
char g_kernel[128] = "";
char g_ltae_kernel[128] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void note_launch(const char* kernel_name) {
  ++g_launches;
  strncpy(g_kernel, kernel_name, sizeof(g_kernel) - 1);
  g_kernel[sizeof(g_kernel) - 1] = '\0';
  if (strncmp(kernel_name, "ltae_forward", 12) == 0) {
    strncpy(g_ltae_kernel, kernel_name, sizeof(g_ltae_kernel) - 1);
    g_ltae_kernel[sizeof(g_ltae_kernel) - 1] = '\0';
  }
}

int check_device() {
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) {
    set_error("cudaGetDevice failed: %s", cudaGetErrorString(err));
    return C2S_ERR_NO_DEVICE;
  }
  // one attribute query per device, cached (re-entrant: worst case two threads write the same value)
  static int cached_major[64] = {0};
  if (dev < 64 && cached_major[dev] == 0) {
    int major = 0;
    err = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (err != cudaSuccess) {
      set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(err));
      return C2S_ERR_NO_DEVICE;
    }
    cached_major[dev] = major;
  }
  if (dev < 64 && cached_major[dev] != 10) {
    set_error("crop2seg_b200 is built for sm_100a only; device %d has compute capability major %d",
              dev, cached_major[dev]);
    return C2S_ERR_NO_DEVICE;
  }
  return C2S_OK;
}

}  // namespace c2s

extern "C" {

int c2s_abi_version(void) { return C2S_ABI_VERSION; }

const char* c2s_last_error(void) { return c2s::g_error; }

int64_t c2s_launch_count(void) { return c2s::g_launches.load(); }

void c2s_reset_launch_count(void) { c2s::g_launches.store(0); }

const char* c2s_last_kernel(void) { return c2s::g_kernel; }
const char* c2s_last_ltae_kernel(void) { return c2s::g_ltae_kernel; }

}  // extern "C"
