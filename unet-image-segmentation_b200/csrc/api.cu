// Library-level entry points and error plumbing of libunet_b200.so.
#include "common.cuh"
#include "ptx.cuh"
#include <string.h>
#include <stdlib.h>

namespace unet {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", where, (int)e, cudaGetErrorString(e));
  return (int)e;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("UNET_B200_PDL");
    cached = (e && e[0] == '1') ? 1 : 0;     // opt-in: measured slower under CUDA-graph replay (common.cuh)
  }
  return cached == 1;
}

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

}  // namespace unet

using namespace unet;

extern "C" int unet_version(void) { return 100; }   // 0.1.0 -> major*10000 + minor*100 + patch
extern "C" int unet_sm_arch(void) { return 100; }
extern "C" const char* unet_last_error(void) { return g_err; }

extern "C" int unet_device_check(int device) {
  int major = 0, minor = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) return set_cuda_error(e, "device_check");
  e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (e != cudaSuccess) return set_cuda_error(e, "device_check");
  UNET_REQUIRE(major == 10, UNET_EUNSUPPORTED, "device %d is sm_%d%d; libunet_b200 contains sm_100a code only", device, major, minor);
  return UNET_OK;
}

extern "C" uint32_t unet_host_dropout_hash(uint64_t idx, uint32_t seed) { return dropout_hash(idx, seed); }
