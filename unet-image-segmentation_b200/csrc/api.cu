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

// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no -lcuda), behind a small per-thread cache: a tensor map is a pure
// function of (type, rank, base pointer, dims, strides, box, element strides, interleave, swizzle, L2 promotion, OOB fill), and an
// engine launches the same few hundred (buffer, shape) combinations every step, so eager launch sequences re-encode nothing after
// the first step.  Direct-mapped, 2048 entries per thread, full-key compare: a stale entry is impossible, a collision just
// re-encodes.  MEASURED (round 2, B200): no visible effect — eager 256x256 batch-8 inference runs 7091 img/s with the cache and
// 7148 without (CUDA-graph replay: 14785): the eager path is bound by the Python / ctypes call sequence (41 launches in 1.1 ms),
// not by the driver's encode; at 512x512 batch 64 eager launches already keep the GPU busy (54.1 ms eager vs 54.6 replayed).
// UNET_B200_TMAP_CACHE=0 bypasses it.
static PFN_encodeTiled g_real_encode = nullptr;

struct TmapKey {
  uint64_t w[24];
  bool operator==(const TmapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmapEntry { TmapKey key; CUtensorMap map; bool valid; };

static CUresult encode_tiled_cached(CUtensorMap* out, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* dims,
                                    const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr,
                                    CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
  if (rank == 0 || rank > 5) return g_real_encode(out, dt, rank, base, dims, strides, box, estr, il, sw, l2, oob);
  TmapKey k;
  memset(&k, 0, sizeof(k));
  k.w[0] = (uint64_t)dt | ((uint64_t)rank << 8) | ((uint64_t)il << 16) | ((uint64_t)sw << 24) | ((uint64_t)l2 << 32) | ((uint64_t)oob << 40);
  k.w[1] = (uint64_t)(uintptr_t)base;
  for (cuuint32_t i = 0; i < rank; ++i) {
    k.w[2 + i] = dims[i];
    k.w[12 + i] = ((uint64_t)box[i] << 32) | estr[i];
  }
  for (cuuint32_t i = 0; i + 1 < rank; ++i) k.w[7 + i] = strides[i];
  uint64_t h = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < 17; ++i) { h ^= k.w[i]; h *= 0xFF51AFD7ED558CCDull; h ^= h >> 29; }
  static thread_local TmapEntry* table = nullptr;
  constexpr int kEntries = 2048;
  if (!table) table = static_cast<TmapEntry*>(calloc(kEntries, sizeof(TmapEntry)));
  if (!table) return g_real_encode(out, dt, rank, base, dims, strides, box, estr, il, sw, l2, oob);
  TmapEntry& e = table[h & (kEntries - 1)];
  if (e.valid && e.key == k) { memcpy(out, &e.map, sizeof(CUtensorMap)); return CUDA_SUCCESS; }
  const CUresult r = g_real_encode(out, dt, rank, base, dims, strides, box, estr, il, sw, l2, oob);
  if (r == CUDA_SUCCESS) { e.key = k; memcpy(&e.map, out, sizeof(CUtensorMap)); e.valid = true; }
  return r;
}

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      g_real_encode = reinterpret_cast<PFN_encodeTiled>(ptr);
      const char* e = getenv("UNET_B200_TMAP_CACHE");
      fn = (e && e[0] == '0') ? g_real_encode : encode_tiled_cached;
    }
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
