// sct_b200 — library-level plumbing behind the C ABI: thread-local error text, device queries and the
// host-side tensor-map (TMA descriptor) factory.  See include/sct_b200.h for the contract.
#include <map>
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <utility>

#include "../../include/sct_b200.h"
#include "common.cuh"

namespace sct {

static thread_local char tl_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}

static std::mutex g_once_mutex;  // guards every once-flag / cache below (backward runs on autograd worker threads)

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(g_once_mutex);
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int ensure_dyn_smem(const void* func, int bytes) {
  int dev = 0;
  SCT_CUDA(cudaGetDevice(&dev));
  static std::map<std::pair<const void*, int>, int> done;  // (kernel, device) -> bytes granted
  std::lock_guard<std::mutex> lock(g_once_mutex);
  auto key = std::make_pair(func, dev);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return 0;
  SCT_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done[key] = bytes;
  return 0;
}

int env_int(const char* name, int dflt) {
  static std::map<std::string, int> cache;
  std::lock_guard<std::mutex> lock(g_once_mutex);
  auto it = cache.find(name);
  if (it != cache.end()) return it->second;
  const char* e = getenv(name);
  const int v = (e != nullptr && e[0] != 0) ? atoi(e) : dflt;
  cache[name] = v;
  return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static CUtensorMapSwizzle to_cu_swizzle(Swz s) {
  switch (s) {
    case SWZ_32: return CU_TENSOR_MAP_SWIZZLE_32B;
    case SWZ_64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case SWZ_128: return CU_TENSOR_MAP_SWIZZLE_128B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, bool is_float, uint64_t inner,
                 uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                 Swz swz) {
  EncodeTiledFn fn = encode_fn();
  SCT_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  SCT_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor base %p not 16-byte aligned", base);
  SCT_CHECK((pitch_bytes & 15) == 0, "row pitch %llu not a multiple of 16 bytes",
            (unsigned long long)pitch_bytes);
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                           : (is_float ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                       : CU_TENSOR_MAP_DATA_TYPE_UINT32);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, to_cu_swizzle(swz), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCT_CHECK(r == CUDA_SUCCESS,
            "cuTensorMapEncodeTiled(2d) failed: %d (inner=%llu outer=%llu pitch=%llu box=%ux%u)", (int)r,
            (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes,
            box_inner, box_outer);
  return 0;
}

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1,
                 uint64_t d2, uint64_t pitch1_bytes, uint64_t pitch2_bytes, uint32_t box0,
                 uint32_t box1, Swz swz) {
  EncodeTiledFn fn = encode_fn();
  SCT_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  SCT_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor base %p not 16-byte aligned", base);
  SCT_CHECK((pitch1_bytes & 15) == 0 && (pitch2_bytes & 15) == 0, "pitches not multiples of 16 bytes");
  CUtensorMapDataType dt =
      elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {pitch1_bytes, pitch2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, to_cu_swizzle(swz), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed: %d (d=%llu,%llu,%llu box=%ux%u)",
            (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, box0, box1);
  return 0;
}

// per-translation-unit bounded-wait flags (see common.cuh)
int gemm_timeout_flag();
int attn_timeout_flag();

}  // namespace sct

extern "C" {

const char* sct_last_error(void) { return sct::tl_error; }

int32_t sct_version(void) { return SCT_B200_VERSION; }

int32_t sct_device_check(void) {
  int dev = 0;
  SCT_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  SCT_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SCT_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  SCT_CHECK(major == 10, "sct_b200 needs an sm_100-class device (B200); found sm_%d%d", major, minor);
  return 0;
}

int32_t sct_debug_timeouts(void) { return sct::gemm_timeout_flag() | (sct::attn_timeout_flag() << 1); }

}  // extern "C"
