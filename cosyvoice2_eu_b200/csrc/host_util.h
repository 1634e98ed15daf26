// Host-side helpers: error plumbing for the C-ABI and TMA tensor-map encoding
// (cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint, so the library
// does not link against libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <stdexcept>
#include <string>

namespace cv2 {

// thread-local last error for cv2_last_error()
inline std::string& last_error_ref() {
  static thread_local std::string e;
  return e;
}

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

[[noreturn]] inline void fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error(buf);
}

#define CV2_CUDA(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) ::cv2::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define CV2_CHECK(cond, ...)                 \
  do {                                       \
    if (!(cond)) ::cv2::fail(__VA_ARGS__);   \
  } while (0)

#define CV2_LAUNCH_CHECK() CV2_CUDA(cudaGetLastError())

// One-time setup that is PER DEVICE (cudaFuncSetAttribute, __constant__ uploads): a process may hold engines on several GPUs
// (get_engine("cuda:1")), and a process-wide `static bool configured` would leave every device but the first unconfigured.
// run(f) calls f() the first time it is reached with a given current device; thread safe.
struct PerDeviceOnce {
  std::mutex m;
  unsigned long long done = 0;   // bit per device ordinal (<= 64 devices)
  template <class F>
  void run(F&& f) {
    int dev = 0;
    CV2_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> g(m);
    if ((done >> (dev & 63)) & 1ull) return;
    f();
    done |= 1ull << (dev & 63);
  }
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CV2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    CV2_CHECK(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled entry point not available");
    fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// 16-bit element tensor map, 128B swizzle, zero OOB fill.  dims[0] is the contiguous dimension.
// strides_bytes[i] is the byte stride of dims[i+1] (rank-1 entries).
inline CUtensorMap make_tmap_16b(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                 const uint32_t* box) {
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; i++) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CV2_CHECK(((uintptr_t)base & 15) == 0, "TMA base %p not 16B aligned", base);
  for (int i = 0; i + 1 < rank; i++) CV2_CHECK((gstr[i] & 15) == 0, "TMA stride %llu not multiple of 16", (unsigned long long)gstr[i]);
  CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                                  es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CV2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
            (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
  return m;
}

}  // namespace cv2
