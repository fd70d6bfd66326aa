// Host-side plumbing shared by the engine: error reporting, device buffers, TMA descriptor cache.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ptts.h"

namespace ptts {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline std::string fmt(const char* f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof buf, f, ap);
  va_end(ap);
  return buf;
}

#define PTTS_CUDA(expr)                                                                                    \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess)                                                                                 \
      throw ::ptts::Error(PTTS_ERR_CUDA, ::ptts::fmt("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                                     __FILE__, __LINE__));                                 \
  } while (0)

#define PTTS_REQUIRE(cond, code, ...)                                  \
  do {                                                                 \
    if (!(cond)) throw ::ptts::Error(code, ::ptts::fmt(__VA_ARGS__));  \
  } while (0)

// RAII device allocation (zero-initialised).
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return;
    PTTS_CUDA(cudaMalloc(&p, count * sizeof(T)));
    PTTS_CUDA(cudaMemset(p, 0, count * sizeof(T)));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void upload(const T* host, size_t count, cudaStream_t s = nullptr) {
    PTTS_REQUIRE(count <= n, PTTS_ERR_INVALID, "upload of %zu elements into buffer of %zu", count, n);
    PTTS_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
};

// cuTensorMapEncodeTiled resolved through the runtime so the library links against cudart only.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    PTTS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    PTTS_REQUIRE(p && q == cudaDriverEntryPointSuccess, PTTS_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// f16 tensor [d2][d1][d0] (d0 contiguous) viewed through a 3-D TMA descriptor with a 128-byte swizzled box
// {64, box1, box2}.  Descriptors are immutable once built, so they are cached by their full key.
struct TmapKey {
  const void* ptr;
  long long d0, d1, d2, s1, s2;
  int box1, box2;
  bool operator<(const TmapKey& o) const { return std::memcmp(this, &o, sizeof *this) < 0; }
};

class TmapCache {
 public:
  const CUtensorMap& get(const __half* ptr, long long d0, long long d1, long long d2, long long stride1_elems,
                         long long stride2_elems, int box1, int box2) {
    TmapKey k;
    std::memset(&k, 0, sizeof k);
    k.ptr = ptr; k.d0 = d0; k.d1 = d1; k.d2 = d2; k.s1 = stride1_elems; k.s2 = stride2_elems;
    k.box1 = box1; k.box2 = box2;
    auto it = maps_.find(k);
    if (it != maps_.end()) return it->second;
    PTTS_REQUIRE(d0 % 64 == 0, PTTS_ERR_INVALID, "TMA inner dimension %lld not a multiple of 64", d0);
    PTTS_REQUIRE(box1 >= 1 && box1 <= 256 && box2 >= 1 && box2 <= 256, PTTS_ERR_INVALID, "bad TMA box %d x %d", box1, box2);
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)stride1_elems * 2, (cuuint64_t)stride2_elems * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box1, (cuuint32_t)box2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(ptr), dims, strides, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PTTS_REQUIRE(r == CUDA_SUCCESS, PTTS_ERR_CUDA,
                 "cuTensorMapEncodeTiled failed (%d) dims %lld,%lld,%lld strides %lld,%lld box 64,%d,%d", (int)r, d0, d1,
                 d2, stride1_elems, stride2_elems, box1, box2);
    return maps_.emplace(k, m).first->second;
  }

  // int8 tensor [d1][d0] (d0 contiguous) through a dense, unswizzled box {64 bytes, box1 rows}: the raw weight-code tile
  // that the GEMM's converter warps expand to f16 in shared memory.
  const CUtensorMap& get_u8(const int8_t* ptr, long long d0, long long d1, int box1) {
    TmapKey k;
    std::memset(&k, 0, sizeof k);
    k.ptr = ptr; k.d0 = d0; k.d1 = d1; k.d2 = 1; k.s1 = d0; k.s2 = d0 * d1;
    k.box1 = box1; k.box2 = -8;  // distinguishes the element type
    auto it = maps_.find(k);
    if (it != maps_.end()) return it->second;
    PTTS_REQUIRE(d0 % 64 == 0 && box1 >= 1 && box1 <= 256, PTTS_ERR_INVALID, "bad int8 TMA geometry %lld x %lld box %d", d0, d1, box1);
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, 1};
    cuuint64_t strides[2] = {(cuuint64_t)d0, (cuuint64_t)(d0 * d1)};
    cuuint32_t box[3] = {64, (cuuint32_t)box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(ptr), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PTTS_REQUIRE(r == CUDA_SUCCESS, PTTS_ERR_CUDA, "cuTensorMapEncodeTiled (u8) failed (%d) dims %lld,%lld box 64,%d", (int)r, d0, d1, box1);
    return maps_.emplace(k, m).first->second;
  }

 private:
  std::map<TmapKey, CUtensorMap> maps_;
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int pow2_at_least(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace ptts
