// fairygen_b200 — host-side glue shared by the C-ABI translation units.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/fairygen_b200.h"

struct fgb_ctx {
  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  // cuTensorMapEncodeTiled, resolved through the runtime so the library does not link libcuda
  // (it must dlopen on a CPU-only box for the symbol/ABI tests).
  CUresult (*encode_tiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
  CUresult (*mem_get_address_range)(CUdeviceptr* base, size_t* size, CUdeviceptr ptr) = nullptr;
  // stream-K tail of the 2-CTA GEMM (gemm.cu: pair_plan_streamk); -1 = not initialised yet
  int sk_enabled = 1, sk_min_kblocks = -1;
  double sk_max_frac = 0.9;
  int32_t* attn_stats = nullptr;   // caller-owned device int32[3] (fgb_attn_set_stats) or NULL
  // Encoded tensor maps, keyed by (base, rows, cols, ld, box): a denoise step re-uses the same ~40 (pointer, shape) pairs for
  // its 600 GEMM / attention launches, so the driver's encoder runs once per pair instead of three times per launch.
  struct TmapSlot {
    const void* base = nullptr;
    int64_t rows = 0, cols = 0, ld = 0;
    int32_t box_rows = 0;
    CUtensorMap map;
  };
  static constexpr int kTmapSlots = 512;
  mutable TmapSlot tmap_cache[kTmapSlots];
};

#define FGB_MAX_PEERS 8  // one NVSwitch box

namespace fgb {

int set_error(int code, const char* fmt, ...);

#define FGB_CHECK_ARG(cond, ...)                                       \
  do {                                                                 \
    if (!(cond)) return ::fgb::set_error(FGB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define FGB_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::fgb::set_error(FGB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                                  \
  } while (0)

#define FGB_LAUNCH_CHECK(name)                                                                       \
  do {                                                                                               \
    cudaError_t _e = cudaGetLastError();                                                             \
    if (_e != cudaSuccess)                                                                           \
      return ::fgb::set_error(FGB_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with leading dimension `ld` elements,
// box = [box_rows, 64 cols] (= 128 bytes per row), 128-byte swizzle; out-of-bounds reads give zeros.
int make_tmap_bf16_2d(const fgb_ctx* ctx, CUtensorMap* map, const void* base, int64_t rows, int64_t cols,
                      int64_t ld, int32_t box_rows, int32_t box_cols = 64);

// cudaFuncSetAttribute applies to the current device only: remember per kernel which device ordinals of this process have
// been configured (one bit each), so that a process driving several GPUs configures every one of them.
inline bool first_use_on_device(unsigned long long& seen) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (seen & bit) return false;
  seen |= bit;
  return true;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace fgb
