// fairygen_b200 — C-ABI context, error reporting and TMA descriptor construction.
#include <cstring>
#include <new>

#include "host.h"

namespace fgb {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int make_tmap_bf16_2d(const fgb_ctx* ctx, CUtensorMap* map, const void* base, int64_t rows, int64_t cols,
                      int64_t ld, int32_t box_rows, int32_t box_cols) {
  if (!ctx || !ctx->encode_tiled) return set_error(FGB_ERR_INVALID, "context has no tensor-map encoder");
  if (!aligned16(base)) return set_error(FGB_ERR_INVALID, "matrix base pointer %p is not 16-byte aligned", base);
  if ((ld * 2) % 16 != 0) return set_error(FGB_ERR_INVALID, "leading dimension %lld elements is not a multiple of 8", (long long)ld);
  if (rows <= 0 || cols <= 0) return set_error(FGB_ERR_INVALID, "empty matrix %lld x %lld", (long long)rows, (long long)cols);
  if (box_rows < 1 || box_rows > 256 || box_cols != 64) return set_error(FGB_ERR_INVALID, "bad TMA box %d x %d", box_rows, box_cols);
  // direct-mapped cache: the map depends on nothing but these five values (a tensor map holds an address, not data)
  uint64_t hkey = reinterpret_cast<uint64_t>(base) * 0x9E3779B97F4A7C15ull;
  hkey ^= (static_cast<uint64_t>(rows) * 0xC2B2AE3D27D4EB4Full) ^ (static_cast<uint64_t>(cols) << 20) ^ (static_cast<uint64_t>(ld) << 40) ^
          static_cast<uint64_t>(box_rows);
  fgb_ctx::TmapSlot& slot = ctx->tmap_cache[(hkey >> 32) % fgb_ctx::kTmapSlots];
  if (slot.base == base && slot.rows == rows && slot.cols == cols && slot.ld == ld && slot.box_rows == box_rows) {
    *map = slot.map;
    return FGB_OK;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = ctx->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                                 box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(FGB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld box=%dx%d)",
                     (int)r, (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
  slot.base = base;
  slot.rows = rows;
  slot.cols = cols;
  slot.ld = ld;
  slot.box_rows = box_rows;
  slot.map = *map;
  return FGB_OK;
}

}  // namespace fgb

extern "C" {

int fgb_abi_version(void) { return FGB_ABI_VERSION; }

const char* fgb_last_error(void) { return fgb::g_err; }

int fgb_create(int device, fgb_ctx** out) {
  if (!out) return fgb::set_error(FGB_ERR_INVALID, "fgb_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fgb::set_error(FGB_ERR_CUDA, "fgb_create: no CUDA device (%s); this library has no CPU fallback",
                          e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= count) return fgb::set_error(FGB_ERR_INVALID, "fgb_create: device %d out of range [0,%d)", device, count);
  FGB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FGB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fgb::set_error(FGB_ERR_UNSUPPORTED, "fgb_create: device %d is sm_%d%d; kernels are built for sm_100a only", device,
                          prop.major, prop.minor);
  fgb_ctx* ctx = new (std::nothrow) fgb_ctx();
  if (!ctx) return fgb::set_error(FGB_ERR_INVALID, "fgb_create: out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    delete ctx;
    return fgb::set_error(FGB_ERR_CUDA, "fgb_create: cuTensorMapEncodeTiled not available from the driver");
  }
  ctx->encode_tiled = reinterpret_cast<decltype(ctx->encode_tiled)>(fn);
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn)
    ctx->mem_get_address_range = reinterpret_cast<decltype(ctx->mem_get_address_range)>(fn);
  else
    cudaGetLastError();
  *out = ctx;
  return FGB_OK;
}

int fgb_destroy(fgb_ctx* ctx) {
  delete ctx;
  return FGB_OK;
}

int fgb_sync_check(fgb_ctx* ctx, void* stream) {
  if (!ctx) return fgb::set_error(FGB_ERR_INVALID, "fgb_sync_check: ctx is NULL");
  FGB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  FGB_CUDA(cudaGetLastError());
  return FGB_OK;
}

int fgb_sm_count(fgb_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

// ---- peer memory over NVLink: CUDA IPC handles for caller-owned buffers (one process per GPU) ----------------------
int fgb_ipc_export(fgb_ctx* ctx, const void* dev_ptr, void* handle_out, int64_t* offset_out) {
  if (!ctx || !dev_ptr || !handle_out || !offset_out) return fgb::set_error(FGB_ERR_INVALID, "fgb_ipc_export: NULL argument");
  if (!ctx->mem_get_address_range) return fgb::set_error(FGB_ERR_UNSUPPORTED, "fgb_ipc_export: cuMemGetAddressRange unavailable");
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = ctx->mem_get_address_range(&base, &size, reinterpret_cast<CUdeviceptr>(dev_ptr));
  if (r != CUDA_SUCCESS) return fgb::set_error(FGB_ERR_CUDA, "fgb_ipc_export: cuMemGetAddressRange failed (%d)", (int)r);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  FGB_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(dev_ptr) - base);
  return FGB_OK;
}

int fgb_ipc_open(fgb_ctx* ctx, const void* handle, int64_t offset, void** peer_ptr) {
  if (!ctx || !handle || !peer_ptr || offset < 0) return fgb::set_error(FGB_ERR_INVALID, "fgb_ipc_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* base = nullptr;
  FGB_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *peer_ptr = static_cast<char*>(base) + offset;
  return FGB_OK;
}

int fgb_ipc_close(fgb_ctx* ctx, void* peer_ptr, int64_t offset) {
  if (!ctx || !peer_ptr) return fgb::set_error(FGB_ERR_INVALID, "fgb_ipc_close: bad argument");
  FGB_CUDA(cudaIpcCloseMemHandle(static_cast<char*>(peer_ptr) - offset));
  return FGB_OK;
}

}  // extern "C"
