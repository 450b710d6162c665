// Library plumbing: error strings, device check, TMA tensor-map encoding, scratch.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace hb {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void clear_error() { g_err[0] = 0; }

int ensure_workspace(Bank* b, size_t bytes, cudaStream_t st) {
  if (bytes <= b->ws_bytes) return HB_OK;
  // Stream-ordered growth: the old block is released after the work already queued on `st` (the
  // only stream a bank is driven from, include/hbird_b200.h) and nothing synchronises the device.
  if (b->ws) {
    HB_CHECK_CUDA(cudaFreeAsync(b->ws, st));
    b->ws = nullptr;
    b->ws_bytes = 0;
  }
  size_t want = bytes + bytes / 4;
  HB_CHECK_CUDA(cudaMallocAsync(&b->ws, want, st));
  b->ws_bytes = want;
  return HB_OK;
}

int device_sm_count() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, int64_t rows, int cols_pad,
                      int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return HB_ERR_UNSUPPORTED;
  }
  if (rows < 1) rows = 1;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols_pad), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols_pad) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                  gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%d box_rows=%d)",
              static_cast<int>(r), static_cast<long long>(rows), cols_pad, box_rows);
    return HB_ERR_CUDA;
  }
  return HB_OK;
}

}  // namespace hb

extern "C" {

int hb_abi_version(void) { return HB_ABI_VERSION; }

const char* hb_last_error(void) { return hb::g_err; }

int hb_device_check(int device, int* num_sms_out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n < 1) {
    hb::set_error("no CUDA device available (%s); hbird_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    (void)cudaGetLastError();
    return HB_ERR_UNSUPPORTED;
  }
  if (device < 0 || device >= n) {
    hb::set_error("invalid GPU id %d (available: 0-%d)", device, n - 1);
    return HB_ERR_INVALID;
  }
  cudaDeviceProp prop;
  HB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    hb::set_error("device %d is sm_%d%d; hbird_b200 kernels are built for sm_100a only", device,
                  prop.major, prop.minor);
    return HB_ERR_UNSUPPORTED;
  }
  if (num_sms_out) *num_sms_out = prop.multiProcessorCount;
  return HB_OK;
}

}  // extern "C"
