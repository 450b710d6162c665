// K5 — scoring: loader-contract mask decode and the mIoU confusion matrix.
// Replaces `(y*255).long()` (+ `y[y==255]=0` on the bank side), hbird_eval.py:219,309-310, and
// PredsmIoU.update's bincount(gt*P+pred), eval_metrics.py:73-109.  Integer work, HBM-bound:
// 2 bytes per pixel.  Each thread run-length-aggregates 16 consecutive pixels in registers, adds
// into a per-block shared-memory histogram (uint32), and blocks flush once into the int64 matrix.
#include "common.cuh"

namespace hb {

__global__ void __launch_bounds__(256)
decode_mask_kernel(const float* __restrict__ y, int64_t n, int remap, uint8_t* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    // (y * 255).long(): fp32 multiply, truncate toward zero
    long long v = static_cast<long long>(__fmul_rn(y[i], 255.0f));
    if (remap && v == 255) v = 0;
    out[i] = static_cast<uint8_t>(v);
  }
}

constexpr int kConfThreads = 256;
constexpr int kConfPixPerThread = 16;

__device__ __forceinline__ void conf_flush(uint32_t* hist, int bin, uint32_t cnt) {
  if (bin >= 0 && cnt) atomicAdd(hist + bin, cnt);
}

__global__ void __launch_bounds__(kConfThreads)
confusion_kernel(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ pred, int64_t n, int Cg,
                 int Cp, int ignore_index, unsigned long long* __restrict__ conf) {
  extern __shared__ uint32_t s_hist[];  // Cg * Cp
  const int bins = Cg * Cp;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();

  const int64_t chunk = static_cast<int64_t>(kConfThreads) * kConfPixPerThread;
  const bool aligned = ((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * chunk; base < n;
       base += static_cast<int64_t>(gridDim.x) * chunk) {
    const int64_t i0 = base + static_cast<int64_t>(threadIdx.x) * kConfPixPerThread;
    if (i0 >= n) continue;
    __align__(16) uint8_t g[kConfPixPerThread];
    __align__(16) uint8_t p[kConfPixPerThread];
    int m = kConfPixPerThread;
    if (aligned && i0 + kConfPixPerThread <= n) {
      *reinterpret_cast<uint4*>(g) = __ldg(reinterpret_cast<const uint4*>(gt + i0));
      *reinterpret_cast<uint4*>(p) = __ldg(reinterpret_cast<const uint4*>(pred + i0));
    } else {
      m = static_cast<int>(n - i0 < kConfPixPerThread ? n - i0 : kConfPixPerThread);
      for (int j = 0; j < kConfPixPerThread; ++j) {
        g[j] = j < m ? gt[i0 + j] : 0;
        p[j] = j < m ? pred[i0 + j] : 0;
      }
    }
    int run_bin = -1;
    uint32_t run_cnt = 0;
#pragma unroll
    for (int j = 0; j < kConfPixPerThread; ++j) {
      const int gv = g[j], pv = p[j];
      const bool ok = j < m && gv != ignore_index && gv < Cg && pv < Cp;
      const int bin = ok ? gv * Cp + pv : -1;
      if (bin != run_bin) {
        conf_flush(s_hist, run_bin, run_cnt);
        run_bin = bin;
        run_cnt = 0;
      }
      run_cnt++;
    }
    conf_flush(s_hist, run_bin, run_cnt);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(conf + i, static_cast<unsigned long long>(c));
  }
}

}  // namespace hb

extern "C" {

int hb_decode_mask(const float* y_dev, int64_t n, int remap_255_to_0, uint8_t* out_dev, void* stream) {
  HB_REQUIRE(n >= 0, "hb_decode_mask: n < 0");
  if (n == 0) return HB_OK;
  HB_REQUIRE(y_dev && out_dev, "hb_decode_mask: NULL pointer");
  int64_t blocks = hb::ceil_div64(n, 256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  hb::decode_mask_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(y_dev, n, remap_255_to_0, out_dev);
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

int hb_confusion_accumulate(const uint8_t* gt_dev, const uint8_t* pred_dev, int64_t n, int C_gt, int C_pred,
                            int ignore_index, int64_t* conf_dev, void* stream) {
  HB_REQUIRE(C_gt >= 1 && C_gt <= 256 && C_pred >= 1 && C_pred <= 256, "hb_confusion_accumulate: class counts (%d, %d) not in [1, 256]", C_gt, C_pred);
  HB_REQUIRE(C_gt * C_pred * 4 <= 200 * 1024, "hb_confusion_accumulate: %d x %d bins exceed shared memory", C_gt, C_pred);
  HB_REQUIRE(n >= 0, "hb_confusion_accumulate: n < 0");
  if (n == 0) return HB_OK;
  HB_REQUIRE(gt_dev && pred_dev && conf_dev, "hb_confusion_accumulate: NULL pointer");
  const size_t smem = sizeof(uint32_t) * C_gt * C_pred;
  if (smem > 48 * 1024)
    HB_CHECK_CUDA(cudaFuncSetAttribute(hb::confusion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int64_t chunk = static_cast<int64_t>(hb::kConfThreads) * hb::kConfPixPerThread;
  int64_t blocks = hb::ceil_div64(n, chunk);
  // blocks per SM are bounded by the histogram's shared-memory footprint
  const int per_sm = smem <= 8 * 1024 ? 8 : (smem <= 48 * 1024 ? 4 : 2);
  if (blocks > 148 * per_sm) blocks = 148 * per_sm;
  hb::confusion_kernel<<<static_cast<unsigned>(blocks), hb::kConfThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      gt_dev, pred_dev, n, C_gt, C_pred, ignore_index, reinterpret_cast<unsigned long long*>(conf_dev));
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

}  // extern "C"
