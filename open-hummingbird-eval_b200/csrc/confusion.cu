// K5 — scoring: loader-contract mask decode and the mIoU confusion matrix.
// Replaces `(y*255).long()` (+ `y[y==255]=0` on the bank side), hbird_eval.py:219,309-310, and
// PredsmIoU.update's bincount(gt*P+pred), eval_metrics.py:73-109.  Integer work, HBM-bound:
// 2 bytes per pixel.  Each thread finds the runs of equal (gt, pred) pairs among 16 consecutive
// pixels, adds one count per run into a per-block shared-memory histogram (uint32), and blocks
// flush once into the int64 matrix.
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

// Segmentation maps are piecewise constant, so the work is organised around RUNS, not pixels:
// a thread packs (gt, pred) byte pairs into 16-bit symbols, finds the run ends of each group of
// 16 pixels with a handful of word-wide XORs, and loops over its runs only.  With GROUPS > 1 a
// thread owns GROUPS*16 consecutive pixels and carries the open run from group to group, so long
// runs cost one shared-memory add per thread instead of one per 16 pixels (same-address adds from
// the 32 lanes of a warp serialise, which is what bounds this kernel on smooth maps).
template <int GROUPS>
__global__ void __launch_bounds__(kConfThreads)
confusion_kernel(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ pred, int64_t n, int Cg,
                 int Cp, int ignore_index, unsigned long long* __restrict__ conf) {
  extern __shared__ uint32_t s_hist[];  // Cg * Cp
  constexpr int kPix = GROUPS * kConfPixPerThread;
  const int bins = Cg * Cp;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();

  const int64_t chunk = static_cast<int64_t>(kConfThreads) * kPix;
  const bool aligned = ((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * chunk; base < n;
       base += static_cast<int64_t>(gridDim.x) * chunk) {
    const int64_t i0 = base + static_cast<int64_t>(threadIdx.x) * kPix;
    if (i0 >= n) continue;
    if (aligned && i0 + kPix <= n) {
      uint4 g[GROUPS], p[GROUPS];
#pragma unroll
      for (int u = 0; u < GROUPS; ++u) {
        g[u] = __ldg(reinterpret_cast<const uint4*>(gt + i0) + u);
        p[u] = __ldg(reinterpret_cast<const uint4*>(pred + i0) + u);
      }
      uint32_t open_sym = 0xffffffffu, open_cnt = 0;  // the run still open at the group boundary
#pragma unroll
      for (int u = 0; u < GROUPS; ++u) {
        // symbol j = gt_j << 8 | pred_j, two per word
        uint32_t w[9];
        w[0] = __byte_perm(g[u].x, p[u].x, 0x1504); w[1] = __byte_perm(g[u].x, p[u].x, 0x3726);
        w[2] = __byte_perm(g[u].y, p[u].y, 0x1504); w[3] = __byte_perm(g[u].y, p[u].y, 0x3726);
        w[4] = __byte_perm(g[u].z, p[u].z, 0x1504); w[5] = __byte_perm(g[u].z, p[u].z, 0x3726);
        w[6] = __byte_perm(g[u].w, p[u].w, 0x1504); w[7] = __byte_perm(g[u].w, p[u].w, 0x3726);
        w[8] = 0;
        uint32_t ends = 0x8000u;  // symbol 15 always ends a run (of this group)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // x halves: s[2k] ^ s[2k+1], s[2k+1] ^ s[2k+2]
          const uint32_t x = w[k] ^ __funnelshift_r(w[k], w[k + 1], 16);
          ends |= ((x & 0xffffu) ? 1u : 0u) << (2 * k);
          if (k < 7) ends |= ((x >> 16) ? 1u : 0u) << (2 * k + 1);
        }
        const uint64_t q0 = (static_cast<uint64_t>(w[1]) << 32) | w[0], q1 = (static_cast<uint64_t>(w[3]) << 32) | w[2];
        const uint64_t q2 = (static_cast<uint64_t>(w[5]) << 32) | w[4], q3 = (static_cast<uint64_t>(w[7]) << 32) | w[6];
        int prev = -1;
        while (ends) {
          const int j = __ffs(ends) - 1;
          ends &= ends - 1;
          const uint64_t lo = (j & 4) ? q1 : q0, hi = (j & 4) ? q3 : q2;
          const uint32_t sym = static_cast<uint32_t>(((j & 8) ? hi : lo) >> ((j & 3) * 16)) & 0xffffu;
          const uint32_t cnt = static_cast<uint32_t>(j - prev);
          prev = j;
          if (sym == open_sym) {
            open_cnt += cnt;
          } else {
            const int gv = open_sym >> 8, pv = open_sym & 255;  // open_sym = ~0: gv fails `< Cg`
            if (gv != ignore_index && gv < Cg && pv < Cp) atomicAdd(s_hist + gv * Cp + pv, open_cnt);
            open_sym = sym;
            open_cnt = cnt;
          }
        }
      }
      const int gv = open_sym >> 8, pv = open_sym & 255;
      if (gv != ignore_index && gv < Cg && pv < Cp) atomicAdd(s_hist + gv * Cp + pv, open_cnt);
    } else {  // ragged tail / unaligned views: plain per-pixel run-length pass
      const int m = static_cast<int>(n - i0 < kPix ? n - i0 : kPix);
      int run_bin = -1;
      uint32_t run_cnt = 0;
      for (int j = 0; j < m; ++j) {
        const int gv = gt[i0 + j], pv = pred[i0 + j];
        const bool ok = gv != ignore_index && gv < Cg && pv < Cp;
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
  const int64_t cap = static_cast<int64_t>(hb::device_sm_count()) * 8;
  if (blocks > cap) blocks = cap;
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
  // blocks per SM are bounded by the histogram's shared-memory footprint
  const int per_sm = smem <= 8 * 1024 ? 8 : (smem <= 48 * 1024 ? 4 : 2);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* conf = reinterpret_cast<unsigned long long*>(conf_dev);
  // 64 pixels per thread once there is enough work to fill the GPU that way, 16 otherwise
  const int sms = hb::device_sm_count();
  const bool wide = n >= static_cast<int64_t>(sms) * per_sm * hb::kConfThreads * 64;
  auto kernel = wide ? hb::confusion_kernel<4> : hb::confusion_kernel<1>;
  if (smem > 48 * 1024)
    HB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int64_t chunk = static_cast<int64_t>(hb::kConfThreads) * hb::kConfPixPerThread * (wide ? 4 : 1);
  int64_t blocks = hb::ceil_div64(n, chunk);
  if (blocks > static_cast<int64_t>(sms) * per_sm) blocks = static_cast<int64_t>(sms) * per_sm;
  kernel<<<static_cast<unsigned>(blocks), hb::kConfThreads, smem, st>>>(gt_dev, pred_dev, n, C_gt, C_pred,
                                                                         ignore_index, conf);
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

}  // extern "C"
