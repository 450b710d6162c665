// K4 — soft label transfer and fused bilinear-upsample + argmax.
// K4a replaces the neighbour gather (hbird_eval.py:632-636) and _cross_attention
// (hbird_eval.py:575-609).  Because bank rows are unit norm, cos(q, m_j) = score_j / ||q||, so the
// 2 GB/batch gather of neighbour FEATURES the reference performs is not needed: only the k label
// records (uint16 class histograms) are gathered.  HBM-bound: k*(4+8) + k*2C + 4C bytes per query.
// K4b replaces hbird_eval.py:235-243.  HBM-bound: 4*S*S*C bytes read + H*W bytes written per image;
// the (B, C, H, W) fp32 intermediate of the reference is never materialised.
#include "common.cuh"

namespace hb {

constexpr int kMaxK = 128;

// One warp per query.
__global__ void __launch_bounds__(256)
label_transfer_kernel(const uint16_t* __restrict__ table, int64_t table_rows, int C,
                      int pp, const float* __restrict__ scores, const int64_t* __restrict__ idx,
                      const float* __restrict__ qnorm, int64_t Q, int k, float beta,
                      float* __restrict__ out) {
  __shared__ float s_w[8][kMaxK];
  __shared__ int64_t s_i[8][kMaxK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t qi = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (qi >= Q) return;
  // F.normalize clamps the norm at eps = 1e-12 (hbird_eval.py:594)
  const float qn = fmaxf(qnorm[qi], 1e-12f);
  // logits = cos / beta, softmax over the k neighbours (hbird_eval.py:603-604)
  float mx = -INFINITY;
  for (int j = lane; j < k; j += 32) {
    const int64_t id = idx[qi * k + j];
    const bool ok = id >= 0 && id < table_rows;
    const float logit = ok ? (scores[qi * k + j] / qn) / beta : -INFINITY;
    s_w[warp][j] = logit;
    s_i[warp][j] = ok ? id : -1;
    mx = fmaxf(mx, logit);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < k; j += 32) {
    const float l = s_w[warp][j];
    const float e = (l == -INFINITY) ? 0.f : expf(l - mx);
    s_w[warp][j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  // softmax weights; a missing neighbour (id < 0) keeps weight 0 and reads row 0, so the gather
  // loop below has no branch and its loads can be issued ahead of the arithmetic
  for (int j = lane; j < k; j += 32) {
    const bool ok = s_i[warp][j] >= 0;
    s_w[warp][j] = ok ? s_w[warp][j] / sum : 0.f;
    if (!ok) s_i[warp][j] = 0;
  }
  __syncwarp();
  const float fpp = static_cast<float>(pp);
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
#pragma unroll 10
    for (int j = 0; j < k; ++j) {
      // soft label = histogram / pixels-per-patch (one_hot(...).mean(3), hbird_eval.py:319-320)
      const float lab = static_cast<float>(__ldg(table + s_i[warp][j] * C + c)) / fpp;
      acc += s_w[warp][j] * lab;
    }
    out[qi * C + c] = acc;
  }
}

// One thread per output pixel; label_hat (B, S*S, C) is read as the (B, C, S, S) view.
__global__ void __launch_bounds__(256)
upsample_argmax_kernel(const float* __restrict__ label_hat, int B, int S, int C, int H, int W,
                       float scale_h, float scale_w, uint8_t* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(B) * H * W;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(t % W);
    const int y = static_cast<int>((t / W) % H);
    const int b = static_cast<int>(t / (static_cast<int64_t>(W) * H));
    // torch area_pixel_compute_source_index, align_corners=False, clamp at 0
    float sy = scale_h * (static_cast<float>(y) + 0.5f) - 0.5f;
    float sx = scale_w * (static_cast<float>(x) + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    int y0 = static_cast<int>(sy);
    int x0 = static_cast<int>(sx);
    y0 = y0 > S - 1 ? S - 1 : y0;
    x0 = x0 > S - 1 ? S - 1 : x0;
    const int y1 = y0 + (y0 < S - 1 ? 1 : 0);
    const int x1 = x0 + (x0 < S - 1 ? 1 : 0);
    float ly1 = sy - static_cast<float>(y0);
    float lx1 = sx - static_cast<float>(x0);
    ly1 = fminf(fmaxf(ly1, 0.f), 1.f);
    lx1 = fminf(fmaxf(lx1, 0.f), 1.f);
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float* base = label_hat + static_cast<int64_t>(b) * S * S * C;
    const float* p00 = base + (static_cast<int64_t>(y0) * S + x0) * C;
    const float* p01 = base + (static_cast<int64_t>(y0) * S + x1) * C;
    const float* p10 = base + (static_cast<int64_t>(y1) * S + x0) * C;
    const float* p11 = base + (static_cast<int64_t>(y1) * S + x1) * C;
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
      // same association as torch's separable linear interpolation: x first, then y
      const float top = __fadd_rn(__fmul_rn(lx0, __ldg(p00 + c)), __fmul_rn(lx1, __ldg(p01 + c)));
      const float bot = __fadd_rn(__fmul_rn(lx0, __ldg(p10 + c)), __fmul_rn(lx1, __ldg(p11 + c)));
      const float v = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
      if (v > best) {  // strict: the first maximum wins, as torch.argmax
        best = v;
        arg = c;
      }
    }
    out[t] = static_cast<uint8_t>(arg);
  }
}


// ---- fused tail: mask decode + bilinear upsample + argmax + confusion histogram ---------------
// Replaces hbird_eval.py:219 (`(y*255).long()`), :235-243 (permute, F.interpolate bilinear
// align_corners=False, argmax) and PredsmIoU.update's bincount (eval_metrics.py:73-109) in ONE pass
// over the output pixels: neither the decoded mask nor the prediction map has to exist in HBM.
// A CTA takes (image, band of output rows) items.  It stages the label_hat rows of the patch cells
// that band interpolates between in shared memory, then every thread owns one output column x of
// 8 consecutive rows: the horizontal blend of the two cell rows (top / bot, per class) is computed
// once per source-row change and reused for the rows below it; the vertical blend + first-maximum
// argmax runs per pixel.  Arithmetic is torch's, operation for operation (x first, then y, no FMA
// contraction).  (gt, pred) pairs go into a per-CTA shared-memory histogram as runs along the
// column and are flushed once per CTA into the int64 matrix: exact integer arithmetic.
constexpr int kTailThreads = 256;
constexpr int kTailRows = 8;  // output rows per thread unit

struct TailParams {
  const float* label_hat;  // (B, S*S, C)
  const float* y;          // (B, H, W) fp32 = class id / 255 (loader contract) or NULL
  const uint8_t* gt;       // (B, H, W) decoded ids, used when y is NULL; both NULL = no scoring
  int B, S, C, H, W;
  float scale_h, scale_w;
  int ignore_index;
  int band_rows;           // output rows per item (multiple of kTailRows)
  int n_bands;
  int max_cell_rows;       // shared-memory capacity in cell rows
  unsigned long long* conf;  // (C, C) int64 or NULL
  uint8_t* pred;           // (B, H, W) or NULL
};

// torch area_pixel_compute_source_index (align_corners=False, clamped at 0) + index/lambda split
__device__ __forceinline__ void src_index(float scale, int dst, int size, int& i0, int& i1, float& l0, float& l1) {
  float s = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  int a = static_cast<int>(s);
  a = a > size - 1 ? size - 1 : a;
  i0 = a;
  i1 = a + (a < size - 1 ? 1 : 0);
  float l = s - static_cast<float>(a);
  l = fminf(fmaxf(l, 0.f), 1.f);
  l1 = l;
  l0 = 1.f - l;
}

template <int CC>
__global__ void __launch_bounds__(kTailThreads)
predict_score_kernel(const TailParams p) {
  extern __shared__ __align__(16) uint8_t tail_smem[];
  float* cell = reinterpret_cast<float*>(tail_smem);                     // (max_cell_rows, S, C)
  uint32_t* hist = reinterpret_cast<uint32_t*>(cell + static_cast<size_t>(p.max_cell_rows) * p.S * p.C);
  const int C = p.C, S = p.S, H = p.H, W = p.W;
  const int bins = C * C;
  const bool score = p.conf != nullptr;
  if (score) {
    for (int i = threadIdx.x; i < bins; i += kTailThreads) hist[i] = 0;
  }
  const int row_floats = S * C;
  const int items = p.B * p.n_bands;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / p.n_bands, band = item % p.n_bands;
    const int Y0 = band * p.band_rows;
    const int Y1 = (Y0 + p.band_rows < H) ? Y0 + p.band_rows : H;
    int r_lo, r_hi, t0, t1;
    float f0, f1;
    src_index(p.scale_h, Y0, S, r_lo, t1, f0, f1);
    src_index(p.scale_h, Y1 - 1, S, t0, r_hi, f0, f1);
    const int n_r = r_hi - r_lo + 1;  // <= max_cell_rows by construction of band_rows
    __syncthreads();                  // the previous item's cells are no longer read; hist zeroed
    const float* src = p.label_hat + (static_cast<int64_t>(b) * S + r_lo) * row_floats;
    for (int i = threadIdx.x; i < n_r * row_floats; i += kTailThreads) cell[i] = __ldg(src + i);
    __syncthreads();
    const int groups = (Y1 - Y0 + kTailRows - 1) / kTailRows;
    for (int u = threadIdx.x; u < groups * W; u += kTailThreads) {
      const int x = u % W;
      const int ya = Y0 + (u / W) * kTailRows;
      int x0, x1;
      float lx0, lx1;
      src_index(p.scale_w, x, S, x0, x1, lx0, lx1);
      float best[kTailRows];
      int arg[kTailRows];
#pragma unroll
      for (int r = 0; r < kTailRows; ++r) {
        best[r] = -INFINITY;
        arg[r] = 0;
      }
      for (int cc0 = 0; cc0 < C; cc0 += CC) {
        float top[CC], bot[CC];
        int cur = -1;
#pragma unroll
        for (int r = 0; r < kTailRows; ++r) {
          const int yy = ya + r;
          if (yy < Y1) {
            int y0, y1;
            float ly0, ly1;
            src_index(p.scale_h, yy, S, y0, y1, ly0, ly1);
            if (y0 != cur) {
              cur = y0;
              const float* p00 = cell + ((y0 - r_lo) * S + x0) * C + cc0;
              const float* p01 = cell + ((y0 - r_lo) * S + x1) * C + cc0;
              const float* p10 = cell + ((y1 - r_lo) * S + x0) * C + cc0;
              const float* p11 = cell + ((y1 - r_lo) * S + x1) * C + cc0;
#pragma unroll
              for (int c = 0; c < CC; ++c) {
                if (cc0 + c < C) {
                  // same association as torch's separable linear interpolation: x first, then y
                  top[c] = __fadd_rn(__fmul_rn(lx0, p00[c]), __fmul_rn(lx1, p01[c]));
                  bot[c] = __fadd_rn(__fmul_rn(lx0, p10[c]), __fmul_rn(lx1, p11[c]));
                }
              }
            }
#pragma unroll
            for (int c = 0; c < CC; ++c) {
              if (cc0 + c < C) {
                const float v = __fadd_rn(__fmul_rn(ly0, top[c]), __fmul_rn(ly1, bot[c]));
                if (v > best[r]) {  // strict: the first maximum wins, as torch.argmax
                  best[r] = v;
                  arg[r] = cc0 + c;
                }
              }
            }
          }
        }
      }
      // emit predictions, score them against the ground truth: runs down the column
      int run_bin = -1;
      uint32_t run_cnt = 0;
#pragma unroll
      for (int r = 0; r < kTailRows; ++r) {
        const int yy = ya + r;
        if (yy < Y1) {
          const int64_t pix = (static_cast<int64_t>(b) * H + yy) * W + x;
          if (p.pred != nullptr) p.pred[pix] = static_cast<uint8_t>(arg[r]);
          if (score) {
            int gv;
            if (p.y != nullptr) {
              // (y * 255).long(): fp32 multiply, truncate toward zero; ids are bytes
              gv = static_cast<int>(static_cast<uint8_t>(static_cast<long long>(__fmul_rn(__ldg(p.y + pix), 255.0f))));
            } else {
              gv = __ldg(p.gt + pix);
            }
            const int bin = (gv != p.ignore_index && gv < C) ? gv * C + arg[r] : -1;
            if (bin != run_bin) {
              if (run_bin >= 0) atomicAdd(hist + run_bin, run_cnt);
              run_bin = bin;
              run_cnt = 0;
            }
            ++run_cnt;
          }
        }
      }
      if (score && run_bin >= 0) atomicAdd(hist + run_bin, run_cnt);
    }
  }
  if (score) {
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += kTailThreads) {
      const uint32_t c = hist[i];
      if (c) atomicAdd(p.conf + i, static_cast<unsigned long long>(c));
    }
  }
}

// Launch the fused tail.  Returns HB_ERR_UNSUPPORTED (without an error message) when the class
// count is too large for the shared-memory histogram; callers then run K4b + K5 separately.
int predict_score_launch(const float* label_hat, int B, int S, int C, int H, int W, const float* y,
                         const uint8_t* gt, int ignore_index, int64_t* conf, uint8_t* pred, cudaStream_t st) {
  TailParams p;
  p.label_hat = label_hat;
  p.y = y;
  p.gt = gt;
  p.B = B; p.S = S; p.C = C; p.H = H; p.W = W;
  // torch area_pixel_compute_scale<float>: input_size / output_size in fp32
  p.scale_h = static_cast<float>(S) / static_cast<float>(H);
  p.scale_w = static_cast<float>(S) / static_cast<float>(W);
  p.ignore_index = ignore_index;
  p.conf = reinterpret_cast<unsigned long long*>(conf);
  p.pred = pred;
  const size_t hist_bytes = conf ? sizeof(uint32_t) * C * C : 0;
  const size_t row_bytes = sizeof(float) * S * C;
  const size_t budget = 200 * 1024;
  int band = 0;
  for (int cand : {32, 16, 8}) {
    // cell rows a band of `cand` output rows can touch: its extent in source rows plus both neighbours
    const int rows = static_cast<int>(static_cast<int64_t>(cand) * S / H) + 3;
    if (hist_bytes + rows * row_bytes <= budget) {
      band = cand;
      p.max_cell_rows = rows < S ? rows : S;
      break;
    }
  }
  if (band == 0) return HB_ERR_UNSUPPORTED;
  p.band_rows = band;
  p.n_bands = (H + band - 1) / band;
  const size_t smem = hist_bytes + p.max_cell_rows * row_bytes;
  const int sms = device_sm_count();
  const int per_sm = smem <= 24 * 1024 ? 4 : (smem <= 48 * 1024 ? 2 : 1);
  int64_t blocks = static_cast<int64_t>(B) * p.n_bands;
  if (blocks > static_cast<int64_t>(sms) * per_sm) blocks = static_cast<int64_t>(sms) * per_sm;
  const int cc = C <= 8 ? 8 : (C <= 16 ? 16 : (C <= 24 ? 24 : 32));
#define HB_TAIL(CCV)                                                                                   \
  do {                                                                                                 \
    if (smem > 48 * 1024)                                                                              \
      HB_CHECK_CUDA(cudaFuncSetAttribute(predict_score_kernel<CCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         static_cast<int>(budget)));                                   \
    predict_score_kernel<CCV><<<static_cast<unsigned>(blocks), kTailThreads, smem, st>>>(p);           \
  } while (0)
  if (cc == 8) HB_TAIL(8);
  else if (cc == 16) HB_TAIL(16);
  else if (cc == 24) HB_TAIL(24);
  else HB_TAIL(32);
#undef HB_TAIL
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

}  // namespace hb

extern "C" {

int hb_label_transfer(const uint16_t* label_table_dev, int64_t table_rows, int C, int patch_pixels,
                      const float* scores_dev, const int64_t* idx_dev, const float* qnorm_dev, int64_t Q,
                      int k, float beta, float* out_label_hat_dev, void* stream) {
  HB_REQUIRE(C >= 1 && C <= 256, "hb_label_transfer: C=%d not in [1, 256]", C);
  HB_REQUIRE(k >= 1 && k <= hb::kMaxK, "hb_label_transfer: k=%d not in [1, %d]", k, hb::kMaxK);
  HB_REQUIRE(patch_pixels >= 1, "hb_label_transfer: patch_pixels < 1");
  HB_REQUIRE(beta > 0.f, "hb_label_transfer: beta must be positive");
  HB_REQUIRE(Q >= 0, "hb_label_transfer: Q < 0");
  if (Q == 0) return HB_OK;
  HB_REQUIRE(label_table_dev && scores_dev && idx_dev && qnorm_dev && out_label_hat_dev, "hb_label_transfer: NULL pointer");
  const unsigned blocks = static_cast<unsigned>(hb::ceil_div64(Q, 8));
  hb::label_transfer_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      label_table_dev, table_rows, C, patch_pixels, scores_dev, idx_dev, qnorm_dev, Q, k, beta, out_label_hat_dev);
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

int hb_upsample_argmax(const float* label_hat_dev, int B, int S, int C, int H, int W, uint8_t* out_pred_dev,
                       void* stream) {
  HB_REQUIRE(B >= 0 && S >= 1 && C >= 1 && C <= 256 && H >= 1 && W >= 1, "hb_upsample_argmax: bad shape B=%d S=%d C=%d H=%d W=%d", B, S, C, H, W);
  if (B == 0) return HB_OK;
  HB_REQUIRE(label_hat_dev && out_pred_dev, "hb_upsample_argmax: NULL pointer");
  const int64_t total = static_cast<int64_t>(B) * H * W;
  int64_t blocks = hb::ceil_div64(total, 256);
  const int64_t cap = static_cast<int64_t>(hb::device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  // torch area_pixel_compute_scale<float>: input_size / output_size in fp32
  const float scale_h = static_cast<float>(S) / static_cast<float>(H);
  const float scale_w = static_cast<float>(S) / static_cast<float>(W);
  hb::upsample_argmax_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      label_hat_dev, B, S, C, H, W, scale_h, scale_w, out_pred_dev);
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

int hb_predict_score(const float* label_hat_dev, int B, int S, int C, int H, int W, const float* y_dev,
                     const uint8_t* gt_dev, int ignore_index, int64_t* conf_dev, uint8_t* out_pred_dev,
                     void* stream) {
  HB_REQUIRE(B >= 0 && S >= 1 && C >= 1 && C <= 256 && H >= 1 && W >= 1, "hb_predict_score: bad shape B=%d S=%d C=%d H=%d W=%d", B, S, C, H, W);
  if (B == 0) return HB_OK;
  HB_REQUIRE(label_hat_dev != nullptr, "hb_predict_score: label_hat is NULL");
  HB_REQUIRE(conf_dev == nullptr || y_dev != nullptr || gt_dev != nullptr, "hb_predict_score: scoring needs y_dev or gt_dev");
  HB_REQUIRE(conf_dev != nullptr || out_pred_dev != nullptr, "hb_predict_score: nothing to do (conf and pred both NULL)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = hb::predict_score_launch(label_hat_dev, B, S, C, H, W, y_dev, gt_dev, ignore_index, conf_dev, out_pred_dev, st);
  if (rc != HB_ERR_UNSUPPORTED) return rc;
  // very large class counts: the (C, C) histogram does not fit beside the cells -> separate kernels
  HB_REQUIRE(out_pred_dev != nullptr, "hb_predict_score: C=%d needs out_pred_dev (the unfused path stores the prediction map)", C);
  rc = hb_upsample_argmax(label_hat_dev, B, S, C, H, W, out_pred_dev, stream);
  if (rc != HB_OK || conf_dev == nullptr) return rc;
  HB_REQUIRE(gt_dev != nullptr, "hb_predict_score: C=%d needs gt_dev (decode the mask with hb_decode_mask first)", C);
  return hb_confusion_accumulate(gt_dev, out_pred_dev, static_cast<int64_t>(B) * H * W, C, C, ignore_index, conf_dev, stream);
}

}  // extern "C"
