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
  if (blocks > 148 * 16) blocks = 148 * 16;
  // torch area_pixel_compute_scale<float>: input_size / output_size in fp32
  const float scale_h = static_cast<float>(S) / static_cast<float>(H);
  const float scale_w = static_cast<float>(S) / static_cast<float>(W);
  hb::upsample_argmax_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      label_hat_dev, B, S, C, H, W, scale_h, scale_w, out_pred_dev);
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

}  // extern "C"
