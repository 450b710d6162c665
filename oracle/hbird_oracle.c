/* CPU oracle in plain C — TEST INFRASTRUCTURE ONLY.
 *
 * A second, independent restatement (the first is oracle/hbird_oracle.py, numpy) of the integer /
 * byte-exact steps of the reference's dense nearest-neighbour evaluation path, plus scalar fp32
 * versions of the search and the label transfer.  Only tests/ may load it (tests/test_oracle_c.py
 * checks it against the numpy oracle and against the golden fixtures produced by the unmodified
 * reference, tests/golden/); the product library never links or calls it.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/Makefile); contraction is off so that fp32
 * expressions round exactly as written.  Single-threaded: it is a checker for small cases.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* hbird/hbird_eval.py:219 and :309-310 — (y * 255).long(); the bank side maps 255 -> 0. */
void hbo_decode_mask(const float* y, int64_t n, int remap_255_to_0, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    const float v = y[i] * 255.0f;    /* fp32 multiply */
    long long id = (long long)v;      /* truncation toward zero */
    if (remap_255_to_0 && id == 255) id = 0;
    out[i] = (uint8_t)id;
  }
}

/* hbird/hbird_eval.py:554-573 (_patchify_gt) + :319-320 (one_hot(...).float().mean(3)): per-patch
 * class histogram; the reference's soft label is hist / (ps*ps).  mask (B, S*ps, S*ps) uint8,
 * hist (B*S*S, C) uint16; pixels with id >= C are not counted (one_hot would raise). */
void hbo_patch_histogram(const uint8_t* mask, int B, int S, int ps, int C, uint16_t* hist) {
  const int W = S * ps;
  memset(hist, 0, sizeof(uint16_t) * (size_t)B * S * S * C);
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < W; ++y)
      for (int x = 0; x < W; ++x) {
        const int cls = mask[((size_t)b * W + y) * W + x];
        if (cls < C) hist[(((size_t)b * S + y / ps) * S + x / ps) * C + cls] += 1;
      }
}

/* hbird/utils/eval_metrics.py:73-104 — PredsmIoU.update: drop gt == ignore_index (ignore < 0: keep
 * all), drop out-of-range ids, bincount(gt * P + pred).  conf (G, P) int64 is ACCUMULATED into. */
void hbo_confusion(const uint8_t* gt, const uint8_t* pred, int64_t n, int G, int P, int ignore_index,
                   int64_t* conf) {
  for (int64_t i = 0; i < n; ++i) {
    const int g = gt[i], p = pred[i];
    if (g == ignore_index || g >= G || p >= P) continue;
    conf[(size_t)g * P + p] += 1;
  }
}

/* hbird/hbird_eval.py:235-243 — label_hat (B, S*S, C) viewed as (B, C, S, S), F.interpolate(size=
 * (H, W), mode="bilinear", align_corners=False), argmax over C (first maximum wins).  The index and
 * weight arithmetic is ATen's area_pixel_compute_source_index / linear interpolation in fp32:
 * x first, then y. */
void hbo_upsample_argmax(const float* label_hat, int B, int S, int C, int H, int W, uint8_t* out) {
  const float scale_h = (float)S / (float)H, scale_w = (float)S / (float)W;
  for (int b = 0; b < B; ++b) {
    const float* base = label_hat + (size_t)b * S * S * C;
    for (int y = 0; y < H; ++y) {
      float sy = scale_h * ((float)y + 0.5f) - 0.5f;
      if (sy < 0.f) sy = 0.f;
      int y0 = (int)sy;
      if (y0 > S - 1) y0 = S - 1;
      const int y1 = y0 + (y0 < S - 1 ? 1 : 0);
      const float ly1 = fminf(fmaxf(sy - (float)y0, 0.f), 1.f), ly0 = 1.f - ly1;
      for (int x = 0; x < W; ++x) {
        float sx = scale_w * ((float)x + 0.5f) - 0.5f;
        if (sx < 0.f) sx = 0.f;
        int x0 = (int)sx;
        if (x0 > S - 1) x0 = S - 1;
        const int x1 = x0 + (x0 < S - 1 ? 1 : 0);
        const float lx1 = fminf(fmaxf(sx - (float)x0, 0.f), 1.f), lx0 = 1.f - lx1;
        const float* p00 = base + ((size_t)y0 * S + x0) * C;
        const float* p01 = base + ((size_t)y0 * S + x1) * C;
        const float* p10 = base + ((size_t)y1 * S + x0) * C;
        const float* p11 = base + ((size_t)y1 * S + x1) * C;
        float best = -INFINITY;
        int arg = 0;
        for (int c = 0; c < C; ++c) {
          const float top = lx0 * p00[c] + lx1 * p01[c];
          const float bot = lx0 * p10[c] + lx1 * p11[c];
          const float v = ly0 * top + ly1 * bot;
          if (v > best) { best = v; arg = c; }
        }
        out[((size_t)b * H + y) * W + x] = (uint8_t)arg;
      }
    }
  }
}

/* hbird/nn/search_faiss.py:39-41,83-90 — GpuIndexFlatIP.search: exact top-k by inner product,
 * descending, ties by ascending index; (-inf, -1) padding when N < k.  Scalar fp32 dot products
 * (sequential summation: last-ulp differences from a blocked SGEMM are expected). */
void hbo_search_ip(const float* q, const float* bank, int64_t Q, int64_t N, int d, int k,
                   int64_t* idx, float* dist) {
  for (int64_t i = 0; i < Q; ++i) {
    int64_t* bi = idx + i * k;
    float* bd = dist + i * k;
    for (int j = 0; j < k; ++j) { bi[j] = -1; bd[j] = -INFINITY; }
    const float* qi = q + i * d;
    for (int64_t r = 0; r < N; ++r) {
      const float* x = bank + r * d;
      float s = 0.f;
      for (int c = 0; c < d; ++c) s += qi[c] * x[c];
      /* rows arrive in ascending index, so strict > keeps the smaller index first on ties */
      if (bi[k - 1] >= 0 && !(s > bd[k - 1])) continue;
      int j = k - 1;
      while (j > 0 && (bi[j - 1] < 0 || s > bd[j - 1])) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
      bd[j] = s;
      bi[j] = r;
    }
  }
}

/* hbird/hbird_eval.py:611-637 (gather) + :575-609 (_cross_attention): label_hat[q] =
 * softmax_j(cos(q, m_j) / beta) . soft_label[idx_j], with cos from F.normalize'd query and bank
 * rows (eps 1e-12).  feature_memory (N, d) and label_memory (N, C) are the reference's tensors. */
void hbo_label_transfer(const float* q, const float* feature_memory, const float* label_memory,
                        const int64_t* idx, int64_t Q, int d, int C, int k, float beta, float* out) {
  for (int64_t i = 0; i < Q; ++i) {
    const float* qi = q + i * d;
    float qq = 0.f;
    for (int c = 0; c < d; ++c) qq += qi[c] * qi[c];
    const float qn = fmaxf(sqrtf(qq), 1e-12f);
    float* logit = (float*)malloc(sizeof(float) * (size_t)k);
    float mx = -INFINITY;
    for (int j = 0; j < k; ++j) {
      const float* m = feature_memory + idx[i * k + j] * d;
      float mm = 0.f, dot = 0.f;
      for (int c = 0; c < d; ++c) { mm += m[c] * m[c]; }
      const float mn = fmaxf(sqrtf(mm), 1e-12f);
      for (int c = 0; c < d; ++c) dot += (qi[c] / qn) * (m[c] / mn);
      logit[j] = dot / beta;
      if (logit[j] > mx) mx = logit[j];
    }
    float sum = 0.f;
    for (int j = 0; j < k; ++j) { logit[j] = expf(logit[j] - mx); sum += logit[j]; }
    for (int c = 0; c < C; ++c) {
      float acc = 0.f;
      for (int j = 0; j < k; ++j) acc += (logit[j] / sum) * label_memory[idx[i * k + j] * C + c];
      out[i * C + c] = acc;
    }
    free(logit);
  }
}
