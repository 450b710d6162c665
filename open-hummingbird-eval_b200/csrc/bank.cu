// K1 — memory-bank construction: fused L2-normalise / cast / pack + per-patch class histogram.
// Replaces hbird_eval.py:309-329 (+ :332-355 when a sampler selection is given) and the
// index.add H2D copy of search_faiss.py:78-81.  HBM-bound: per row it reads 4*d B of fp32
// features and ps*ps B of mask, writes 2*dpad B bf16 (+ 4*d B fp32 copy) + 2*C B histogram.
#include "common.cuh"

namespace hb {

constexpr int kPackWarps = 8;
constexpr int kPackPrefetch = 256;  // mask pixels per patch prefetched into registers (16x16)

// One warp per bank row.
//   LABEL_MODE 0: histogram from the uint8 mask (B, S*ps, S*ps)
//   LABEL_MODE 1: histogram recovered from fp32 soft labels (n, C)
template <int LABEL_MODE>
__global__ void __launch_bounds__(kPackWarps * 32, 4)
pack_rows_kernel(const float* __restrict__ feats, const uint8_t* __restrict__ mask,
                 const float* __restrict__ soft, const int32_t* __restrict__ sel, int64_t n,
                 int d, int dpad, int C, int S, int ps, int pp, int normalise, int l2,
                 __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32,
                 uint16_t* __restrict__ out_hist) {
  extern __shared__ uint32_t s_hist[];  // kPackWarps * C, then kPackPrefetch pixel offsets
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* hist = s_hist + warp * C;
  const int d4 = d >> 2;  // d % 4 == 0 is enforced on the host
  const int W = S * ps;  // pp = pixels per patch (ps*ps when a mask is given; the bank's value otherwise)
  // offset of pixel t of a patch inside the mask, tabulated once per block (no per-pixel division)
  uint32_t* pix_off = s_hist + kPackWarps * C;
  if (LABEL_MODE == 0) {
    for (int t = threadIdx.x; t < kPackPrefetch && t < pp; t += blockDim.x) pix_off[t] = (t / ps) * W + (t % ps);
    __syncthreads();
  }

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kPackWarps + warp; row < n;
       row += static_cast<int64_t>(gridDim.x) * kPackWarps) {
    const int64_t src = sel ? static_cast<int64_t>(sel[row]) : row;
    const float4* in4 = reinterpret_cast<const float4*>(feats + src * d);

    // ---- label record, part 1: issue the mask loads of the first kPackPrefetch pixels now so
    // that their latency overlaps the feature pass ----
    int cls_pre[kPackPrefetch / 32];
    const uint8_t* mrow = nullptr;
    if (LABEL_MODE == 0) {
      const int S2 = S * S;
      const int b = static_cast<int>(src / S2), p = static_cast<int>(src % S2);
      const int py = p / S, px = p % S;
      mrow = mask + (static_cast<int64_t>(b) * W + py * ps) * W + px * ps;
#pragma unroll
      for (int u = 0; u < kPackPrefetch / 32; ++u) {
        const int t = u * 32 + lane;
        cls_pre[u] = t < pp ? static_cast<int>(__ldg(mrow + pix_off[t])) : -1;
      }
    }

    // ---- features: ||f||_2, then f / ||f||_2 (true division, no epsilon) ----
    float ss = 0.f;
    for (int i = lane; i < d4; i += 32) {
      float4 v = __ldg(in4 + i);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    const float nrm = normalise ? sqrtf(ss) : 1.0f;
    __nv_bfloat16* ob = out_bf16 + row * dpad;
    float4* of = out_f32 ? reinterpret_cast<float4*>(out_f32 + row * d) : nullptr;
    float ss_out = 0.f;  // ||stored row||^2, only needed by the L2 metric
    for (int i = lane; i < d4; i += 32) {
      float4 v = __ldg(in4 + i);  // second read hits L1/L2
      v.x /= nrm; v.y /= nrm; v.z /= nrm; v.w /= nrm;
      ss_out += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      if (of) of[i] = v;
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
      __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(ob + 4 * i) = pk;
    }
    for (int i = d + lane; i < dpad; i += 32) ob[i] = __float2bfloat16(0.f);
    if (l2) {
      // L2 metric (search_faiss.py:45-46): ||q-x||^2 = ||q||^2 - 2 (q.x - ||x||^2/2).  The tensor
      // pass ranks by q.x - ||x||^2/2: three extra bank columns carry -||x||^2/2 split into
      // bf16 hi/mid/lo parts (~24 bits), the query carries 1.0 in the same columns.
      ss_out = warp_sum(ss_out);
      __syncwarp();
      if (lane == 0) {
        const float h = -0.5f * ss_out;
        const __nv_bfloat16 h0 = __float2bfloat16(h);
        const float r1 = h - __bfloat162float(h0);
        const __nv_bfloat16 h1 = __float2bfloat16(r1);
        const __nv_bfloat16 h2 = __float2bfloat16(r1 - __bfloat162float(h1));
        ob[d] = h0; ob[d + 1] = h1; ob[d + 2] = h2;
      }
    }

    // ---- label record: per-patch class histogram ----
    if (LABEL_MODE == 0) {
      for (int c = lane; c < C; c += 32) hist[c] = 0;
      __syncwarp();
#pragma unroll
      for (int u = 0; u < kPackPrefetch / 32; ++u) {
        if (u * 32 < pp) {
          const int cls = cls_pre[u];
          // warp-aggregate equal classes, one shared-memory add per distinct class
          const unsigned peers = __match_any_sync(0xffffffffu, cls);
          if (cls >= 0 && cls < C && lane == (__ffs(peers) - 1)) hist[cls] += __popc(peers);
          __syncwarp();
        }
      }
      for (int t0 = kPackPrefetch; t0 < pp; t0 += 32) {  // patches larger than 16x16
        const int t = t0 + lane;
        const int cls = t < pp ? static_cast<int>(mrow[(t / ps) * W + (t % ps)]) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, cls);
        if (cls >= 0 && cls < C && lane == (__ffs(peers) - 1)) hist[cls] += __popc(peers);
        __syncwarp();
      }
      for (int c = lane; c < C; c += 32) out_hist[row * C + c] = static_cast<uint16_t>(hist[c]);
      __syncwarp();
    } else {
      const float fpp = static_cast<float>(pp);
      for (int c = lane; c < C; c += 32)
        out_hist[row * C + c] = static_cast<uint16_t>(rintf(soft[row * C + c] * fpp));
    }
  }
}

__global__ void export_rows_kernel(const __nv_bfloat16* __restrict__ bf, const float* __restrict__ f32,
                                   const uint16_t* __restrict__ hist, int64_t row0, int64_t n,
                                   int d, int dpad, int C, int pp, float* __restrict__ feats_out,
                                   float* __restrict__ labels_out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (feats_out) {
    for (int64_t i = tid; i < n * d; i += stride) {
      const int64_t r = i / d;
      const int c = static_cast<int>(i % d);
      feats_out[i] = f32 ? f32[(row0 + r) * d + c] : __bfloat162float(bf[(row0 + r) * dpad + c]);
    }
  }
  if (labels_out) {
    const float fpp = static_cast<float>(pp);
    for (int64_t i = tid; i < n * C; i += stride)
      labels_out[i] = static_cast<float>(hist[row0 * C + i]) / fpp;
  }
}

// N1 — bounded-memory sampler (hbird_eval.py:447-517): one block per image.  For every patch: the set
// of classes present (bitmap), then per image the number of patches containing each class, then
// score = sum of those frequencies over the patch's classes, times the caller's U(0,1) draw (the
// reference's CPU RNG stream, uploaded), 1e6 for an empty patch; the K smallest scores, ascending
// (ties: lower patch index), are the picks.  Integer/float arithmetic is exact (small integers).
constexpr int kSampleThreads = 256;
constexpr int kSampleMaxPatches = 4096;

__global__ void __launch_bounds__(kSampleThreads)
sample_patches_kernel(const uint8_t* __restrict__ mask, int S, int ps, int C,
                      const float* __restrict__ uniform, int K, int32_t* __restrict__ sel) {
  extern __shared__ unsigned char s_raw[];
  const int SS = S * S, W = S * ps;
  int n2 = 1;
  while (n2 < SS) n2 <<= 1;
  uint32_t* present = reinterpret_cast<uint32_t*>(s_raw);           // SS * 8 words (256 class bits)
  int* freq = reinterpret_cast<int*>(present + SS * 8);             // 256
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(freq + 256);  // n2
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < 256; c += blockDim.x) freq[c] = 0;
  __syncthreads();
  for (int p = threadIdx.x; p < SS; p += blockDim.x) {
    uint32_t bits[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int py = p / S, px = p % S;
    const uint8_t* base = mask + (static_cast<int64_t>(b) * W + py * ps) * W + px * ps;
    for (int i = 0; i < ps; ++i)
      for (int j = 0; j < ps; ++j) {
        const int cls = base[i * W + j];
        if (cls < C) bits[cls >> 5] |= 1u << (cls & 31);
      }
#pragma unroll
    for (int wd = 0; wd < 8; ++wd) {
      present[p * 8 + wd] = bits[wd];
      uint32_t m = bits[wd];
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        atomicAdd(&freq[wd * 32 + bit], 1);
      }
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < n2; p += blockDim.x) {
    unsigned long long key = ~0ull;  // padding sorts last
    if (p < SS) {
      float score = 0.f;
      bool any = false;
#pragma unroll
      for (int wd = 0; wd < 8; ++wd) {
        uint32_t m = present[p * 8 + wd];
        any |= m != 0;
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          score += static_cast<float>(freq[wd * 32 + bit]);
        }
      }
      score = any ? __fmul_rn(score, uniform[static_cast<int64_t>(b) * SS + p]) : 1e6f;
      key = (static_cast<unsigned long long>(f32_to_ordered(score)) << 32) | static_cast<uint32_t>(p);
    }
    keys[p] = key;
  }
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int partner = i ^ j;
        if (partner > i) {
          const bool asc = (i & k) == 0;
          const unsigned long long a = keys[i], c2 = keys[partner];
          if ((a > c2) == asc) {
            keys[i] = c2;
            keys[partner] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < K; j += blockDim.x)
    sel[static_cast<int64_t>(b) * K + j] = b * SS + static_cast<int32_t>(keys[j] & 0xffffffffu);
}

static int launch_pack(Bank* b, const float* feats, const uint8_t* mask, const float* soft,
                       const int32_t* sel, int64_t n, int S, int ps, int normalise,
                       cudaStream_t st) {
  if (n == 0) return HB_OK;
  const int64_t row0 = b->rows;
  int64_t blocks = ceil_div64(n, kPackWarps);
  const int64_t max_blocks = static_cast<int64_t>(b->num_sms) * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  const size_t smem = sizeof(uint32_t) * (kPackWarps * b->C + kPackPrefetch);
  __nv_bfloat16* ob = b->feat_bf16 + row0 * b->dpad;
  float* of = b->feat_f32 ? b->feat_f32 + row0 * b->d : nullptr;
  uint16_t* oh = b->label_hist + row0 * b->C;
  if (mask) {
    pack_rows_kernel<0><<<static_cast<unsigned>(blocks), kPackWarps * 32, smem, st>>>(
        feats, mask, nullptr, sel, n, b->d, b->dpad, b->C, S, ps, b->pp, normalise, (b->flags & HB_BANK_L2) ? 1 : 0, ob, of, oh);
  } else {
    pack_rows_kernel<1><<<static_cast<unsigned>(blocks), kPackWarps * 32, smem, st>>>(
        feats, nullptr, soft, sel, n, b->d, b->dpad, b->C, S, ps, b->pp, normalise, (b->flags & HB_BANK_L2) ? 1 : 0, ob, of, oh);
  }
  HB_CHECK_CUDA(cudaGetLastError());
  b->rows += n;
  return HB_OK;
}

}  // namespace hb

using hb::Bank;

extern "C" {

int hb_bank_create(int device, int d, int num_classes, int patch_pixels, int64_t capacity_rows,
                   unsigned flags, hb_bank_t** bank_out) {
  HB_REQUIRE(bank_out != nullptr, "hb_bank_create: bank_out is NULL");
  *bank_out = nullptr;
  int num_sms = 0;
  int rc = hb_device_check(device, &num_sms);
  if (rc != HB_OK) return rc;
  HB_REQUIRE(d >= 8 && d % 8 == 0 && d <= 8192, "hb_bank_create: d=%d must be a multiple of 8 in [8, 8192]", d);
  HB_REQUIRE(num_classes >= 1 && num_classes <= 256, "hb_bank_create: num_classes=%d not in [1, 256]", num_classes);
  HB_REQUIRE(patch_pixels >= 1 && patch_pixels <= 65535, "hb_bank_create: patch_pixels=%d not in [1, 65535]", patch_pixels);
  HB_REQUIRE(capacity_rows >= 1 && capacity_rows < (int64_t(1) << 31), "hb_bank_create: capacity_rows=%lld not in [1, 2^31)", (long long)capacity_rows);
  HB_CHECK_CUDA(cudaSetDevice(device));
  Bank* b = new Bank();
  b->device = device;
  b->num_sms = num_sms;
  b->d = d;
  b->dpad = (d + ((flags & HB_BANK_L2) ? 3 : 0) + 63) / 64 * 64;  // L2: + 3 norm columns
  b->C = num_classes;
  b->pp = patch_pixels;
  b->flags = flags;
  b->capacity = capacity_rows;
  cudaError_t e = cudaMalloc(&b->feat_bf16, sizeof(__nv_bfloat16) * capacity_rows * b->dpad);
  if (e == cudaSuccess && (flags & HB_BANK_KEEP_F32))
    e = cudaMalloc(&b->feat_f32, sizeof(float) * capacity_rows * d);
  if (e == cudaSuccess) e = cudaMalloc(&b->label_hist, sizeof(uint16_t) * capacity_rows * num_classes);
  if (e != cudaSuccess) {
    hb::set_error("hb_bank_create: cudaMalloc failed for %lld rows x d=%d: %s", (long long)capacity_rows, d, cudaGetErrorString(e));
    (void)cudaGetLastError();
    hb_bank_destroy(reinterpret_cast<hb_bank_t*>(b));
    return e == cudaErrorMemoryAllocation ? HB_ERR_OOM : HB_ERR_CUDA;
  }
  *bank_out = reinterpret_cast<hb_bank_t*>(b);
  return HB_OK;
}

int hb_bank_destroy(hb_bank_t* bank) {
  if (!bank) return HB_OK;
  Bank* b = reinterpret_cast<Bank*>(bank);
  cudaSetDevice(b->device);
  // cudaFree itself waits for queued work that still uses a block; no explicit device sync
  if (b->feat_bf16) cudaFree(b->feat_bf16);
  if (b->feat_f32) cudaFree(b->feat_f32);
  if (b->label_hist) cudaFree(b->label_hist);
  if (b->ws) cudaFree(b->ws);
  for (int i = 0; i < 64; ++i) {
    if (b->ev_begin[i]) cudaEventDestroy(b->ev_begin[i]);
    if (b->ev_end[i]) cudaEventDestroy(b->ev_end[i]);
    if (b->ev_rerank[i]) cudaEventDestroy(b->ev_rerank[i]);
    if (b->ev_rerank0[i]) cudaEventDestroy(b->ev_rerank0[i]);
  }
  for (int i = 0; i < 2; ++i) {
    if (b->pipe[i].buf) cudaFree(b->pipe[i].buf);
    if (b->pipe[i].done) cudaEventDestroy(b->pipe[i].done);
  }
  (void)cudaGetLastError();
  delete b;
  return HB_OK;
}

int hb_bank_append(hb_bank_t* bank, const float* feats_dev, const uint8_t* mask_dev, int B, int S,
                   int ps, const int32_t* sel_dev, int64_t n, void* stream) {
  HB_REQUIRE(bank != nullptr, "hb_bank_append: bank is NULL");
  Bank* b = reinterpret_cast<Bank*>(bank);
  HB_REQUIRE(!b->finalized, "hb_bank_append: bank already finalized");
  HB_REQUIRE(n >= 0, "hb_bank_append: n < 0");
  if (n == 0) return HB_OK;
  HB_REQUIRE(feats_dev && mask_dev, "hb_bank_append: NULL feature or mask pointer");
  HB_REQUIRE(B >= 1 && S >= 1 && ps >= 1, "hb_bank_append: bad geometry B=%d S=%d ps=%d", B, S, ps);
  HB_REQUIRE(ps * ps == b->pp, "hb_bank_append: ps*ps=%d does not match the bank's patch_pixels=%d", ps * ps, b->pp);
  HB_REQUIRE(sel_dev != nullptr || n == static_cast<int64_t>(B) * S * S,
             "hb_bank_append: n=%lld must equal B*S*S=%lld when no selection is given", (long long)n, (long long)B * S * S);
  HB_REQUIRE(b->rows + n <= b->capacity, "hb_bank_append: %lld + %lld rows exceed capacity %lld", (long long)b->rows, (long long)n, (long long)b->capacity);
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  return hb::launch_pack(b, feats_dev, mask_dev, nullptr, sel_dev, n, S, ps, 1, static_cast<cudaStream_t>(stream));
}

int hb_bank_append_soft(hb_bank_t* bank, const float* feats_dev, const float* soft_dev, int64_t n,
                        int normalise, void* stream) {
  HB_REQUIRE(bank != nullptr, "hb_bank_append_soft: bank is NULL");
  Bank* b = reinterpret_cast<Bank*>(bank);
  HB_REQUIRE(!b->finalized, "hb_bank_append_soft: bank already finalized");
  HB_REQUIRE(n >= 0, "hb_bank_append_soft: n < 0");
  if (n == 0) return HB_OK;
  HB_REQUIRE(feats_dev && soft_dev, "hb_bank_append_soft: NULL pointer");
  HB_REQUIRE(b->rows + n <= b->capacity, "hb_bank_append_soft: %lld + %lld rows exceed capacity %lld", (long long)b->rows, (long long)n, (long long)b->capacity);
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  return hb::launch_pack(b, feats_dev, nullptr, soft_dev, nullptr, n, 1, 1, normalise, static_cast<cudaStream_t>(stream));
}

int hb_sample_patches(const uint8_t* mask_dev, int B, int S, int ps, int C, const float* uniform_dev,
                      int K, int32_t* sel_out_dev, void* stream) {
  HB_REQUIRE(B >= 0 && S >= 1 && ps >= 1 && C >= 1 && C <= 256, "hb_sample_patches: bad shape B=%d S=%d ps=%d C=%d", B, S, ps, C);
  HB_REQUIRE(S * S <= hb::kSampleMaxPatches, "hb_sample_patches: S*S=%d exceeds %d patches per image", S * S, hb::kSampleMaxPatches);
  HB_REQUIRE(K >= 1 && K <= S * S, "hb_sample_patches: K=%d not in [1, S*S=%d]", K, S * S);
  if (B == 0) return HB_OK;
  HB_REQUIRE(mask_dev && uniform_dev && sel_out_dev, "hb_sample_patches: NULL pointer");
  int n2 = 1;
  while (n2 < S * S) n2 <<= 1;
  const size_t smem = static_cast<size_t>(S) * S * 32 + 256 * 4 + static_cast<size_t>(n2) * 8;
  HB_CHECK_CUDA(cudaFuncSetAttribute(hb::sample_patches_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  hb::sample_patches_kernel<<<B, hb::kSampleThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      mask_dev, S, ps, C, uniform_dev, K, sel_out_dev);
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

int hb_bank_finalize(hb_bank_t* bank) {
  HB_REQUIRE(bank != nullptr, "hb_bank_finalize: bank is NULL");
  Bank* b = reinterpret_cast<Bank*>(bank);
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  int rc = hb::make_tmap_2d_bf16(&b->tmap_bank_cg1, b->feat_bf16, b->rows, b->dpad, 256);
  if (rc != HB_OK) return rc;
  rc = hb::make_tmap_2d_bf16(&b->tmap_bank_cg2, b->feat_bf16, b->rows, b->dpad, 128);
  if (rc != HB_OK) return rc;
  b->finalized = true;
  return HB_OK;
}

int64_t hb_bank_rows(const hb_bank_t* bank) { return bank ? reinterpret_cast<const Bank*>(bank)->rows : -1; }
int64_t hb_bank_capacity(const hb_bank_t* bank) { return bank ? reinterpret_cast<const Bank*>(bank)->capacity : -1; }
const uint16_t* hb_bank_label_table(const hb_bank_t* bank) { return bank ? reinterpret_cast<const Bank*>(bank)->label_hist : nullptr; }

int hb_bank_export(const hb_bank_t* bank, int64_t row0, int64_t n, float* feats_out_dev,
                   float* labels_out_dev, void* stream) {
  HB_REQUIRE(bank != nullptr, "hb_bank_export: bank is NULL");
  const Bank* b = reinterpret_cast<const Bank*>(bank);
  HB_REQUIRE(row0 >= 0 && n >= 0 && row0 + n <= b->rows, "hb_bank_export: rows [%lld, %lld) outside [0, %lld)", (long long)row0, (long long)(row0 + n), (long long)b->rows);
  if (n == 0 || (!feats_out_dev && !labels_out_dev)) return HB_OK;
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  hb::export_rows_kernel<<<b->num_sms * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      b->feat_bf16, b->feat_f32, b->label_hist, row0, n, b->d, b->dpad, b->C, b->pp, feats_out_dev, labels_out_dev);
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

}  // extern "C"
