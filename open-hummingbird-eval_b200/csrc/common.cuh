// Shared host/device helpers for the hbird_b200 C-ABI library (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hbird_b200.h"

namespace hb {

// ---- host: error reporting ------------------------------------------------------
void set_error(const char* fmt, ...);
void clear_error();

#define HB_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      hb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                   \
                    cudaGetErrorString(e__));                                            \
      return (e__ == cudaErrorMemoryAllocation) ? HB_ERR_OOM : HB_ERR_CUDA;              \
    }                                                                                    \
  } while (0)

#define HB_REQUIRE(cond, ...)                                                            \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      hb::set_error(__VA_ARGS__);                                                        \
      return HB_ERR_INVALID;                                                             \
    }                                                                                    \
  } while (0)

// A search between its two halves (search.cu): hb_search_begin = query prep + K2 (tensor-core
// pass) fills the slot's candidate buffer, hb_search_finish* = K2b reads it — possibly on another
// stream, while the K2 of the next batch already runs (two slots alternate).
struct PipeSlot {
  bool begun = false;
  int kp = 0, n_chunks = 0;
  int64_t Q = 0, q_pad = 0;
  const uint64_t* cand = nullptr;
  const float* qnorm = nullptr;
  int timing_slot = -1;
  cudaEvent_t done = nullptr;  // recorded by the finish on its stream; reuse of the slot waits for it
  void* buf = nullptr;         // the slot's own scratch: norms | candidate keys (grown on demand)
  size_t buf_bytes = 0;
};

// One shard of the memory bank, resident in HBM.
struct Bank {
  int device = 0;
  int num_sms = 0;
  int d = 0;          // feature dim
  int dpad = 0;       // bf16 row pitch in elements (multiple of 64)
  int C = 0;          // classes
  int pp = 0;         // pixels per patch (ps*ps)
  unsigned flags = 0;
  int64_t capacity = 0;
  int64_t rows = 0;
  bool finalized = false;
  __nv_bfloat16* feat_bf16 = nullptr;  // (capacity, dpad)
  float* feat_f32 = nullptr;           // (capacity, d) if HB_BANK_KEEP_F32
  uint16_t* label_hist = nullptr;      // (capacity, C)
  CUtensorMap tmap_bank_cg1;           // box (64, 256)
  CUtensorMap tmap_bank_cg2;           // box (64, 128)
  // search scratch (grown on demand, owned by the bank)
  void* ws = nullptr;
  size_t ws_bytes = 0;
  int cfg_cta_group = 0;
  int cfg_max_chunks = 0;
  int cfg_prefetch_tiles = -1;  // -1 = auto
  int cfg_ablate = 0;
  // co-residency of the post-processing kernels with a running search kernel (hb_coresidency_config)
  bool cfg_lean_search = false;  // 128-register build of the search kernel
  int cfg_rerank_warps = 4;      // warps (= queries) per CTA of the re-rank kernel
  int cfg_rerank_carveout = -1;  // preferred shared-memory carve-out (%) of the re-rank kernel, -1 = driver default
  unsigned long long* cfg_stats = nullptr;  // instrumented-build counters (hb_search_stats)
  bool cfg_pace = true;         // L2 pacing window of the search kernel
  int last_launches = 0;
  int fit_cache[2][3][6] = {};  // co-resident clusters per (cta_group, k', kernel variant); 0 = unknown
  // optional kernel timing (hb_search_timing)
  bool timing = false;
  int timing_count = 0;
  cudaEvent_t ev_begin[64] = {};
  cudaEvent_t ev_end[64] = {};
  cudaEvent_t ev_rerank0[64] = {};  // before / after K2b (+ fused K4a / scatter), on the finish stream
  cudaEvent_t ev_rerank[64] = {};
  PipeSlot pipe[2];
};

// ---- fused shard exchange (exchange.cu, rerank.cu) ----------------------------------------
constexpr int kMaxPeers = 16;

// Where K2b's output rows go when the bank is row-sharded over `world` GPUs: query qi belongs to
// the rank p with qsplit[p] <= qi < qsplit[p+1] and is stored straight into p's window (peer
// memory over NVLink), slot `rank`; the last CTA then publishes `step` in every peer's flag.
struct Scatter {
  int world = 0;  // 0: plain local output
  int rank = 0;
  uint32_t step = 0;
  int64_t qsplit[kMaxPeers + 1] = {};
  float* scores[kMaxPeers] = {};     // peer p: window buffer (step & 1), slot `rank`
  int64_t* idx[kMaxPeers] = {};
  uint32_t* flag[kMaxPeers] = {};    // peer p: arrival flag of source `rank`
  unsigned int* done_ctas = nullptr; // local counter of finished CTAs
  // Threshold exchange (hb_exchange_config mode 1), two phases instead of one re-rank of the shard's
  // whole top-k':  phase 1 = shortlist_kernel merges the chunk lists of a query into its sorted bf16
  // top-k' and broadcasts four order statistics of it to every peer; phase 2 = after all peers'
  // statistics have arrived, every shard derives the same lower bound of the GLOBAL k'-th best bf16
  // score and re-ranks (fp32 row gather) only its candidates at or above it.
  int phase = 0;
  uint4* stats[kMaxPeers] = {};      // peer p: statistics buffer (step & 1), source slot `rank`, one uint4 per query
  uint32_t* flag2[kMaxPeers] = {};   // peer p: arrival flag of this rank's statistics
  const uint4* stats_in = nullptr;   // own window, statistics buffer (step & 1): [source][q_cap]
  const uint32_t* flags2_in = nullptr;  // own window: statistics arrival flags
  int64_t q_cap = 0;                 // queries per source slot of the statistics buffer
  unsigned int* done_ctas2 = nullptr;
  unsigned int* timeout_flag = nullptr;
  unsigned long long timeout_ns = 0;
};

// Optional fused label transfer (K4a) at the end of K2b / K3 / K3x: when `table` is set, the warp
// that holds a query's final neighbour list also writes label_hat[q] (C floats).
struct LabelOut {
  const uint16_t* table = nullptr;  // (table_rows, C) class histograms, indexed by GLOBAL bank row
  int64_t table_rows = 0;
  int C = 0;
  int pp = 1;                       // pixels per patch: soft label = histogram / pp
  float beta = 0.02f;
  const float* qnorm = nullptr;     // ||q||_2, indexed like the kernel's query index
  float* out = nullptr;             // (queries, C)
};

// One rank's end of the exchange: a cudaMalloc'ed window other ranks map through CUDA IPC.
//   window = [result flags: kMaxPeers x u32 | statistics flags: kMaxPeers x u32, padded to 256 B]
//            2 x [scores: world x cap x kmax f32][idx: world x cap x kmax i64]
//            2 x [statistics: world x (world * cap) x uint4]
struct Exchange {
  int device = 0, rank = 0, world = 1;
  int64_t cap = 0;  // rows per slot
  int kmax = 0;
  uint32_t step = 0;        // exchanges started so far
  int64_t last_rows = 0;    // slice rows of the last scatter
  int last_k = 0;
  size_t bytes = 0;
  bool connected = false;
  bool ipc_opened[kMaxPeers] = {};
  uint8_t* window[kMaxPeers] = {};  // [rank] = own allocation, others = mapped peers
  unsigned int* done_ctas = nullptr;
  unsigned int* timeout_flag = nullptr;
  unsigned long long timeout_ms = 600000;  // bound of the merge kernel's wait for a peer
  int mode = 0;             // 0: every shard re-ranks its whole top-k'; 1: threshold exchange (see Scatter)
  // phase 2 of a threshold exchange that has been started (hb_search_scatter) but not yet issued
  struct Pending {
    bool active = false;
    Bank* bank = nullptr;
    int slot = 0;
    const float* q = nullptr;
    int64_t Q = 0, q_pad = 0, idx_offset = 0;
    int k = 0, kp = 0;
    const uint64_t* cand = nullptr;
    Scatter sc;
  } pending;
  int64_t q_cap() const { return cap * world; }
  size_t stats_off(int parity) const {
    return 256 + 2 * buffer_bytes() + static_cast<size_t>(parity) * sizeof(uint4) * static_cast<size_t>(q_cap()) * world;
  }
  size_t window_bytes() const { return stats_off(2); }
  size_t scores_off(int parity) const { return 256 + static_cast<size_t>(parity) * buffer_bytes(); }
  size_t idx_off(int parity) const { return scores_off(parity) + sizeof(float) * slot_elems() * world; }
  size_t slot_elems() const { return static_cast<size_t>(cap) * kmax; }
  size_t buffer_bytes() const { return (sizeof(float) + sizeof(int64_t)) * slot_elems() * world; }
};

int ensure_workspace(Bank* b, size_t bytes, cudaStream_t st);
// SM count of the current device (cached per device id).
int device_sm_count();
// Fused tail (label_transfer.cu): decode + upsample + argmax + confusion.  HB_ERR_UNSUPPORTED (no
// message) when C is too large for the shared-memory histogram.
int predict_score_launch(const float* label_hat, int B, int S, int C, int H, int W, const float* y,
                         const uint8_t* gt, int ignore_index, int64_t* conf, uint8_t* pred, cudaStream_t st);
// K2 + K2b (search.cu): bf16 tcgen05 pass + exact fp32 re-rank.  Outputs are optional: out_scores /
// out_idx (local top-k), sc (scatter into the owner ranks' exchange windows), lo (fused label
// transfer; lo->qnorm is filled in by the callee), dump (validation: raw score matrix only).
int search_impl(Bank* b, const float* q, int64_t Q, int k, int kp, int64_t idx_offset,
                float* out_scores, int64_t* out_idx, float* out_qnorm, float* dump, int cg_override,
                cudaStream_t st, const Scatter* sc, const LabelOut* lo);
// The two halves of search_impl (which is begin + finish of slot 0 on one stream).
int search_begin_impl(Bank* b, const float* q, int64_t Q, int kp, int slot, float* out_qnorm, float* dump,
                      int cg_override, cudaStream_t st, cudaEvent_t prepared);
int search_finish_impl(Bank* b, int slot, const float* q, int k, int64_t idx_offset, float* out_scores,
                       int64_t* out_idx, const Scatter* sc, const LabelOut* lo, cudaStream_t st);
int rerank_launch(const Bank* b, const float* q, int64_t Q, int k, int kp, int n_chunks,
                  int64_t q_pad, const uint64_t* cand, int64_t idx_offset, float* out_scores,
                  int64_t* out_idx, const Scatter* sc, const LabelOut* lo, cudaStream_t st);  // rerank.cu
int merge_window_launch(const Exchange* x, uint32_t step, int64_t rows, int k, float* out_scores,
                        int64_t* out_idx, const LabelOut* lo, cudaStream_t st);  // rerank.cu
// Encode a 2-D bf16 row-major (rows, cols_pad) tensor map with a (64 x box_rows) box and
// 128-byte swizzle.  Resolved through cudaGetDriverEntryPoint: no link-time libcuda.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, int64_t rows, int cols_pad,
                      int box_rows);

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device: small utilities --------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Order-preserving map float -> uint32 (larger float <=> larger uint); NaN never passes the
// `>` filters upstream so it is not special-cased here.
__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float f) {
  uint32_t b;
#ifdef __CUDA_ARCH__
  b = __float_as_uint(f);
#else
  __builtin_memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  __builtin_memcpy(&f, &b, 4);
  return f;
#endif
}
// Candidate key: high word = ordered score, low word = ~row so that among equal scores the
// smaller row index is the larger key (faiss keeps the smaller id first on ties).
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return (static_cast<uint64_t>(f32_to_ordered(score)) << 32) | static_cast<uint32_t>(~row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) {
  return ordered_to_f32(static_cast<uint32_t>(k >> 32));
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) {
  return ~static_cast<uint32_t>(k);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__

}  // namespace hb
