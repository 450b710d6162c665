// K2 — the search: S = Q . Bank^T on the 5th-gen tensor cores with the top-k' selection fused
// into the epilogue, so the (Q x N) score matrix never leaves the SM.
// Replaces GpuIndexFlatIP.search as called from search_faiss.py:83-90 (first, bf16 pass; the
// exact fp32 re-rank that restores the reference's metric is K2b in rerank.cu).
//
// Structure (one persistent CTA pair with cta_group::2 -- or single CTA with cta_group::1 -- per SM):
//   warp 8      TMA producer: streams 128x64 query tiles and (256/CG)x64 bank tiles (bf16,
//               128B-swizzled) through a STAGES-deep shared-memory ring; L2-prefetches bank tiles
//               ahead and paces itself against the other CTAs of the wave (see kPaceTiles).
//   warp 9      tcgen05.mma issuer (pair leader only): 128*CG x 256 x 16 UMMAs into one of two
//               256-column TMEM accumulators; tcgen05.commit frees ring slots and publishes finished
//               accumulators.  Also owns TMEM alloc/dealloc.  Both issue warps stay converged and
//               issue under elect.sync so that descriptors live in uniform registers.
//   warps 0..7  epilogue: each thread owns ONE query row (TMEM lane) and every other 32-column chunk
//               of the tile (two interleaved column sets, so runs of consecutive bank rows -- the
//               patches of one training image -- are shared out between the two lists) and keeps its k'/2 best candidates as a sorted list (scores in registers, bank
//               rows in shared memory).  Per 32 columns: tcgen05.ld, max-trees over groups of 8, one
//               warp-wide OR against tau = the list's last score.  Survivors (rare) go to a 16-entry
//               per-thread queue in shared memory with independent predicated stores; when some
//               lane's queue is nearly full the whole warp folds queues into lists with fully
//               unrolled bitonic networks on registers, in lock-step (no divergence), so the fold
//               is amortised over the warp's 32 queries.  The TMEM buffer is handed back as soon
//               as its last chunk is in registers; accumulator double-buffering overlaps the scan
//               with the MMAs of the next tile.  The epilogue must not spill: with 227 KB of shared
//               memory carved out there is no L1 behind local memory.
// Work item = (query block of 128*CG rows) x (bank chunk of consecutive 256-row tiles); items are
// dealt round-robin, chunk-major, to the persistent CTAs so that concurrently running CTAs walk
// the same bank tiles (L2 reuse) with different queries.  Each item emits k' unsorted candidate
// keys per query; thresholds are shared between the lists of one query through global memory.
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"
#include "search_plan.h"

namespace hb {

constexpr int BM = 128;   // query rows per CTA (TMEM lanes)
constexpr int BN = 256;   // bank rows per tile (UMMA N)
constexpr int BK = 64;    // bf16 per k-block: one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;   // 4 TMEM lane quarters x 2 column halves
constexpr int kQueueCap = 16;  // pending-candidate queue entries per epilogue thread
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kChunk = 32;     // accumulator columns per tcgen05.ld
constexpr int kGroups = kChunk / 8;
constexpr int kSearchThreads = 32 * (2 + kEpiWarps);
constexpr int kProducerWarp = kEpiWarps;      // warp 8
constexpr int kMmaWarp = kEpiWarps + 1;       // warp 9
constexpr int kTmemCols = 512;
// L2 pacing: CTAs that stream the same bank chunk may not run more than (kPaceLag + 1) groups of
// kPaceTiles tiles ahead of the slowest one, so a bank tile is fetched from HBM once per wave and
// then hit in the 126 MB L2 by everyone else.
constexpr int kPaceTiles = 32;
constexpr int kPaceLag = 2;
constexpr int64_t kSmallBankRows = 16384;  // banks up to this size always run with k' = 128 lists
constexpr int kBoardPeriod = 8;  // tiles between two reads of the threshold board (power of two)
constexpr int kBoardDense = 24;  // ... after the first kBoardDense tiles of a work item, which read it every tile

struct SearchParams {
  int64_t n_rows;      // valid bank rows
  int64_t n_queries;   // valid query rows
  int num_kblocks;     // dpad / 64
  int n_qblocks;       // query blocks of 128*CG rows
  int n_chunks;        // bank chunks per query block
  int n_tiles;         // ceil(n_rows / 256)
  uint64_t* cand;      // (n_chunks, n_qblocks*128*CG, KP) candidate keys
  uint32_t* board;     // (2, 2*n_chunks, q_pad) published list statistics, ordered-float bits (0 = none yet):
                       // plane 0 = every list's KL-th best, plane 1 = its (KL/2)-th best (see the epilogue)
  float* dump;         // optional (n_queries, n_rows) raw scores (validation only)
  unsigned long long* stats;  // optional (CTAs, 8 warps, 8) cycle counters (instrumented build only)
  int prefetch_tiles;  // > 0: L2-prefetch bank tiles this many tiles ahead (split over the CTAs)
  uint32_t* pace;      // (rounds, pace_groups) arrival counters of the L2 pacing window (NULL = off)
  int pace_groups;     // counters per round
  int ablate;          // measurement only (selects an ablated kernel build): 1 = release accumulators
                       // unread, 2 = scan but never insert
};

template <int CG, int STAGES, int KP>
struct SearchSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / CG) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRingBytes = STAGES * kStageBytes;
  static constexpr int kQueueBytes = kQueueCap * kEpiThreads * 8;
  static constexpr int kRowsBytes = (KP / 2) * kEpiThreads * 4;  // bank rows of the list entries
  static constexpr int kBarOffset = kRingBytes + kQueueBytes + kRowsBytes;
  static constexpr int kNumBars = 2 * STAGES + 4;
  static constexpr int kTotal = kBarOffset + kNumBars * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;  // slack to align the ring to 1024 B
  static_assert(kDynamic <= 227 * 1024, "search kernel shared memory exceeds 227 KB");
};

// ---- per-thread candidate list in registers ---------------------------------------------------
// s[] scores / r[] bank rows, every index a compile-time constant so the arrays never leave the
// register file.  All lanes of a warp run these networks in lock-step on their own lists.
template <bool DESC>
__device__ __forceinline__ void cmpx(float& sa, uint32_t& ra, float& sb, uint32_t& rb) {
  // after the call (a, b) is ordered: descending if DESC else ascending
  const bool sw = DESC ? (sa < sb) : (sa > sb);
  const float ts = sw ? sb : sa;
  const uint32_t tr = sw ? rb : ra;
  sb = sw ? sa : sb;
  rb = sw ? ra : rb;
  sa = ts;
  ra = tr;
}
// Bitonic merge of s[LO .. LO+N): input bitonic, output sorted (DESC or ascending).
template <int N, int LO, bool DESC, int M>
__device__ __forceinline__ void reg_bitonic_merge(float (&s)[M], uint32_t (&r)[M]) {
#pragma unroll
  for (int j = N / 2; j > 0; j >>= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      if ((i & j) == 0) cmpx<DESC>(s[LO + i], r[LO + i], s[LO + i + j], r[LO + i + j]);
  }
}
// Full bitonic sort of s[0 .. N).
template <int N, bool DESC, int M>
__device__ __forceinline__ void reg_bitonic_sort(float (&s)[M], uint32_t (&r)[M]) {
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if ((i & j) == 0) {
          const bool up = ((i & k) == 0) == DESC;  // direction of this sub-sequence
          if (up) cmpx<true>(s[i], r[i], s[i + j], r[i + j]);
          else cmpx<false>(s[i], r[i], s[i + j], r[i + j]);
        }
      }
    }
  }
}
// Fold `cnt` queued (score,row) pairs (shared memory, entry j at queue[j*kEpiThreads]) into the
// sorted-descending list of KL entries.  Between folds only the scores stay in registers (the scan
// needs nothing but the last one); the bank rows of the entries rest in shared memory
// (rows[j*kEpiThreads]) and visit registers for the duration of the merge network.  A spill would be
// ruinous here: with 227 KB of shared memory carved out there is next to no L1 behind local memory.
template <int KL, int QC>
__device__ __forceinline__ void fold_queue(float (&ls)[KL], uint32_t* rows, const uint2* queue, int cnt) {
  constexpr int QB = 8;  // queue entries folded per pass: bounds the registers the merge needs
  uint32_t lr[KL];
#pragma unroll
  for (int j = 0; j < KL; ++j) lr[j] = rows[j * kEpiThreads];
#pragma unroll 1
  for (int base = 0; base < QC; base += QB) {
    if (base > 0 && !__any_sync(0xffffffffu, cnt > base)) break;
    float qs[QB];
    uint32_t qr[QB];
#pragma unroll
    for (int j = 0; j < QB; ++j) {
      const uint2 e = queue[(base + j) * kEpiThreads];
      const bool ok = base + j < cnt;
      qs[j] = ok ? __uint_as_float(e.x) : -INFINITY;
      qr[j] = ok ? e.y : 0xffffffffu;
    }
    reg_bitonic_sort<QB, false>(qs, qr);  // ascending
    // the QB smallest list entries (descending) against the ascending batch: the element-wise max
    // keeps the QB largest of their union, as a bitonic sequence
#pragma unroll
    for (int j = 0; j < QB; ++j) {
      const bool take = qs[j] > ls[KL - QB + j];
      ls[KL - QB + j] = take ? qs[j] : ls[KL - QB + j];
      lr[KL - QB + j] = take ? qr[j] : lr[KL - QB + j];
    }
    // tail ascending: [descending head | ascending tail] is bitonic, one merge sorts the list
    reg_bitonic_merge<QB, KL - QB, false>(ls, lr);
    reg_bitonic_merge<KL, 0, true>(ls, lr);
  }
#pragma unroll
  for (int j = 0; j < KL; ++j) rows[j * kEpiThreads] = lr[j];
}

// LB = the thread count the register budget is derived from (__launch_bounds__): kSearchThreads gives
// ptxas the whole register file (168 registers per thread); 512 caps it at 128, which leaves room on
// every SM sub-partition for a warp of another kernel (see hb_coresidency_config).  The launch is
// always kSearchThreads wide.
template <int CG, int STAGES, int KP, int MODE, int LB>
__global__ void __launch_bounds__(LB, 1)
search_topk_kernel(const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ CUtensorMap tmap_bank, const SearchParams p) {
  using L = SearchSmem<CG, STAGES, KP>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;

  const uint32_t bar_base = smem_base + L::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kBarOffset + L::kNumBars * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int cluster_id = (CG == 2) ? static_cast<int>(ptx::cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int n_clusters = (CG == 2) ? static_cast<int>(ptx::num_clusters_x()) : static_cast<int>(gridDim.x);

  // ---- one-time setup ----
  if (warp == kProducerWarp && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_bank);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(tfull_bar(b), 1);
      ptx::mbar_init(tempty_bar(b), kEpiWarps * CG);
    }
    ptx::fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    ptx::tmem_alloc<CG>(smem_u32(tmem_slot), kTmemCols);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_items = p.n_qblocks * p.n_chunks;
  const int nkb = p.num_kblocks;

  if (warp == kProducerWarp) {
    // =========================== TMA producer (converged warp, one elected lane issues) ===========================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < total_items; item += n_clusters) {
        const int qb = item % p.n_qblocks, chunk = item / p.n_qblocks;
        const int t0 = chunk_tile_begin(p.n_tiles, p.n_chunks, chunk);
        const int t1 = chunk_tile_begin(p.n_tiles, p.n_chunks, chunk + 1);
        const int32_t q_row = (qb * CG + static_cast<int>(cta_rank)) * BM;
        const int round = item / n_clusters;
        const uint32_t n_part = static_cast<uint32_t>(min(n_clusters, total_items - round * n_clusters));
        uint32_t* pace = p.pace ? p.pace + static_cast<size_t>(round) * p.pace_groups : nullptr;
        for (int tile = t0; tile < t1; ++tile) {
          const int32_t b_row = tile * BN + static_cast<int>(cta_rank) * (BN / CG);
          const bool pf_tile = p.prefetch_tiles > 0 && tile + p.prefetch_tiles < t1;
          const int rel = tile - t0;
          if (pace != nullptr && (rel % kPaceTiles) == 0 && rel / kPaceTiles > kPaceLag) {
            // do not start group g before every CTA of this wave has loaded group g - 1 - kPaceLag
            const volatile uint32_t* ctr = pace + (rel / kPaceTiles - 1 - kPaceLag);
            const uint64_t t_start = ptx::globaltimer_ns();
            while (*ctr < n_part) {
              __nanosleep(256);
              if (ptx::globaltimer_ns() - t_start > 4000000000ull) ptx::hang_trap(5, 0, 0);
            }
          }
          for (int kb = 0; kb < nkb; ++kb) {
            ptx::mbar_wait(empty_bar(stage), phase ^ 1u, 1);
            if (ptx::elect_one()) {
              const uint32_t a_dst = smem_base + stage * L::kStageBytes;
              // completion bytes of both CTAs of a pair land on the pair leader's barrier
              const uint32_t full_addr = (CG == 2) ? ptx::mapa(full_bar(stage), 0) : full_bar(stage);
              if (is_leader) ptx::mbar_arrive_expect_tx(full_bar(stage), L::kStageBytes * CG);
              ptx::tma_load_2d<CG>(a_dst, &tmap_q, full_addr, kb * BK, q_row);
              ptx::tma_load_2d<CG>(a_dst + L::kABytes, &tmap_bank, full_addr, kb * BK, b_row);
              // bank tiles are shared by all CTAs walking this chunk: each (tile, k-block) box is
              // pulled into L2 ahead of time by exactly one of them
              if (pf_tile && (tile * nkb + kb) % n_clusters == cluster_id)
                ptx::tma_prefetch_2d(&tmap_bank, kb * BK, b_row + p.prefetch_tiles * BN);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          if (pace != nullptr && is_leader && lane == 0 && ((rel % kPaceTiles) == kPaceTiles - 1 || tile == t1 - 1))
            atomicAdd(pace + rel / kPaceTiles, 1u);  // this CTA (pair) has issued all loads of the group
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer (pair leader; converged warp, one elected lane issues) ===========================
    // The issue loop must stay well under the 512 tensor-pipe cycles one k-block is worth: every
    // value feeding the descriptors is warp-uniform, so they live in uniform registers.
    if (is_leader) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BM * CG, BN);
      const uint64_t adesc0 = ptx::make_smem_desc_sw128(smem_base);
      const uint64_t bdesc0 = ptx::make_smem_desc_sw128(smem_base + L::kABytes);
      constexpr uint64_t kStageStep = L::kStageBytes >> 4;  // descriptor address field is in 16-byte units
      int stage = 0;
      uint32_t phase = 0;
      int abuf = 0;
      uint32_t aphase = 0;
      for (int item = cluster_id; item < total_items; item += n_clusters) {
        const int chunk = item / p.n_qblocks;
        const int t0 = chunk_tile_begin(p.n_tiles, p.n_chunks, chunk);
        const int t1 = chunk_tile_begin(p.n_tiles, p.n_chunks, chunk + 1);
        for (int tile = t0; tile < t1; ++tile) {
          ptx::mbar_wait(tempty_bar(abuf), aphase ^ 1u, 2);  // epilogue drained this accumulator
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(abuf * BN);
          for (int kb = 0; kb < nkb; ++kb) {
            ptx::mbar_wait(full_bar(stage), phase, 3);  // TMA bytes have landed
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint64_t adesc = adesc0 + kStageStep * stage;
              const uint64_t bdesc = bdesc0 + kStageStep * stage;
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                // advance 16 bf16 = 32 B inside the swizzle atom: +2 in the (addr >> 4) field
                ptx::umma_bf16<CG>(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc,
                                   static_cast<uint32_t>((kb | k) != 0));
              }
              ptx::umma_commit<CG>(empty_bar(stage));  // frees the ring slot when the MMAs retire
              if (kb == nkb - 1) ptx::umma_commit<CG>(tfull_bar(abuf));  // accumulator complete
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          abuf ^= 1;
          if (abuf == 0) aphase ^= 1u;
        }
      }
    }
  } else if (warp < kEpiWarps) {
    // =========================== epilogue: fused top-k' ===========================
    constexpr int KL = KP / 2;                    // list length per (row, column set)
    constexpr int QC = kQueueCap;
    constexpr uint32_t kQStride = kEpiThreads * 8;  // bytes between consecutive queue entries of a thread
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = warp >> 2;                   // which of the two per-row lists this thread keeps
    const int row_in_tile = quarter * 32 + lane;  // query row owned by this thread
    const uint2* queue = reinterpret_cast<const uint2*>(smem + L::kRingBytes) + threadIdx.x;  // entry j at [j*kEpiThreads]
    const uint32_t queue_addr = smem_base + L::kRingBytes + threadIdx.x * 8u;
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem + L::kRingBytes + L::kQueueBytes) + threadIdx.x;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t tempty_leader0 = (CG == 2) ? ptx::mapa(tempty_bar(0), 0) : tempty_bar(0);
    const uint32_t tempty_leader1 = (CG == 2) ? ptx::mapa(tempty_bar(1), 0) : tempty_bar(1);
    int abuf = 0;
    uint32_t aphase = 0;
    // hand the current TMEM accumulator back to the MMA issuer (the pair leader's barrier)
    auto release_accumulator = [&]() {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        const uint32_t bar = abuf ? tempty_leader1 : tempty_leader0;
        if (CG == 2) ptx::mbar_arrive_cluster(bar); else ptx::mbar_arrive_local(bar);
      }
    };
    const int64_t q_pad = static_cast<int64_t>(p.n_qblocks) * BM * CG;
    const int n_slots = 2 * p.n_chunks;
    for (int item = cluster_id; item < total_items; item += n_clusters) {
      const int qb = item % p.n_qblocks, chunk = item / p.n_qblocks;
      const int t0 = chunk_tile_begin(p.n_tiles, p.n_chunks, chunk);
      const int t1 = chunk_tile_begin(p.n_tiles, p.n_chunks, chunk + 1);
      const int64_t q_row = static_cast<int64_t>(qb * CG + static_cast<int>(cta_rank)) * BM + row_in_tile;
      float ls[KL];
#pragma unroll
      for (int j = 0; j < KL; ++j) {
        ls[j] = -INFINITY;
        rows[j * kEpiThreads] = 0xffffffffu;  // "no candidate"
      }
      uint32_t qtop = queue_addr;  // shared-memory address of this thread's next free queue entry
      long long st_wait = 0, st_load = 0, st_slow = 0, st_fold = 0, st_nfold = 0, st_nslow = 0, st_tiles = 0, t_a = 0;
      // Threshold sharing through the board.  Every list that scans bank rows for this query (2 per
      // chunk, chunks on different CTAs, at the same time or one after the other) publishes two
      // statistics after each fold: its KL-th best score (plane 0) and its (KL/2)-th best (plane 1).
      // If two lists each hold KL scores >= x, or four lists each hold KL/2 scores >= x, then the
      // bank holds 2*KL = k' rows scoring >= x, so a row below x is not among the query's best k'.
      // x = max(2nd largest of plane 0, 4th largest of plane 1) over a window of up to 8 lists is
      // therefore a threshold every list may prune with.  Together with a list's own KL-th best:
      // a row of the query's bf16 top-k' is dropped only if KL better rows share its list.
      const int slot = 2 * chunk + half;
      int w0 = 2 * chunk - 4;
      if (w0 > n_slots - 8) w0 = n_slots - 8;
      if (w0 < 0) w0 = 0;
      const int w1 = (w0 + 8 < n_slots) ? w0 + 8 : n_slots;
      auto board_bound = [&]() -> float {
        const uint32_t* brd = p.board + static_cast<int64_t>(w0) * q_pad + q_row;
        const int64_t plane = static_cast<int64_t>(n_slots) * q_pad;
        uint32_t x[8], y[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const bool in = w0 + s < w1;  // warp-uniform; a slot must not be counted twice
          x[s] = in ? __ldcg(brd + s * q_pad) : 0u;
          y[s] = in ? __ldcg(brd + plane + s * q_pad) : 0u;
        }
        uint32_t f1 = 0, f2 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          f2 = max(f2, min(f1, x[s]));
          f1 = max(f1, x[s]);
          const uint32_t a = min(h1, y[s]);
          h1 = max(h1, y[s]);
          const uint32_t b = min(h2, a);
          h2 = max(h2, a);
          const uint32_t c = min(h3, b);
          h3 = max(h3, b);
          h4 = max(h4, c);
        }
        const uint32_t bound = max(f2, h4);
        return bound ? ordered_to_f32(bound) : -INFINITY;
      };
      auto publish = [&]() {
        uint32_t* pub = p.board + static_cast<int64_t>(slot) * q_pad + q_row;
        __stcg(pub, f32_to_ordered(ls[KL - 1]));
        __stcg(pub + static_cast<int64_t>(n_slots) * q_pad, f32_to_ordered(ls[KL / 2 - 1]));
      };
      float shared_tau = -INFINITY;  // best bound read from the board so far
      float tau = -INFINITY;         // max(shared_tau, own KL-th best)
      for (int tile = t0; tile < t1; ++tile) {
        // every tile while the lists of this query are young (their statistics move fast and short
        // chunks are over before a sparse schedule pays), every kBoardPeriod tiles afterwards
        if ((tile - t0 < kBoardDense || ((tile - t0) & (kBoardPeriod - 1)) == 0) && MODE != 3 && MODE != 4) {
          shared_tau = fmaxf(shared_tau, board_bound());
          tau = fmaxf(tau, shared_tau);
        }
        if (MODE == 2) t_a = clock64();
        ptx::mbar_wait(tfull_bar(abuf), aphase, 4);
        ptx::tc_fence_after();
        if (MODE == 2) { st_wait += clock64() - t_a; ++st_tiles; }
        const int64_t col_base = static_cast<int64_t>(tile) * BN;
        const int64_t rem = p.n_rows - col_base;
        const int nvalid = rem >= BN ? BN : static_cast<int>(rem);  // valid columns of this tile (>= 1)
        const uint32_t tacc = tmem_lane + static_cast<uint32_t>(abuf * BN);
        // the two lists of a row take alternate 32-column chunks of a tile, and which list takes the
        // even ones changes pseudo-randomly from tile to tile: bank rows with a regular stride (the
        // same patch position in consecutive images) do not pile up in one list
        const int c_first = (half ^ static_cast<int>((static_cast<uint32_t>(tile) * 0x9E3779B1u) >> 31)) * kChunk;
#pragma unroll 1
        for (int c0 = c_first; c0 < BN; c0 += 2 * kChunk) {  // this warp's interleaved column set
          if (c0 >= nvalid || MODE == 3) break;  // warp-uniform
          uint32_t v[kChunk];
          if (MODE == 2) t_a = clock64();
          ptx::tmem_ld_chunk(tacc + c0, v);
          ptx::tmem_ld_wait();
          if (MODE == 2) st_load += clock64() - t_a;
          const bool last_chunk = c0 + 2 * kChunk >= nvalid || c0 + 2 * kChunk >= BN;  // no further chunk of this set
          if (MODE == 1 && q_row < p.n_queries) {
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
              if (c0 + j < nvalid) p.dump[q_row * p.n_rows + col_base + c0 + j] = __uint_as_float(v[j]);
          }
          const int nv = nvalid - c0;  // valid columns of this chunk (may exceed kChunk)
          if (nv < kChunk) {  // last, partial tile of the bank: padded rows never win
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
              if (j >= nv) v[j] = 0xff800000u;  // -inf
          }
          // fast reject: per-group maxima against tau, one warp-wide OR of the 4-bit result
          float mg[kGroups];
#pragma unroll
          for (int g = 0; g < kGroups; ++g) {
            float m = __uint_as_float(v[8 * g]);
#pragma unroll
            for (int j = 1; j < 8; ++j) m = fmaxf(m, __uint_as_float(v[8 * g + j]));
            mg[g] = m;
          }
          uint32_t gmask = 0;
#pragma unroll
          for (int g = 0; g < kGroups; ++g) gmask |= (mg[g] > tau ? 1u : 0u) << g;
          uint32_t hot = __reduce_or_sync(0xffffffffu, gmask);  // groups in which some lane has a survivor
          if (MODE == 4) hot = 0;
          if (hot == 0) {
            if (last_chunk) release_accumulator();
          } else {
            if (MODE == 2) { t_a = clock64(); ++st_nslow; }
            // Survivors are appended to the thread's queue with predicated stores through a running
            // address.  A group adds at most 8 entries per lane; if some lane has fewer than 8 free
            // the warp folds first -- that needs the registers v occupies, so the chunk is re-read
            // from TMEM afterwards (the accumulator is still ours) and appending resumes there.
            const uint32_t code0 = static_cast<uint32_t>(col_base) + static_cast<uint32_t>(c0);
            int resume = 0;
            for (;;) {
              int blocked = -1;
#pragma unroll
              for (int g = 0; g < kGroups; ++g) {
                if (blocked < 0 && g >= resume && ((hot >> g) & 1u)) {
                  if (__any_sync(0xffffffffu, qtop > queue_addr + (QC - 8) * kQStride)) {
                    blocked = g;
                  } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      if (__uint_as_float(v[8 * g + j]) > tau) {
                        // the bank row is formed here, under the predicate, and not hoisted to the
                        // entry of the slow path for all 32 columns (most groups are not hot)
                        ptx::st_shared_v2(qtop, v[8 * g + j], ptx::add_volatile(code0, static_cast<uint32_t>(8 * g + j)));
                        qtop += kQStride;
                      }
                    }
                  }
                }
              }
              if (blocked < 0) break;
              if (MODE == 2) ++st_nfold;
              fold_queue<KL, QC>(ls, rows, queue, static_cast<int>((qtop - queue_addr) / kQStride));  // v is dead here
              qtop = queue_addr;
              publish();
              tau = fmaxf(shared_tau, ls[KL - 1]);
              resume = blocked;
              ptx::tmem_ld_chunk(tacc + c0, v);
              ptx::tmem_ld_wait();
              if (nv < kChunk) {
#pragma unroll
                for (int j = 0; j < kChunk; ++j)
                  if (j >= nv) v[j] = 0xff800000u;
              }
            }
            if (last_chunk) release_accumulator();
            if (MODE == 2) st_slow += clock64() - t_a;
          }
        }
        if (c_first >= nvalid || MODE == 3) release_accumulator();  // nothing was read
        // routine folds happen here, after the accumulator was released, whenever some lane could not
        // take another full group: a fold costs ~1k cycles of this warp, and the MMA of the tile after
        // next waits for the slowest of all epilogue warps
        if (__any_sync(0xffffffffu, qtop > queue_addr + (QC - 8) * kQStride)) {
          if (MODE == 2) { t_a = clock64(); ++st_nfold; }
          fold_queue<KL, QC>(ls, rows, queue, static_cast<int>((qtop - queue_addr) / kQStride));
          qtop = queue_addr;
          publish();
          tau = fmaxf(shared_tau, ls[KL - 1]);
          if (MODE == 2) st_fold += clock64() - t_a;
        }
        abuf ^= 1;
        if (abuf == 0) aphase ^= 1u;
      }
      if (__any_sync(0xffffffffu, qtop > queue_addr)) {
        fold_queue<KL, QC>(ls, rows, queue, static_cast<int>((qtop - queue_addr) / kQStride));
        publish();  // lists that start later begin with this one's final statistics
      }
      if (MODE == 2 && lane == 0) {
        unsigned long long* o = p.stats + (static_cast<size_t>(blockIdx.x) * kEpiWarps + warp) * 8;
        atomicAdd(o + 0, static_cast<unsigned long long>(st_wait));
        atomicAdd(o + 1, static_cast<unsigned long long>(st_load));
        atomicAdd(o + 2, static_cast<unsigned long long>(st_slow));
        atomicAdd(o + 3, static_cast<unsigned long long>(st_fold));
        atomicAdd(o + 4, static_cast<unsigned long long>(st_nslow));
        atomicAdd(o + 5, static_cast<unsigned long long>(st_nfold));
        atomicAdd(o + 6, static_cast<unsigned long long>(st_tiles));
      }
      // emit this item's candidates (k'/2 per list; the re-rank kernel merges them)
      uint64_t* out = p.cand + (static_cast<int64_t>(chunk) * q_pad + q_row) * KP + half * KL;
#pragma unroll
      for (int j = 0; j < KL; ++j) {
        const uint32_t r = rows[j * kEpiThreads];
        out[j] = (r == 0xffffffffu) ? 0ull : make_key(ls[j], r);
      }
    }
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

// Query preparation: fp32 (Q, d) -> bf16 (Q, dpad) zero padded, plus ||q||_2.
// (hbird_eval.py:624-625 hands the raw, un-normalised query rows to the backend.)
__global__ void __launch_bounds__(256)
prep_queries_kernel(const float* __restrict__ q, int64_t Q, int d, int dpad, int l2,
                    __nv_bfloat16* __restrict__ qb, float* __restrict__ qnorm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d4 = d >> 2;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp; row < Q;
       row += static_cast<int64_t>(gridDim.x) * 8) {
    const float4* in4 = reinterpret_cast<const float4*>(q + row * d);
    __nv_bfloat16* ob = qb + row * dpad;
    float ss = 0.f;
    for (int i = lane; i < d4; i += 32) {
      const float4 v = __ldg(in4 + i);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
      __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(ob + 4 * i) = pk;
    }
    // L2 banks carry -||x||^2/2 in columns d..d+2 (bank.cu); the query multiplies them by 1
    for (int i = d + lane; i < dpad; i += 32) ob[i] = __float2bfloat16((l2 && i < d + 3) ? 1.f : 0.f);
    ss = warp_sum(ss);
    if (lane == 0 && qnorm) qnorm[row] = sqrtf(ss);
  }
}

// Launch (or, with probe != nullptr, only size) the persistent search grid.  The kernel's CTAs wait
// for each other (pair barriers, L2 pacing), so every CTA must be resident at once: the grid is
// min(SMs / CG, work items, what the driver says fits) clusters.
template <int CG, int STAGES, int KP, int MODE = 0, int LB = kSearchThreads>
static int launch_search(const Bank* b, const CUtensorMap& tmap_q, const SearchParams& p,
                         cudaStream_t st, int* probe, int n_clusters) {
  using L = SearchSmem<CG, STAGES, KP>;
  auto kern = search_topk_kernel<CG, STAGES, KP, MODE, LB>;
  HB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic));
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kSearchThreads);
  cfg.dynamicSmemBytes = L::kDynamic;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (probe != nullptr) {
    int want = std::max(1, std::min(b->num_sms / CG, p.n_qblocks * p.n_chunks));
    cfg.gridDim = dim3(static_cast<unsigned>(want * CG));
    int fit = 0;
    if (cudaOccupancyMaxActiveClusters(&fit, kern, &cfg) == cudaSuccess && fit > 0) want = std::min(want, fit);
    else (void)cudaGetLastError();
    *probe = want;
    return HB_OK;
  }
  cfg.gridDim = dim3(static_cast<unsigned>(n_clusters * CG));
  const CUtensorMap& tmap_b = (CG == 2) ? b->tmap_bank_cg2 : b->tmap_bank_cg1;
  HB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap_q, tmap_b, p));
  return HB_OK;
}

// Dispatch over (cta_group, k').  Ring depth is what fits beside the queues in 227 KB.
static int dispatch_search(const Bank* b, int cg, int kp, const CUtensorMap& tmap_q,
                           const SearchParams& p, cudaStream_t st, int* probe, int n_clusters) {
  if (p.dump != nullptr) {  // validation build of the same kernel that also writes the raw scores
    if (cg == 2) return launch_search<2, 5, 64, 1>(b, tmap_q, p, st, probe, n_clusters);
    return launch_search<1, 3, 64, 1>(b, tmap_q, p, st, probe, n_clusters);
  }
  if (p.ablate == 1) return cg == 2 ? launch_search<2, 5, 64, 3>(b, tmap_q, p, st, probe, n_clusters)
                                    : launch_search<1, 3, 64, 3>(b, tmap_q, p, st, probe, n_clusters);
  if (p.ablate == 2) return cg == 2 ? launch_search<2, 5, 64, 4>(b, tmap_q, p, st, probe, n_clusters)
                                    : launch_search<1, 3, 64, 4>(b, tmap_q, p, st, probe, n_clusters);
  if (p.stats != nullptr) {  // instrumented build: per-warp cycle counters of the epilogue
    if (cg == 2) return launch_search<2, 5, 64, 2>(b, tmap_q, p, st, probe, n_clusters);
    return launch_search<1, 3, 64, 2>(b, tmap_q, p, st, probe, n_clusters);
  }
  if (cg == 2 && kp == 64 && b->cfg_lean_search)  // 128-register build of the default configuration
    return launch_search<2, 5, 64, 0, 512>(b, tmap_q, p, st, probe, n_clusters);
  if (cg == 2) {
    if (kp == 32) return launch_search<2, 5, 32>(b, tmap_q, p, st, probe, n_clusters);
    if (kp == 64) return launch_search<2, 5, 64>(b, tmap_q, p, st, probe, n_clusters);
    if (kp == 128) return launch_search<2, 4, 128>(b, tmap_q, p, st, probe, n_clusters);
  } else {
    if (kp == 32) return launch_search<1, 3, 32>(b, tmap_q, p, st, probe, n_clusters);
    if (kp == 64) return launch_search<1, 3, 64>(b, tmap_q, p, st, probe, n_clusters);
    if (kp == 128) return launch_search<1, 2, 128>(b, tmap_q, p, st, probe, n_clusters);
  }
  set_error("hb_search: k_prime=%d not in {32, 64, 128}", kp);
  return HB_ERR_INVALID;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Cross-stream slot hand-over uses events recorded outside of any capture; inside a CUDA-graph capture
// the whole search sits on the captured stream and needs (and may use) none of them.
static bool stream_is_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &status) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return status != cudaStreamCaptureStatusNone;
}

int search_begin_impl(Bank* b, const float* q, int64_t Q, int kp, int slot, float* out_qnorm, float* dump,
                      int cg_override, cudaStream_t st, cudaEvent_t prepared) {
  // measured on B200 (profiles/): CTA pairs (cta_group::2: half the B-operand shared-memory traffic
  // per SM) win at every bank size and feature dim; cta_group 1 stays selectable
  int cg = cg_override ? cg_override : (b->cfg_cta_group ? b->cfg_cta_group : 2);
  if (b->num_sms < 2) cg = 1;
  // Small banks have too few tiles to spread a query's best rows over many lists: keep k'/2 = 64
  // per list there (strict bf16 top-64 per list; the cost is irrelevant at this size).
  if (b->rows <= kSmallBankRows && kp < 128 && dump == nullptr) kp = 128;
  const SearchPlan plan = plan_search(b->rows, Q, cg, b->num_sms, b->cfg_max_chunks);
  const int64_t q_pad = static_cast<int64_t>(plan.n_qblocks) * BM * cg;
  PipeSlot& ps = b->pipe[slot];
  // the slot's buffers may still be read by the finish of the search that used it last
  const bool capturing = stream_is_capturing(st);
  if (ps.done != nullptr && !capturing) HB_CHECK_CUDA(cudaStreamWaitEvent(st, ps.done, 0));

  int n_units_used = 1;
  const int variant = dump ? 1 : (b->cfg_ablate ? 2 + b->cfg_ablate : (b->cfg_stats ? 2 : (b->cfg_lean_search ? 5 : 0)));
  int& cached_fit = b->fit_cache[cg - 1][kp == 128 ? 2 : (kp == 64 ? 1 : 0)][variant];
  const int want_units = std::max(1, std::min(b->num_sms / cg, plan.n_qblocks * plan.n_chunks));
  if (cached_fit > 0) {
    n_units_used = std::min(want_units, cached_fit);
  } else {
    SearchParams probe_p{};
    probe_p.dump = dump;
    probe_p.stats = b->cfg_stats;
    probe_p.ablate = b->cfg_ablate;
    CUtensorMap unused{};
    // probe with a grid that wants every SM, so the cached answer is the device's capacity
    probe_p.n_qblocks = b->num_sms;
    probe_p.n_chunks = 1;
    int rc0 = dispatch_search(b, cg, kp, unused, probe_p, st, &cached_fit, 0);
    if (rc0 != HB_OK) return rc0;
    n_units_used = std::min(want_units, cached_fit);
  }
  // Scratch.  Shared by every search (used by prep + K2 only, which run one after the other on the
  // begin stream): bf16 queries | threshold board (2 planes x 2 lists per chunk x padded queries) |
  // pacing counters.  Per pipeline slot, in a block of its own (read by the slot's finish, possibly on
  // another stream while the next begin already runs): norms | candidate keys.
  const int n_rounds = (plan.n_qblocks * plan.n_chunks + n_units_used - 1) / n_units_used;
  const int pace_groups = (plan.n_tiles / plan.n_chunks + 1) / kPaceTiles + 2;
  const size_t off_q = 0;
  const size_t off_board = align_up(off_q + sizeof(__nv_bfloat16) * static_cast<size_t>(Q) * b->dpad, 256);
  const size_t board_bytes = sizeof(uint32_t) * 4 * static_cast<size_t>(plan.n_chunks) * q_pad;
  const size_t off_pace = align_up(off_board + board_bytes, 256);
  const size_t total = align_up(off_pace + sizeof(uint32_t) * static_cast<size_t>(n_rounds) * pace_groups, 256);
  int rc = ensure_workspace(b, total, st);
  if (rc != HB_OK) return rc;
  const size_t norm_bytes = align_up(sizeof(float) * static_cast<size_t>(Q), 256);
  const size_t slot_bytes = norm_bytes + align_up(sizeof(uint64_t) * static_cast<size_t>(plan.n_chunks) * q_pad * kp, 256);
  if (slot_bytes > ps.buf_bytes) {
    // stream-ordered: `st` already waits for the slot's last finish (ps.done above)
    if (ps.buf) HB_CHECK_CUDA(cudaFreeAsync(ps.buf, st));
    ps.buf = nullptr;
    ps.buf_bytes = 0;
    HB_CHECK_CUDA(cudaMallocAsync(&ps.buf, slot_bytes + slot_bytes / 4, st));
    ps.buf_bytes = slot_bytes + slot_bytes / 4;
  }
  uint8_t* ws = static_cast<uint8_t*>(b->ws);
  __nv_bfloat16* q_bf16 = reinterpret_cast<__nv_bfloat16*>(ws + off_q);
  uint32_t* board = reinterpret_cast<uint32_t*>(ws + off_board);
  uint32_t* pace = reinterpret_cast<uint32_t*>(ws + off_pace);
  uint8_t* slot_base = static_cast<uint8_t*>(ps.buf);
  float* qnorm = out_qnorm ? out_qnorm : reinterpret_cast<float*>(slot_base);
  uint64_t* cand = reinterpret_cast<uint64_t*>(slot_base + norm_bytes);
  // board and pacing counters are contiguous: one memset
  HB_CHECK_CUDA(cudaMemsetAsync(board, 0, total - off_board, st));

  b->last_launches = 0;
  int64_t blocks = std::min<int64_t>(ceil_div64(Q, 8), static_cast<int64_t>(b->num_sms) * 8);
  prep_queries_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(q, Q, b->d, b->dpad, (b->flags & HB_BANK_L2) ? 1 : 0, q_bf16, qnorm);
  HB_CHECK_CUDA(cudaGetLastError());
  b->last_launches++;
  // A pipelined caller holds the previous batch's post-processing back until this point: released
  // earlier, its small CTAs flood the SMs in the gap before the search kernel is launched and the
  // search CTAs (which need a whole SM each) then wait for them to drain; released here, both are
  // pending together and the search's higher stream priority places it first.
  if (prepared != nullptr) HB_CHECK_CUDA(cudaEventRecord(prepared, st));

  CUtensorMap tmap_q;
  rc = make_tmap_2d_bf16(&tmap_q, q_bf16, Q, b->dpad, BM);
  if (rc != HB_OK) return rc;

  SearchParams p;
  p.n_rows = b->rows;
  p.n_queries = Q;
  p.num_kblocks = b->dpad / BK;
  p.n_qblocks = plan.n_qblocks;
  p.n_chunks = plan.n_chunks;
  p.n_tiles = plan.n_tiles;
  p.cand = cand;
  p.board = board;
  p.pace = b->cfg_pace ? pace : nullptr;
  p.pace_groups = pace_groups;
  p.dump = dump;
  p.stats = b->cfg_stats;
  p.prefetch_tiles = b->cfg_prefetch_tiles >= 0 ? b->cfg_prefetch_tiles : (cg == 2 ? 4 : 0);
  p.ablate = b->cfg_ablate;
  const int tslot = b->timing_count & 63;
  if (b->timing) HB_CHECK_CUDA(cudaEventRecord(b->ev_begin[tslot], st));
  rc = dispatch_search(b, cg, kp, tmap_q, p, st, nullptr, n_units_used);
  if (rc != HB_OK) return rc;
  ps.timing_slot = -1;
  if (b->timing) {
    HB_CHECK_CUDA(cudaEventRecord(b->ev_end[tslot], st));
    ps.timing_slot = tslot;
    b->timing_count++;
  }
  b->last_launches++;
  ps.begun = dump == nullptr;
  ps.kp = kp;
  ps.n_chunks = plan.n_chunks;
  ps.Q = Q;
  ps.q_pad = q_pad;
  ps.cand = cand;
  ps.qnorm = qnorm;
  return HB_OK;
}

int search_finish_impl(Bank* b, int slot, const float* q, int k, int64_t idx_offset, float* out_scores,
                       int64_t* out_idx, const Scatter* sc, const LabelOut* lo, cudaStream_t st) {
  PipeSlot& ps = b->pipe[slot];
  if (!ps.begun) {
    set_error("hb_search_finish: pipeline slot %d holds no begun search (call hb_search_begin first)", slot);
    return HB_ERR_STATE;
  }
  LabelOut label;
  if (lo != nullptr) {
    label = *lo;
    label.qnorm = ps.qnorm;
  }
  if (ps.timing_slot >= 0) HB_CHECK_CUDA(cudaEventRecord(b->ev_rerank0[ps.timing_slot], st));
  int rc = rerank_launch(b, q, ps.Q, k, ps.kp, ps.n_chunks, ps.q_pad, ps.cand, idx_offset, out_scores, out_idx, sc,
                         lo ? &label : nullptr, st);
  if (rc != HB_OK) return rc;
  if (ps.timing_slot >= 0) HB_CHECK_CUDA(cudaEventRecord(b->ev_rerank[ps.timing_slot], st));
  if (!stream_is_capturing(st)) {
    if (ps.done == nullptr) HB_CHECK_CUDA(cudaEventCreateWithFlags(&ps.done, cudaEventDisableTiming));
    HB_CHECK_CUDA(cudaEventRecord(ps.done, st));
  }
  ps.begun = false;
  b->last_launches++;
  return HB_OK;
}

int search_impl(Bank* b, const float* q, int64_t Q, int k, int kp, int64_t idx_offset,
                float* out_scores, int64_t* out_idx, float* out_qnorm, float* dump, int cg_override,
                cudaStream_t st, const Scatter* sc, const LabelOut* lo) {
  int rc = search_begin_impl(b, q, Q, kp, 0, out_qnorm, dump, cg_override, st, nullptr);
  if (rc != HB_OK || dump != nullptr) return rc;  // validation call: raw scores only
  return search_finish_impl(b, 0, q, k, idx_offset, out_scores, out_idx, sc, lo, st);
}

}  // namespace hb

using hb::Bank;

extern "C" {

int hb_search(hb_bank_t* bank, const float* q_dev, int64_t Q, int k, int k_prime, int64_t idx_offset,
              float* out_scores_dev, int64_t* out_idx_dev, float* out_qnorm_dev, void* stream) {
  HB_REQUIRE(bank != nullptr, "hb_search: bank is NULL");
  Bank* b = reinterpret_cast<Bank*>(bank);
  if (!b->finalized) {
    hb::set_error("hb_search: bank not finalized (call hb_bank_finalize first)");
    return HB_ERR_STATE;
  }
  HB_REQUIRE(Q >= 0 && Q < (int64_t(1) << 31), "hb_search: Q=%lld out of range", (long long)Q);
  HB_REQUIRE(k >= 1 && k <= k_prime, "hb_search: need 1 <= k (%d) <= k_prime (%d)", k, k_prime);
  HB_REQUIRE(k_prime == 32 || k_prime == 64 || k_prime == 128, "hb_search: k_prime=%d not in {32, 64, 128}", k_prime);
  if (Q == 0) return HB_OK;
  HB_REQUIRE(q_dev && out_scores_dev && out_idx_dev, "hb_search: NULL pointer");
  HB_REQUIRE(b->rows >= 1, "hb_search: the bank is empty");
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  return hb::search_impl(b, q_dev, Q, k, k_prime, idx_offset, out_scores_dev, out_idx_dev, out_qnorm_dev,
                         nullptr, 0, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

static int fill_label_out(const Bank* b, const uint16_t* label_table_dev, int64_t table_rows, float beta,
                          float* out_label_hat_dev, hb::LabelOut* lo, const char* who) {
  HB_REQUIRE(out_label_hat_dev != nullptr, "%s: out_label_hat_dev is NULL", who);
  HB_REQUIRE(beta > 0.f, "%s: beta must be positive", who);
  HB_REQUIRE((b->flags & HB_BANK_L2) == 0, "%s: label transfer needs an inner-product bank (unit-norm rows)", who);
  lo->table = label_table_dev ? label_table_dev : b->label_hist;
  lo->table_rows = label_table_dev ? table_rows : b->rows;
  HB_REQUIRE(lo->table_rows >= 1, "%s: empty label table", who);
  lo->C = b->C;
  lo->pp = b->pp;
  lo->beta = beta;
  lo->out = out_label_hat_dev;
  return HB_OK;
}

int hb_search_transfer(hb_bank_t* bank, const uint16_t* label_table_dev, int64_t table_rows,
                       const float* q_dev, int64_t Q, int k, int k_prime, int64_t idx_offset, float beta,
                       float* out_scores_dev, int64_t* out_idx_dev, float* out_qnorm_dev,
                       float* out_label_hat_dev, void* stream) {
  HB_REQUIRE(bank != nullptr, "hb_search_transfer: bank is NULL");
  Bank* b = reinterpret_cast<Bank*>(bank);
  if (!b->finalized) {
    hb::set_error("hb_search_transfer: bank not finalized (call hb_bank_finalize first)");
    return HB_ERR_STATE;
  }
  HB_REQUIRE(Q >= 0 && Q < (int64_t(1) << 31), "hb_search_transfer: Q=%lld out of range", (long long)Q);
  HB_REQUIRE(k >= 1 && k <= k_prime, "hb_search_transfer: need 1 <= k (%d) <= k_prime (%d)", k, k_prime);
  HB_REQUIRE(k_prime == 32 || k_prime == 64 || k_prime == 128, "hb_search_transfer: k_prime=%d not in {32, 64, 128}", k_prime);
  HB_REQUIRE((out_scores_dev == nullptr) == (out_idx_dev == nullptr), "hb_search_transfer: give both or neither of out_scores/out_idx");
  if (Q == 0) return HB_OK;
  HB_REQUIRE(q_dev != nullptr, "hb_search_transfer: q_dev is NULL");
  HB_REQUIRE(b->rows >= 1, "hb_search_transfer: the bank is empty");
  hb::LabelOut lo;
  int rc = fill_label_out(b, label_table_dev, table_rows, beta, out_label_hat_dev, &lo, "hb_search_transfer");
  if (rc != HB_OK) return rc;
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  return hb::search_impl(b, q_dev, Q, k, k_prime, idx_offset, out_scores_dev, out_idx_dev, out_qnorm_dev,
                         nullptr, 0, static_cast<cudaStream_t>(stream), nullptr, &lo);
}

int hb_search_begin(hb_bank_t* bank, const float* q_dev, int64_t Q, int k_prime, int slot, float* out_qnorm_dev,
                    void* prepared_event, void* stream) {
  HB_REQUIRE(bank != nullptr, "hb_search_begin: bank is NULL");
  Bank* b = reinterpret_cast<Bank*>(bank);
  if (!b->finalized) {
    hb::set_error("hb_search_begin: bank not finalized (call hb_bank_finalize first)");
    return HB_ERR_STATE;
  }
  HB_REQUIRE(slot == 0 || slot == 1, "hb_search_begin: slot=%d not in {0, 1}", slot);
  HB_REQUIRE(Q >= 1 && Q < (int64_t(1) << 31), "hb_search_begin: Q=%lld out of range", (long long)Q);
  HB_REQUIRE(k_prime == 32 || k_prime == 64 || k_prime == 128, "hb_search_begin: k_prime=%d not in {32, 64, 128}", k_prime);
  HB_REQUIRE(q_dev != nullptr, "hb_search_begin: q_dev is NULL");
  HB_REQUIRE(b->rows >= 1, "hb_search_begin: the bank is empty");
  if (b->pipe[slot].begun) {
    hb::set_error("hb_search_begin: pipeline slot %d already holds a begun search (finish it first)", slot);
    return HB_ERR_STATE;
  }
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  return hb::search_begin_impl(b, q_dev, Q, k_prime, slot, out_qnorm_dev, nullptr, 0, static_cast<cudaStream_t>(stream),
                               static_cast<cudaEvent_t>(prepared_event));
}

int hb_search_finish(hb_bank_t* bank, int slot, const float* q_dev, int k, int64_t idx_offset,
                     const uint16_t* label_table_dev, int64_t table_rows, float beta, float* out_scores_dev,
                     int64_t* out_idx_dev, float* out_label_hat_dev, void* stream) {
  HB_REQUIRE(bank != nullptr, "hb_search_finish: bank is NULL");
  Bank* b = reinterpret_cast<Bank*>(bank);
  HB_REQUIRE(slot == 0 || slot == 1, "hb_search_finish: slot=%d not in {0, 1}", slot);
  HB_REQUIRE(q_dev != nullptr, "hb_search_finish: q_dev is NULL");
  HB_REQUIRE((out_scores_dev == nullptr) == (out_idx_dev == nullptr), "hb_search_finish: give both or neither of out_scores/out_idx");
  HB_REQUIRE(out_scores_dev != nullptr || out_label_hat_dev != nullptr, "hb_search_finish: no output requested");
  HB_REQUIRE(k >= 1 && (!b->pipe[slot].begun || k <= b->pipe[slot].kp), "hb_search_finish: k=%d exceeds the k_prime of the begun search", k);
  hb::LabelOut lo;
  if (out_label_hat_dev != nullptr) {
    int rc = fill_label_out(b, label_table_dev, table_rows, beta, out_label_hat_dev, &lo, "hb_search_finish");
    if (rc != HB_OK) return rc;
  }
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  return hb::search_finish_impl(b, slot, q_dev, k, idx_offset, out_scores_dev, out_idx_dev, nullptr,
                                out_label_hat_dev ? &lo : nullptr, static_cast<cudaStream_t>(stream));
}

int hb_search_abort(hb_bank_t* bank) {
  HB_REQUIRE(bank != nullptr, "hb_search_abort: bank is NULL");
  Bank* b = reinterpret_cast<Bank*>(bank);
  b->pipe[0].begun = false;
  b->pipe[1].begun = false;
  return HB_OK;
}

int hb_eval_step(hb_bank_t* bank, const uint16_t* label_table_dev, int64_t table_rows, const float* q_dev,
                 int B, int S, int H, int W, const float* y_dev, int k, int k_prime, int64_t idx_offset,
                 float beta, int ignore_index, float* label_hat_dev, int64_t* conf_dev,
                 uint8_t* out_pred_dev, float* out_scores_dev, int64_t* out_idx_dev, void* stream) {
  HB_REQUIRE(bank != nullptr, "hb_eval_step: bank is NULL");
  HB_REQUIRE(B >= 0 && S >= 1 && H >= 1 && W >= 1, "hb_eval_step: bad shape B=%d S=%d H=%d W=%d", B, S, H, W);
  if (B == 0) return HB_OK;
  HB_REQUIRE(y_dev != nullptr && conf_dev != nullptr, "hb_eval_step: y_dev / conf_dev is NULL");
  const int64_t Q = static_cast<int64_t>(B) * S * S;
  int rc = hb_search_transfer(bank, label_table_dev, table_rows, q_dev, Q, k, k_prime, idx_offset, beta,
                              out_scores_dev, out_idx_dev, nullptr, label_hat_dev, stream);
  if (rc != HB_OK) return rc;
  Bank* b = reinterpret_cast<Bank*>(bank);
  rc = hb::predict_score_launch(label_hat_dev, B, S, b->C, H, W, y_dev, nullptr, ignore_index, conf_dev,
                                out_pred_dev, static_cast<cudaStream_t>(stream));
  if (rc == HB_ERR_UNSUPPORTED)
    hb::set_error("hb_eval_step: %d classes exceed the fused tail's shared-memory histogram; use "
                  "hb_search_transfer + hb_decode_mask + hb_predict_score", b->C);
  if (rc == HB_OK) b->last_launches++;
  return rc;
}

int hb_search_stats(hb_bank_t* bank, unsigned long long* stats_dev) {
  HB_REQUIRE(bank != nullptr, "hb_search_stats: bank is NULL");
  reinterpret_cast<Bank*>(bank)->cfg_stats = stats_dev;
  return HB_OK;
}

int hb_search_pacing(hb_bank_t* bank, int enable) {
  HB_REQUIRE(bank != nullptr, "hb_search_pacing: bank is NULL");
  reinterpret_cast<Bank*>(bank)->cfg_pace = enable != 0;
  return HB_OK;
}

int hb_search_tune(hb_bank_t* bank, int prefetch_tiles, int ablate) {
  HB_REQUIRE(bank != nullptr, "hb_search_tune: bank is NULL");
  HB_REQUIRE(prefetch_tiles >= -1 && prefetch_tiles <= 64, "hb_search_tune: prefetch_tiles=%d not in [-1, 64]", prefetch_tiles);
  HB_REQUIRE(ablate >= 0 && ablate <= 2, "hb_search_tune: ablate=%d not in {0,1,2}", ablate);
  reinterpret_cast<Bank*>(bank)->cfg_prefetch_tiles = prefetch_tiles;
  reinterpret_cast<Bank*>(bank)->cfg_ablate = ablate;
  return HB_OK;
}

int hb_coresidency_config(hb_bank_t* bank, int lean_search, int rerank_warps_per_cta, int rerank_shared_carveout) {
  HB_REQUIRE(bank != nullptr, "hb_coresidency_config: bank is NULL");
  HB_REQUIRE(rerank_warps_per_cta == 0 || rerank_warps_per_cta == 1 || rerank_warps_per_cta == 2 || rerank_warps_per_cta == 4,
             "hb_coresidency_config: rerank_warps_per_cta=%d not in {0, 1, 2, 4}", rerank_warps_per_cta);
  HB_REQUIRE(rerank_shared_carveout >= -1 && rerank_shared_carveout <= 100,
             "hb_coresidency_config: rerank_shared_carveout=%d not in [-1, 100]", rerank_shared_carveout);
  Bank* b = reinterpret_cast<Bank*>(bank);
  b->cfg_lean_search = lean_search != 0;
  b->cfg_rerank_warps = rerank_warps_per_cta ? rerank_warps_per_cta : 4;
  b->cfg_rerank_carveout = rerank_shared_carveout;
  return HB_OK;
}

int hb_search_config(hb_bank_t* bank, int cta_group, int max_chunks) {
  HB_REQUIRE(bank != nullptr, "hb_search_config: bank is NULL");
  HB_REQUIRE(cta_group >= 0 && cta_group <= 2, "hb_search_config: cta_group=%d not in {0,1,2}", cta_group);
  HB_REQUIRE(max_chunks >= 0, "hb_search_config: max_chunks < 0");
  Bank* b = reinterpret_cast<Bank*>(bank);
  b->cfg_cta_group = cta_group;
  b->cfg_max_chunks = max_chunks;
  return HB_OK;
}

int hb_search_timing(hb_bank_t* bank, int enable) {
  HB_REQUIRE(bank != nullptr, "hb_search_timing: bank is NULL");
  Bank* b = reinterpret_cast<Bank*>(bank);
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  if (enable && b->ev_begin[0] == nullptr) {
    for (int i = 0; i < 64; ++i) {
      HB_CHECK_CUDA(cudaEventCreate(&b->ev_begin[i]));
      HB_CHECK_CUDA(cudaEventCreate(&b->ev_end[i]));
      HB_CHECK_CUDA(cudaEventCreate(&b->ev_rerank0[i]));
      HB_CHECK_CUDA(cudaEventCreate(&b->ev_rerank[i]));
    }
  }
  b->timing = enable != 0;
  b->timing_count = 0;
  return HB_OK;
}

int hb_search_kernel_time(hb_bank_t* bank, float* mean_ms_out, int* count_out) {
  HB_REQUIRE(bank != nullptr && mean_ms_out != nullptr && count_out != nullptr, "hb_search_kernel_time: NULL argument");
  Bank* b = reinterpret_cast<Bank*>(bank);
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  const int n = b->timing_count < 64 ? b->timing_count : 64;
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    HB_CHECK_CUDA(cudaEventSynchronize(b->ev_end[i]));
    HB_CHECK_CUDA(cudaEventElapsedTime(&ms, b->ev_begin[i], b->ev_end[i]));
    sum += ms;
  }
  *mean_ms_out = n ? static_cast<float>(sum / n) : 0.f;
  *count_out = n;
  return HB_OK;
}

int hb_search_rerank_time(hb_bank_t* bank, float* mean_ms_out, int* count_out) {
  HB_REQUIRE(bank != nullptr && mean_ms_out != nullptr && count_out != nullptr, "hb_search_rerank_time: NULL argument");
  Bank* b = reinterpret_cast<Bank*>(bank);
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  const int n = b->timing_count < 64 ? b->timing_count : 64;
  double sum = 0.0;
  int used = 0;
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    // a search that stopped after K2 (validation dump) never recorded this slot's event: skip it
    if (cudaEventSynchronize(b->ev_rerank[i]) != cudaSuccess ||
        cudaEventElapsedTime(&ms, b->ev_rerank0[i], b->ev_rerank[i]) != cudaSuccess) {
      (void)cudaGetLastError();
      continue;
    }
    sum += ms;
    ++used;
  }
  *mean_ms_out = used ? static_cast<float>(sum / used) : 0.f;
  *count_out = used;
  return HB_OK;
}

int hb_plan_search(int64_t rows, int64_t Q, int cta_group, int num_sms, int max_chunks, int* out4) {
  HB_REQUIRE(out4 != nullptr, "hb_plan_search: out4 is NULL");
  HB_REQUIRE(rows >= 1 && Q >= 1, "hb_plan_search: rows and Q must be positive");
  HB_REQUIRE(cta_group == 1 || cta_group == 2, "hb_plan_search: cta_group=%d not in {1,2}", cta_group);
  HB_REQUIRE(num_sms >= 1 && max_chunks >= 0, "hb_plan_search: bad num_sms/max_chunks");
  const hb::SearchPlan p = hb::plan_search(rows, Q, cta_group, num_sms, max_chunks);
  out4[0] = p.n_tiles;
  out4[1] = p.n_qblocks;
  out4[2] = p.n_chunks;
  out4[3] = p.n_units;
  return HB_OK;
}

int hb_search_last_launches(const hb_bank_t* bank) {
  return bank ? reinterpret_cast<const Bank*>(bank)->last_launches : 0;
}

int hb_search_dump_scores(hb_bank_t* bank, const float* q_dev, int64_t Q, float* out_dev, int cta_group,
                          void* stream) {
  HB_REQUIRE(bank != nullptr, "hb_search_dump_scores: bank is NULL");
  Bank* b = reinterpret_cast<Bank*>(bank);
  if (!b->finalized) {
    hb::set_error("hb_search_dump_scores: bank not finalized");
    return HB_ERR_STATE;
  }
  HB_REQUIRE(Q >= 1 && q_dev && out_dev, "hb_search_dump_scores: bad arguments");
  HB_REQUIRE(cta_group >= 0 && cta_group <= 2, "hb_search_dump_scores: cta_group=%d", cta_group);
  HB_REQUIRE(Q * b->rows <= (int64_t(1) << 28), "hb_search_dump_scores: Q*rows too large for a debug dump");
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  return hb::search_impl(b, q_dev, Q, 1, 64, 0, nullptr, nullptr, nullptr, out_dev, cta_group,
                         static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

}  // extern "C"
