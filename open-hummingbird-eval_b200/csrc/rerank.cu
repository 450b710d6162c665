// K2b — exact fp32 re-rank, and K3 — cross-shard merge.
// K2b restores the reference metric: the k' candidates per query that survive the bf16 tensor-core
// pass are re-scored as fp32 dot products of the ORIGINAL fp32 query with the fp32 copy of the
// unit-norm bank row — the arithmetic GpuIndexFlatIP performs (search_faiss.py:39-41,89) — and the
// top-k of those, sorted by descending score (ties: smaller row first), is what hb_search returns.
// HBM-bound gather: k' * d * 4 bytes per query.
#include "common.cuh"

namespace hb {

// Sort 32*R keys held R per lane (element e = r*32 + lane) across one warp, bitonic network.
template <int R, bool DESC>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t (&a)[R], int lane) {
  constexpr int n = 32 * R;
#pragma unroll
  for (int k = 2; k <= n; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j < 32) {
        // partner lives in lane ^ j, same register slot
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int e = r * 32 + lane;
          const uint64_t other = __shfl_xor_sync(0xffffffffu, a[r], j);
          const bool lower = (e & j) == 0;
          const bool asc = ((e & k) == 0) != DESC;
          const uint64_t lo = a[r] < other ? a[r] : other;
          const uint64_t hi = a[r] < other ? other : a[r];
          a[r] = (lower == asc) ? lo : hi;
        }
      } else {
        // partner lives in this lane, register slot r ^ (j/32)
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if ((r & jr) == 0) {
            const int r2 = r | jr;
            const int e = r * 32 + lane;
            const bool asc = ((e & k) == 0) != DESC;
            const uint64_t lo = a[r] < a[r2] ? a[r] : a[r2];
            const uint64_t hi = a[r] < a[r2] ? a[r2] : a[r];
            a[r] = asc ? lo : hi;
            a[r2] = asc ? hi : lo;
          }
        }
      }
    }
  }
}

// Sort a BITONIC sequence of 32*R keys (same layout): the last pass of the network above, log2(32R) stages.
template <int R, bool DESC>
__device__ __forceinline__ void warp_bitonic_merge(uint64_t (&a)[R], int lane) {
  constexpr int n = 32 * R;
#pragma unroll
  for (int j = n >> 1; j > 0; j >>= 1) {
    if (j < 32) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint64_t other = __shfl_xor_sync(0xffffffffu, a[r], j);
        const bool lower = (lane & j) == 0;
        const uint64_t lo = a[r] < other ? a[r] : other;
        const uint64_t hi = a[r] < other ? other : a[r];
        a[r] = (lower != DESC) ? lo : hi;
      }
    } else {
      const int jr = j >> 5;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if ((r & jr) == 0) {
          const int r2 = r | jr;
          const uint64_t lo = a[r] < a[r2] ? a[r] : a[r2];
          const uint64_t hi = a[r] < a[r2] ? a[r2] : a[r];
          a[r] = DESC ? hi : lo;
          a[r2] = DESC ? lo : hi;
        }
      }
    }
  }
}

// One (query, chunk) block of the search kernel's candidate buffer is two runs of k'/2 keys, each sorted
// by descending score (the two register lists of an epilogue row, search.cu; empty entries are key 0 at the
// end).  Read with the second run backwards the block is a bitonic sequence, which one merge pass sorts —
// no full sort.  (Equal scores may sit in any row order inside a run: the merge then still orders by score,
// which is all the selection needs; the exact re-rank sorts its own keys from scratch.)
template <int R>
__device__ __forceinline__ void load_chunk_bitonic(const uint64_t* __restrict__ src, int lane, uint64_t (&a)[R]) {
  constexpr int KP = 32 * R, KL = KP / 2;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    a[r] = src[e < KL ? e : KP - 1 - (e - KL)];
  }
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// K4a fused into the kernels that produce a query's final neighbour list (K2b when the bank is not
// sharded, the merge kernels K3 / K3x when it is): the warp that holds the k (score, row) pairs in
// its lanes turns them into label_hat[q] = sum_j softmax_j(cos_j / beta) * soft_label[row_j]
// (hbird_eval.py:575-609,632-636) without the list ever being re-read from HBM.  cos_j =
// score_j / ||q|| because bank rows are unit norm; soft label = uint16 histogram / pixels per
// patch.  Same arithmetic, in the same order, as the stand-alone label_transfer_kernel.
// Element e = r*32 + lane of the list is (sc[r], id[r]); id < 0 = no neighbour.
template <int R>
__device__ __forceinline__ void
label_transfer_lanes(const LabelOut& lo, int64_t qi, int64_t out_row, int k, const float (&sc)[R],
                     const int64_t (&id)[R]) {
  const int lane = threadIdx.x & 31;
  // F.normalize clamps the norm at eps = 1e-12 (hbird_eval.py:594)
  const float qn = fmaxf(lo.qnorm[qi], 1e-12f);
  float logit[R];
  float mx = -INFINITY;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    const bool ok = e < k && id[r] >= 0 && id[r] < lo.table_rows;
    logit[r] = ok ? (sc[r] / qn) / lo.beta : -INFINITY;
    mx = fmaxf(mx, logit[r]);
  }
  mx = warp_max(mx);
  float ex[R];
  float sum = 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    ex[r] = (logit[r] == -INFINITY) ? 0.f : expf(logit[r] - mx);
    sum += ex[r];
  }
  sum = warp_sum(sum);
  // softmax weight and table row of element e = r*32 + lane stay in this lane's registers and are
  // broadcast with shuffles below: the kernels stay free of shared memory.  A missing neighbour keeps
  // weight 0 and reads row 0 (no branch in the gather loop).
  float w[R];
  int64_t row[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool ok = logit[r] != -INFINITY;
    w[r] = ok ? ex[r] / sum : 0.f;
    row[r] = ok ? id[r] : 0;
  }
  const float fpp = static_cast<float>(lo.pp);
  const int C = lo.C;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + lane;
    const bool live = c < C;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int n = k - r * 32 < 32 ? k - r * 32 : 32;  // neighbours held in register slot r (warp-uniform)
#pragma unroll 8
      for (int j = 0; j < n; ++j) {
        const float wj = __shfl_sync(0xffffffffu, w[r], j);
        const int64_t rj = __shfl_sync(0xffffffffu, row[r], j);
        // soft label = histogram / pixels-per-patch (one_hot(...).mean(3), hbird_eval.py:319-320)
        const float lab = live ? static_cast<float>(__ldg(lo.table + rj * C + c)) / fpp : 0.f;
        acc += wj * lab;
      }
    }
    if (live) lo.out[out_row * C + c] = acc;
  }
}

// j-th largest (j counted from 1) of the values held by lanes 0 .. G-1; 0 when G < j.  Warp-uniform result.
__device__ __forceinline__ uint32_t nth_largest(uint32_t v, int G, int j, int lane) {
  if (G < j) return 0u;
  int rank = 0;  // lanes whose (value, lane) pair is greater than this lane's
  for (int s = 0; s < G; ++s) {
    const uint32_t o = __shfl_sync(0xffffffffu, v, s);
    rank += (o > v || (o == v && s < lane)) ? 1 : 0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, lane < G && rank == j - 1);
  return __shfl_sync(0xffffffffu, v, __ffs(m) - 1);
}

// The bf16-pass candidates of query qi, all chunks merged: its best k' keys, sorted descending over
// (register r, lane) = element r*32 + lane.
template <int R>
__device__ __forceinline__ void
merge_chunk_lists(const uint64_t* __restrict__ cand, int n_chunks, int64_t q_pad, int64_t qi, int lane,
                  uint64_t (&top)[R]) {
  constexpr int KP = 32 * R;
  load_chunk_bitonic<R>(cand + qi * KP, lane, top);
  warp_bitonic_merge<R, true>(top, lane);
  for (int c = 1; c < n_chunks; ++c) {
    uint64_t nxt[R];
    load_chunk_bitonic<R>(cand + (static_cast<int64_t>(c) * q_pad + qi) * KP, lane, nxt);
    warp_bitonic_merge<R, false>(nxt, lane);
    // top descending, nxt ascending: the element-wise max holds the k' largest of the union, as a
    // bitonic sequence again
#pragma unroll
    for (int r = 0; r < R; ++r) top[r] = top[r] > nxt[r] ? top[r] : nxt[r];
    warp_bitonic_merge<R, true>(top, lane);
  }
}

// SUBSET = phase 2 of the threshold exchange: `cand` holds the query's sorted shortlist (written by
// shortlist_kernel), and only its entries at or above the cross-shard bound are re-ranked.
template <int R, bool L2, bool LABEL, bool SUBSET>
__device__ __forceinline__ void
rerank_query(const float* __restrict__ q, const float* __restrict__ bank_f32,
             const __nv_bfloat16* __restrict__ bank_bf16, const uint64_t* __restrict__ cand,
             int n_chunks, int64_t q_pad, int64_t Q, int d, int dpad, int k, int64_t idx_offset,
             float* __restrict__ out_scores, int64_t* __restrict__ out_idx, const Scatter& sc,
             const LabelOut& lo) {
  constexpr int KP = 32 * R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t qi = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;  // 1, 2 or 4 warps per CTA
  if (qi >= Q) return;

  // ---- union of the per-chunk candidate lists -> best k' by the bf16-pass score ----
  uint64_t top[R];
  int n_keep = KP;  // leading entries of the sorted list that are re-ranked (warp-uniform)
  if (SUBSET) {
#pragma unroll
    for (int r = 0; r < R; ++r) top[r] = __ldcg(cand + qi * KP + r * 32 + lane);
    // Every shard published the scores at ranks k', k'/2, k'/4, k'/8 of its own sorted list.  If j
    // shards each hold k'/j candidates scoring >= x, then k' bank rows score >= x: the j-th largest of
    // the rank-k'/j statistics is a lower bound of the global k'-th best bf16 score, for j = 1, 2, 4, 8,
    // and every shard computes the same bound.  0 = "that shard has fewer candidates" (prunes nothing).
    uint4 st = make_uint4(0u, 0u, 0u, 0u);
    if (lane < sc.world) st = __ldcg(sc.stats_in + static_cast<int64_t>(lane) * sc.q_cap + qi);
    uint32_t bound = nth_largest(st.x, sc.world, 1, lane);
    bound = max(bound, nth_largest(st.y, sc.world, 2, lane));
    bound = max(bound, nth_largest(st.z, sc.world, 4, lane));
    bound = max(bound, nth_largest(st.w, sc.world, 8, lane));
    n_keep = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
      n_keep += __popc(__ballot_sync(0xffffffffu, top[r] != 0ull && static_cast<uint32_t>(top[r] >> 32) >= bound));
  } else {
    merge_chunk_lists<R>(cand, n_chunks, q_pad, qi, lane, top);
  }

  // ---- exact fp32 scores of the k' candidates ----
  const float4* q4 = reinterpret_cast<const float4*>(q + qi * d);
  const int d4 = d >> 2;
  uint64_t exact[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    exact[r] = 0ull;
    for (int c0 = 0; c0 < 32; c0 += 4) {
      if (SUBSET && r * 32 + c0 >= n_keep) break;  // the list is sorted: nothing further passes the bound
      uint32_t rows[4];
      bool valid[4];
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint64_t key = __shfl_sync(0xffffffffu, top[r], c0 + u);
        valid[u] = key != 0ull && (!SUBSET || r * 32 + c0 + u < n_keep);
        rows[u] = valid[u] ? key_row(key) : 0u;
        acc[u] = 0.f;
      }
      if (bank_f32 != nullptr) {
        for (int i = lane; i < d4; i += 32) {
          const float4 qv = __ldg(q4 + i);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(bank_f32 + static_cast<int64_t>(rows[u]) * d) + i);
            if (L2) {
              const float a = qv.x - m.x, b2 = qv.y - m.y, c = qv.z - m.z, e = qv.w - m.w;
              acc[u] -= a * a + b2 * b2 + c * c + e * e;  // -||q-x||^2: larger is nearer
            } else {
              acc[u] += qv.x * m.x + qv.y * m.y + qv.z * m.z + qv.w * m.w;
            }
          }
        }
      } else {
        for (int i = lane; i < d4; i += 32) {
          const float4 qv = __ldg(q4 + i);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint2 pk = __ldg(reinterpret_cast<const uint2*>(bank_bf16 + static_cast<int64_t>(rows[u]) * dpad) + i);
            const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&pk.x);
            const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&pk.y);
            if (L2) {
              const float a = qv.x - __low2float(lo), b2 = qv.y - __high2float(lo);
              const float c = qv.z - __low2float(hi), e = qv.w - __high2float(hi);
              acc[u] -= a * a + b2 * b2 + c * c + e * e;
            } else {
              acc[u] += qv.x * __low2float(lo) + qv.y * __high2float(lo) + qv.z * __low2float(hi) + qv.w * __high2float(hi);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float s = warp_sum(acc[u]);
        if (lane == c0 + u && valid[u]) exact[r] = make_key(s, rows[u]);
      }
    }
  }
  warp_bitonic_sort<R, true>(exact, lane);

  // ---- emit the top-k: locally, or into the window of the rank that owns this query ----
  float* os = out_scores ? out_scores + qi * k : nullptr;
  int64_t* oi = out_idx ? out_idx + qi * k : nullptr;
  if (sc.world) {
    int p = 0;
    while (p + 1 < sc.world && qi >= sc.qsplit[p + 1]) ++p;
    const int64_t row = qi - sc.qsplit[p];
    os = sc.scores[p] + row * k;  // peer memory (NVLink) unless p == rank
    oi = sc.idx[p] + row * k;
  }
  float fs[R];
  int64_t fi[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    const bool ok = exact[r] != 0ull;
    // L2 banks report squared distances, ascending (GpuIndexFlatL2, search_faiss.py:45-46,89)
    const float v = ok ? key_score(exact[r]) : -INFINITY;
    fs[r] = L2 ? -v : v;
    fi[r] = ok ? static_cast<int64_t>(key_row(exact[r])) + idx_offset : -1;
    if (e < k && os != nullptr) {
      os[e] = fs[r];
      oi[e] = fi[r];
    }
  }
  if (LABEL && !L2) label_transfer_lanes<R>(lo, qi, qi, k, fs, fi);
}

// LABEL = with the fused label transfer.  No variant owns shared memory.  Issued on a second stream
// while the next batch's search runs (hb_search_begin / _finish), these CTAs fill the SMs the search
// kernel's CTAs vacate as they finish, and the launch gaps; they are NOT co-resident with a search
// CTA of the default build: its ten 168-register warps fill the register file of two of the four SM
// sub-partitions, and a CTA is placed only if every one of its warps finds room (measured:
// tools/pipe_timeline.py; hb_coresidency_config selects a 128-register search build beside which
// they do run, which does not pay under the power cap).
template <int R, bool L2, bool LABEL, bool SUBSET = false>
__global__ void __launch_bounds__(128)
rerank_kernel(const float* __restrict__ q, const float* __restrict__ bank_f32,
              const __nv_bfloat16* __restrict__ bank_bf16, const uint64_t* __restrict__ cand,
              int n_chunks, int64_t q_pad, int64_t Q, int d, int dpad, int k, int64_t idx_offset,
              float* __restrict__ out_scores, int64_t* __restrict__ out_idx,
              const __grid_constant__ Scatter sc, const LabelOut lo) {
  if (SUBSET && __ldcg(sc.timeout_flag) != 0u) return;  // a peer's statistics never arrived (whole grid returns)
  rerank_query<R, L2, LABEL, SUBSET>(q, bank_f32, bank_bf16, cand, n_chunks, q_pad, Q, d, dpad, k, idx_offset,
                             out_scores, out_idx, sc, lo);
  if (sc.world) {
    // Fused exchange: every thread's peer stores are ordered before the CTA counts itself done;
    // the last CTA of the grid then raises this rank's arrival flag on every peer.
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned prev = atomicAdd(sc.done_ctas, 1u);
      if (prev == gridDim.x - 1) {
        *sc.done_ctas = 0u;  // ready for the next launch on this stream
        __threadfence_system();
        for (int p = 0; p < sc.world; ++p) st_release_sys(sc.flag[p], sc.step);
      }
    }
  }
}

// Phase 1 of the threshold exchange: one warp per query merges the chunk lists into the sorted bf16
// top-k', stores it back over the query's chunk-0 list (the only reader of the other chunks is this
// warp) and writes the list's order statistics into every rank's window; the last CTA publishes the
// step on every peer's statistics flag (same protocol as the result scatter).
template <int R>
__global__ void __launch_bounds__(128)
shortlist_kernel(uint64_t* __restrict__ cand, int n_chunks, int64_t q_pad, int64_t Q,
                 const __grid_constant__ Scatter sc) {
  constexpr int KP = 32 * R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t qi = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  if (qi < Q) {
    uint64_t top[R];
    merge_chunk_lists<R>(cand, n_chunks, q_pad, qi, lane, top);
#pragma unroll
    for (int r = 0; r < R; ++r) cand[qi * KP + r * 32 + lane] = top[r];
    // ordered bf16 score at ranks k', k'/2, k'/4, k'/8 (element rank-1 of the list; 0 = no such candidate)
    constexpr int e0 = KP - 1, e1 = KP / 2 - 1, e2 = KP / 4 - 1, e3 = KP / 8 - 1;
    uint4 st;
    st.x = static_cast<uint32_t>(__shfl_sync(0xffffffffu, top[e0 / 32], e0 % 32) >> 32);
    st.y = static_cast<uint32_t>(__shfl_sync(0xffffffffu, top[e1 / 32], e1 % 32) >> 32);
    st.z = static_cast<uint32_t>(__shfl_sync(0xffffffffu, top[e2 / 32], e2 % 32) >> 32);
    st.w = static_cast<uint32_t>(__shfl_sync(0xffffffffu, top[e3 / 32], e3 % 32) >> 32);
    if (lane < sc.world) sc.stats[lane][qi] = st;  // lane p stores into rank p's window
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(sc.done_ctas2, 1u);
    if (prev == gridDim.x - 1) {
      *sc.done_ctas2 = 0u;
      __threadfence_system();
      for (int p = 0; p < sc.world; ++p) st_release_sys(sc.flag2[p], sc.step);
    }
  }
}

__global__ void exchange_wait_kernel(const uint32_t* flags, uint32_t step, int G, unsigned long long timeout_ns,
                                     unsigned int* timeout_flag);

int rerank_launch(const Bank* b, const float* q, int64_t Q, int k, int kp,
                  int n_chunks, int64_t q_pad, const uint64_t* cand, int64_t idx_offset,
                  float* out_scores, int64_t* out_idx, const Scatter* sc, const LabelOut* lo,
                  cudaStream_t st) {
  const int wpb = (b->cfg_rerank_warps == 1 || b->cfg_rerank_warps == 2) ? b->cfg_rerank_warps : 4;
  const unsigned blocks = static_cast<unsigned>(ceil_div64(Q, wpb));
  const unsigned threads = 32u * static_cast<unsigned>(wpb);
  const Scatter scatter = sc ? *sc : Scatter();
  const LabelOut label = lo ? *lo : LabelOut();
  if (scatter.world && scatter.phase == 1) {  // threshold exchange, phase 1: shortlist + statistics broadcast
    const unsigned sblocks = static_cast<unsigned>(ceil_div64(Q, 4));
    uint64_t* cw = const_cast<uint64_t*>(cand);
    if (kp == 32) shortlist_kernel<1><<<sblocks, 128, 0, st>>>(cw, n_chunks, q_pad, Q, scatter);
    else if (kp == 64) shortlist_kernel<2><<<sblocks, 128, 0, st>>>(cw, n_chunks, q_pad, Q, scatter);
    else if (kp == 128) shortlist_kernel<4><<<sblocks, 128, 0, st>>>(cw, n_chunks, q_pad, Q, scatter);
    else {
      set_error("rerank: k_prime=%d not in {32, 64, 128}", kp);
      return HB_ERR_INVALID;
    }
    HB_CHECK_CUDA(cudaGetLastError());
    return HB_OK;
  }
  if (scatter.world && scatter.phase == 2) {  // phase 2: wait for every peer's statistics, re-rank the survivors
    exchange_wait_kernel<<<1, 32, 0, st>>>(scatter.flags2_in, scatter.step, scatter.world, scatter.timeout_ns,
                                           scatter.timeout_flag);
    HB_CHECK_CUDA(cudaGetLastError());
    if (kp == 32) rerank_kernel<1, false, false, true><<<blocks, threads, 0, st>>>(q, b->feat_f32, b->feat_bf16, cand, 1, q_pad, Q, b->d, b->dpad, k, idx_offset, out_scores, out_idx, scatter, label);
    else if (kp == 64) rerank_kernel<2, false, false, true><<<blocks, threads, 0, st>>>(q, b->feat_f32, b->feat_bf16, cand, 1, q_pad, Q, b->d, b->dpad, k, idx_offset, out_scores, out_idx, scatter, label);
    else rerank_kernel<4, false, false, true><<<blocks, threads, 0, st>>>(q, b->feat_f32, b->feat_bf16, cand, 1, q_pad, Q, b->d, b->dpad, k, idx_offset, out_scores, out_idx, scatter, label);
    HB_CHECK_CUDA(cudaGetLastError());
    return HB_OK;
  }
#define HB_RERANK_ARGS q, b->feat_f32, b->feat_bf16, cand, n_chunks, q_pad, Q, b->d, b->dpad, k, idx_offset, out_scores, out_idx, scatter, label
#define HB_RERANK_ONE(R, L2, LABEL)                                                               \
  do {                                                                                            \
    if (b->cfg_rerank_carveout >= 0)                                                              \
      HB_CHECK_CUDA(cudaFuncSetAttribute(rerank_kernel<R, L2, LABEL>,                             \
                                         cudaFuncAttributePreferredSharedMemoryCarveout,          \
                                         b->cfg_rerank_carveout));                                \
    rerank_kernel<R, L2, LABEL><<<blocks, threads, 0, st>>>(HB_RERANK_ARGS);                      \
  } while (0)
#define HB_RERANK(R)                                                                              \
  do {                                                                                            \
    if (b->flags & HB_BANK_L2) HB_RERANK_ONE(R, true, false);                                     \
    else if (label.table != nullptr) HB_RERANK_ONE(R, false, true);                               \
    else HB_RERANK_ONE(R, false, false);                                                          \
  } while (0)
  if (kp == 32) HB_RERANK(1);
  else if (kp == 64) HB_RERANK(2);
  else if (kp == 128) HB_RERANK(4);
  else {
    set_error("rerank: k_prime=%d not in {32, 64, 128}", kp);
    return HB_ERR_INVALID;
  }
#undef HB_RERANK
#undef HB_RERANK_ONE
#undef HB_RERANK_ARGS
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

// K3: merge G per-shard sorted top-k lists per query.  One warp per query; k <= 128.
// List g of query qi starts at g * slot_stride + qi * k.
template <int R>
__device__ __forceinline__ void
merge_query(const float* ss, const int64_t* si, int G, int64_t slot_stride, int64_t qi, int k,
            float* __restrict__ out_scores, int64_t* __restrict__ out_idx, const LabelOut& lo) {
  // keys here carry a 64-bit index, so sort (ordered score, then index) pairs held as two words
  const int lane = threadIdx.x & 31;
  // candidate slots: position in the gathered (G, k) list, encoded as g*k + j in the low word
  uint64_t top[R];
#pragma unroll
  for (int r = 0; r < R; ++r) top[r] = 0ull;
  const int total = G * k;
  for (int base = 0; base < total; base += 32 * R) {
    uint64_t nxt[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int e = base + r * 32 + lane;
      nxt[r] = 0ull;
      if (e < total) {
        const int g = e / k, j = e % k;
        const int64_t off = static_cast<int64_t>(g) * slot_stride + qi * k + j;
        const int64_t id = __ldcg(si + off);
        // position e doubles as the tie-break: shards hold ascending global rows, and within a
        // shard ties are already ordered by row, so smaller e <=> smaller global index
        if (id >= 0) nxt[r] = make_key(__ldcg(ss + off), static_cast<uint32_t>(e));
      }
    }
    warp_bitonic_sort<R, false>(nxt, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) top[r] = top[r] > nxt[r] ? top[r] : nxt[r];
    warp_bitonic_sort<R, true>(top, lane);
  }
  float fs[R];
  int64_t fi[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    fs[r] = -INFINITY;
    fi[r] = -1;
    if (e < k) {
      if (top[r] != 0ull) {
        const uint32_t pos = key_row(top[r]);
        const int g = pos / k, j = pos % k;
        const int64_t off = static_cast<int64_t>(g) * slot_stride + qi * k + j;
        fi[r] = __ldcg(si + off);
        fs[r] = __ldcg(ss + off);
      }
      if (out_scores != nullptr) {
        out_scores[qi * k + e] = fs[r];
        out_idx[qi * k + e] = fi[r];
      }
    }
  }
  if (lo.table != nullptr) label_transfer_lanes<R>(lo, qi, qi, k, fs, fi);
}

template <int R>
__global__ void __launch_bounds__(128)
merge_topk_kernel(const float* __restrict__ ss, const int64_t* __restrict__ si, int G, int64_t Q,
                  int k, float* __restrict__ out_scores, int64_t* __restrict__ out_idx,
                  const LabelOut lo) {
  const int64_t qi = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (qi >= Q) return;
  merge_query<R>(ss, si, G, Q * k, qi, k, out_scores, out_idx, lo);
}

// K3x: the receiving half of the fused exchange, in two launches.
// exchange_wait_kernel — ONE warp polls the G step flags of this rank's window until every source rank
// has published `step` (the rows were stored by the peers' K2b kernels over NVLink; acquire at system
// scope).  A single polling warp instead of a wait at the head of every merge CTA: the merge is usually
// issued while the next batch's search runs (pipeline.py), and hundreds of CTAs polling peer-visible
// memory slowed that search by several per cent.  The wait is bounded: after `timeout_ns` the kernel
// records which rank is missing in `timeout_flag` — no trap, the context stays usable;
// hb_exchange_status() turns the flag into an error on the host.
// merge_window_kernel — stream-ordered after it: merges the G lists of each query of the local slice
// (+ fused label transfer); writes nothing if the wait gave up.
__global__ void
exchange_wait_kernel(const uint32_t* flags, uint32_t step, int G, unsigned long long timeout_ns,
                     unsigned int* timeout_flag) {
  if (threadIdx.x < G) {
    unsigned long long t0 = 0;
    unsigned backoff = 64;
    // flags count exchanges; a peer may already be one step ahead (its data for that step went
    // to the other buffer), hence >= on the wrapped difference
    while (static_cast<int32_t>(ld_acquire_sys(flags + threadIdx.x) - step) < 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > timeout_ns) {
        atomicCAS(timeout_flag, 0u, 0x80000000u | threadIdx.x);  // first missing rank wins
        break;
      }
      __nanosleep(backoff);
      if (backoff < 2048) backoff *= 2;
    }
  }
  __threadfence_system();
}

template <int R>
__global__ void __launch_bounds__(128)
merge_window_kernel(const float* ss, const int64_t* si, int G, int64_t slot_stride, int64_t rows, int k,
                    const unsigned int* timeout_flag, float* __restrict__ out_scores,
                    int64_t* __restrict__ out_idx, const LabelOut lo) {
  if (__ldcg(timeout_flag) != 0u) return;  // a peer never arrived: leave the outputs alone
  const int64_t qi = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (qi >= rows) return;
  merge_query<R>(ss, si, G, slot_stride, qi, k, out_scores, out_idx, lo);
}

int merge_window_launch(const Exchange* x, uint32_t step, int64_t rows, int k, float* out_scores,
                        int64_t* out_idx, const LabelOut* lo, cudaStream_t st) {
  const LabelOut label = lo ? *lo : LabelOut();
  const uint8_t* win = x->window[x->rank];
  const int parity = static_cast<int>(step & 1u);
  const float* ss = reinterpret_cast<const float*>(win + x->scores_off(parity));
  const int64_t* si = reinterpret_cast<const int64_t*>(win + x->idx_off(parity));
  const uint32_t* flags = reinterpret_cast<const uint32_t*>(win);
  const unsigned long long timeout_ns = x->timeout_ms * 1000000ull;  // hang protection only (default 10 min)
  const int64_t stride = static_cast<int64_t>(x->slot_elems());
  // the wait runs even for an empty slice: it keeps the ranks within one step of each other, which is
  // what makes two window buffers enough
  exchange_wait_kernel<<<1, 32, 0, st>>>(flags, step, x->world, timeout_ns, x->timeout_flag);
  HB_CHECK_CUDA(cudaGetLastError());
  if (rows <= 0) return HB_OK;
  const unsigned blocks = static_cast<unsigned>(ceil_div64(rows, 4));
#define HB_MERGE_WIN(R)                                                                         \
  merge_window_kernel<R><<<blocks, 128, 0, st>>>(ss, si, x->world, stride, rows, k, x->timeout_flag, out_scores, out_idx, label)
  if (k <= 32) HB_MERGE_WIN(1);
  else if (k <= 64) HB_MERGE_WIN(2);
  else HB_MERGE_WIN(4);
#undef HB_MERGE_WIN
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

}  // namespace hb

static int merge_topk_impl(const float* shard_scores_dev, const int64_t* shard_idx_dev, int G, int64_t Q, int k,
                           float* out_scores_dev, int64_t* out_idx_dev, const hb::LabelOut& lo, void* stream) {
  const unsigned blocks = static_cast<unsigned>(hb::ceil_div64(Q, 4));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k <= 32) hb::merge_topk_kernel<1><<<blocks, 128, 0, st>>>(shard_scores_dev, shard_idx_dev, G, Q, k, out_scores_dev, out_idx_dev, lo);
  else if (k <= 64) hb::merge_topk_kernel<2><<<blocks, 128, 0, st>>>(shard_scores_dev, shard_idx_dev, G, Q, k, out_scores_dev, out_idx_dev, lo);
  else hb::merge_topk_kernel<4><<<blocks, 128, 0, st>>>(shard_scores_dev, shard_idx_dev, G, Q, k, out_scores_dev, out_idx_dev, lo);
  HB_CHECK_CUDA(cudaGetLastError());
  return HB_OK;
}

extern "C" int hb_merge_topk(const float* shard_scores_dev, const int64_t* shard_idx_dev, int G,
                             int64_t Q, int k, float* out_scores_dev, int64_t* out_idx_dev,
                             void* stream) {
  HB_REQUIRE(G >= 1 && G <= 1024, "hb_merge_topk: G=%d not in [1, 1024]", G);
  HB_REQUIRE(k >= 1 && k <= 128, "hb_merge_topk: k=%d not in [1, 128]", k);
  HB_REQUIRE(Q >= 0, "hb_merge_topk: Q < 0");
  if (Q == 0) return HB_OK;
  HB_REQUIRE(shard_scores_dev && shard_idx_dev && out_scores_dev && out_idx_dev, "hb_merge_topk: NULL pointer");
  return merge_topk_impl(shard_scores_dev, shard_idx_dev, G, Q, k, out_scores_dev, out_idx_dev, hb::LabelOut(), stream);
}

extern "C" int hb_merge_topk_transfer(const float* shard_scores_dev, const int64_t* shard_idx_dev, int G,
                                      int64_t Q, int k, const uint16_t* label_table_dev, int64_t table_rows,
                                      int C, int patch_pixels, const float* qnorm_dev, float beta,
                                      float* out_scores_dev, int64_t* out_idx_dev,
                                      float* out_label_hat_dev, void* stream) {
  HB_REQUIRE(G >= 1 && G <= 1024, "hb_merge_topk_transfer: G=%d not in [1, 1024]", G);
  HB_REQUIRE(k >= 1 && k <= 128, "hb_merge_topk_transfer: k=%d not in [1, 128]", k);
  HB_REQUIRE(C >= 1 && C <= 256 && patch_pixels >= 1 && beta > 0.f, "hb_merge_topk_transfer: bad C/patch_pixels/beta");
  HB_REQUIRE(Q >= 0, "hb_merge_topk_transfer: Q < 0");
  if (Q == 0) return HB_OK;
  HB_REQUIRE(shard_scores_dev && shard_idx_dev && label_table_dev && qnorm_dev && out_label_hat_dev,
             "hb_merge_topk_transfer: NULL pointer");
  HB_REQUIRE((out_scores_dev == nullptr) == (out_idx_dev == nullptr), "hb_merge_topk_transfer: give both or neither of out_scores/out_idx");
  hb::LabelOut lo;
  lo.table = label_table_dev;
  lo.table_rows = table_rows;
  lo.C = C;
  lo.pp = patch_pixels;
  lo.beta = beta;
  lo.qnorm = qnorm_dev;
  lo.out = out_label_hat_dev;
  return merge_topk_impl(shard_scores_dev, shard_idx_dev, G, Q, k, out_scores_dev, out_idx_dev, lo, stream);
}
