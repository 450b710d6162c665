// K3x — fused shard exchange over NVLink peer memory.
// The reference merges its per-GPU shard results on the host (faiss.IndexShards,
// search_faiss.py:53-63).  Here each rank owns a row shard of the bank, every rank searches every
// query, and only the rank that post-processes a query needs its merged neighbours.  So the
// exchange is an all-to-all of (score f32, idx i64)[rows, k] slices, and it is fused into the
// kernels on either side of it instead of being a collective call:
//   * K2b (rerank.cu) stores each query's top-k straight into the owner rank's window through a
//     CUDA-IPC mapping (NVLink stores), and its last CTA publishes a step counter on every peer;
//   * K3x (merge_window_kernel) waits on those counters and merges the G lists of its slice.
// No NCCL call, no staging copy, no host synchronisation; two window buffers alternate by step
// (a rank can be at most one exchange ahead of a peer, because its own merge waits for that
// peer's previous scatter).
#include <string.h>

#include "common.cuh"

using hb::Bank;
using hb::Exchange;

extern "C" {

int hb_exchange_create(int device, int rank, int world, int64_t slice_capacity, int max_k,
                       hb_exchange_t** out) {
  HB_REQUIRE(out != nullptr, "hb_exchange_create: out is NULL");
  *out = nullptr;
  HB_REQUIRE(world >= 1 && world <= hb::kMaxPeers, "hb_exchange_create: world=%d not in [1, %d]", world, hb::kMaxPeers);
  HB_REQUIRE(rank >= 0 && rank < world, "hb_exchange_create: rank=%d not in [0, %d)", rank, world);
  HB_REQUIRE(slice_capacity >= 1 && slice_capacity < (int64_t(1) << 31), "hb_exchange_create: slice_capacity=%lld out of range", (long long)slice_capacity);
  HB_REQUIRE(max_k >= 1 && max_k <= 128, "hb_exchange_create: max_k=%d not in [1, 128]", max_k);
  int rc = hb_device_check(device, nullptr);
  if (rc != HB_OK) return rc;
  HB_CHECK_CUDA(cudaSetDevice(device));
  Exchange* x = new Exchange();
  x->device = device;
  x->rank = rank;
  x->world = world;
  x->cap = slice_capacity;
  x->kmax = max_k;
  x->bytes = x->window_bytes();
  cudaError_t e = cudaMalloc(&x->window[rank], x->bytes);
  if (e == cudaSuccess) e = cudaMemset(x->window[rank], 0, 256);
  // [0] CTAs done with the result scatter, [1] timeout flag, [2] CTAs done with the statistics broadcast
  if (e == cudaSuccess) e = cudaMalloc(&x->done_ctas, 4 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(x->done_ctas, 0, 4 * sizeof(unsigned int));
  if (e != cudaSuccess) {
    hb::set_error("hb_exchange_create: allocating a %zu-byte window failed: %s", x->bytes, cudaGetErrorString(e));
    (void)cudaGetLastError();
    hb_exchange_destroy(reinterpret_cast<hb_exchange_t*>(x));
    return e == cudaErrorMemoryAllocation ? HB_ERR_OOM : HB_ERR_CUDA;
  }
  x->timeout_flag = x->done_ctas + 1;
  HB_CHECK_CUDA(cudaDeviceSynchronize());
  x->connected = world == 1;
  *out = reinterpret_cast<hb_exchange_t*>(x);
  return HB_OK;
}

int hb_exchange_disconnect(hb_exchange_t* xchg) {
  if (!xchg) return HB_OK;
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < x->world; ++p) {
    if (p == x->rank) continue;
    if (x->ipc_opened[p] && x->window[p]) cudaIpcCloseMemHandle(x->window[p]);
    x->ipc_opened[p] = false;
    x->window[p] = nullptr;
  }
  (void)cudaGetLastError();
  x->connected = x->world == 1;
  return HB_OK;
}

int hb_exchange_destroy(hb_exchange_t* xchg) {
  if (!xchg) return HB_OK;
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  hb_exchange_disconnect(xchg);
  if (x->window[x->rank]) cudaFree(x->window[x->rank]);
  if (x->done_ctas) cudaFree(x->done_ctas);
  (void)cudaGetLastError();
  delete x;
  return HB_OK;
}

int hb_exchange_handle(hb_exchange_t* xchg, void* handle_out, int handle_bytes) {
  HB_REQUIRE(xchg && handle_out, "hb_exchange_handle: NULL argument");
  HB_REQUIRE(handle_bytes == HB_EXCHANGE_HANDLE_BYTES, "hb_exchange_handle: handle_bytes=%d, expected %d", handle_bytes, HB_EXCHANGE_HANDLE_BYTES);
  static_assert(sizeof(cudaIpcMemHandle_t) == HB_EXCHANGE_HANDLE_BYTES, "IPC handle size");
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  HB_CHECK_CUDA(cudaSetDevice(x->device));
  cudaIpcMemHandle_t h;
  HB_CHECK_CUDA(cudaIpcGetMemHandle(&h, x->window[x->rank]));
  memcpy(handle_out, &h, sizeof(h));
  return HB_OK;
}

int hb_exchange_connect(hb_exchange_t* xchg, const void* handles, int n_handles) {
  HB_REQUIRE(xchg && handles, "hb_exchange_connect: NULL argument");
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  HB_REQUIRE(n_handles == x->world, "hb_exchange_connect: %d handles for world=%d", n_handles, x->world);
  HB_CHECK_CUDA(cudaSetDevice(x->device));
  const uint8_t* hs = static_cast<const uint8_t*>(handles);
  for (int p = 0; p < x->world; ++p) {
    if (p == x->rank || x->window[p]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + static_cast<size_t>(p) * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      hb::set_error("hb_exchange_connect: cannot map the window of rank %d (no peer access between the "
                    "GPUs, or CUDA IPC unavailable): %s", p, cudaGetErrorString(e));
      (void)cudaGetLastError();
      return HB_ERR_UNSUPPORTED;
    }
    x->window[p] = static_cast<uint8_t*>(ptr);
    x->ipc_opened[p] = true;
  }
  x->connected = true;
  return HB_OK;
}

int hb_exchange_connect_local(hb_exchange_t* xchg, hb_exchange_t* const* peers, int n_peers) {
  HB_REQUIRE(xchg && peers, "hb_exchange_connect_local: NULL argument");
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  HB_REQUIRE(n_peers == x->world, "hb_exchange_connect_local: %d peers for world=%d", n_peers, x->world);
  for (int p = 0; p < x->world; ++p) {
    const Exchange* y = reinterpret_cast<const Exchange*>(peers[p]);
    HB_REQUIRE(y && y->rank == p && y->world == x->world && y->cap == x->cap && y->kmax == x->kmax,
               "hb_exchange_connect_local: peer %d does not match (rank/world/capacity/max_k)", p);
    if (p != x->rank) x->window[p] = y->window[p];
  }
  x->connected = true;
  return HB_OK;
}

// Argument checks shared by the one-call and the split form of the scatter, and the Scatter record
// (where each query's rows go) for exchange number `step`.
static int prepare_scatter(Bank* b, Exchange* x, int64_t Q, int k, int k_prime, const int64_t* qsplit_host,
                           uint32_t step, hb::Scatter* sc, const char* who) {
  if (!b->finalized) {
    hb::set_error("%s: bank not finalized (call hb_bank_finalize first)", who);
    return HB_ERR_STATE;
  }
  if (!x->connected) {
    hb::set_error("%s: exchange not connected (call hb_exchange_connect first)", who);
    return HB_ERR_STATE;
  }
  HB_REQUIRE(b->device == x->device, "%s: bank on device %d, exchange on device %d", who, b->device, x->device);
  HB_REQUIRE(Q >= 1 && Q < (int64_t(1) << 31), "%s: Q=%lld out of range", who, (long long)Q);
  HB_REQUIRE(k >= 1 && k <= k_prime && k <= x->kmax, "%s: need 1 <= k (%d) <= min(k_prime %d, max_k %d)", who, k, k_prime, x->kmax);
  HB_REQUIRE(k_prime == 32 || k_prime == 64 || k_prime == 128, "%s: k_prime=%d not in {32, 64, 128}", who, k_prime);
  HB_REQUIRE(qsplit_host != nullptr, "%s: NULL pointer", who);
  HB_REQUIRE(b->rows >= 1, "%s: the bank shard is empty", who);
  // the cross-shard merge keeps the LARGEST scores; squared-L2 results are ascending distances
  HB_REQUIRE((b->flags & HB_BANK_L2) == 0, "%s: squared-L2 banks cannot be row-sharded (inner-product metric only)", who);
  HB_REQUIRE(qsplit_host[0] == 0 && qsplit_host[x->world] == Q, "%s: qsplit must run from 0 to Q", who);
  for (int p = 0; p < x->world; ++p) {
    const int64_t n = qsplit_host[p + 1] - qsplit_host[p];
    HB_REQUIRE(n >= 0 && n <= x->cap, "%s: slice %d has %lld rows, window capacity is %lld", who, p, (long long)n, (long long)x->cap);
  }
  const int parity = static_cast<int>(step & 1u);
  sc->world = x->world;
  sc->rank = x->rank;
  sc->step = step;
  sc->done_ctas = x->done_ctas;
  for (int p = 0; p <= x->world; ++p) sc->qsplit[p] = qsplit_host[p];
  for (int p = 0; p < x->world; ++p) {
    uint8_t* win = x->window[p];
    sc->scores[p] = reinterpret_cast<float*>(win + x->scores_off(parity)) + x->slot_elems() * x->rank;
    sc->idx[p] = reinterpret_cast<int64_t*>(win + x->idx_off(parity)) + x->slot_elems() * x->rank;
    sc->flag[p] = reinterpret_cast<uint32_t*>(win) + x->rank;
    sc->stats[p] = reinterpret_cast<uint4*>(win + x->stats_off(parity)) + x->q_cap() * x->rank;
    sc->flag2[p] = reinterpret_cast<uint32_t*>(win) + hb::kMaxPeers + x->rank;
  }
  sc->phase = (x->mode == 1 && x->world > 1) ? 1 : 0;  // a single shard has nobody to swap statistics with
  sc->stats_in = reinterpret_cast<const uint4*>(x->window[x->rank] + x->stats_off(parity));
  sc->flags2_in = reinterpret_cast<const uint32_t*>(x->window[x->rank]) + hb::kMaxPeers;
  sc->q_cap = x->q_cap();
  sc->done_ctas2 = x->done_ctas + 2;
  sc->timeout_flag = x->timeout_flag;
  sc->timeout_ns = x->timeout_ms * 1000000ull;
  return HB_OK;
}

// Threshold exchange: remember what phase 2 needs (the slot's candidate buffer now holds the sorted
// shortlists) until hb_exchange_rerank / hb_exchange_merge* issues it.
static void remember_phase2(Exchange* x, Bank* b, int slot, const float* q, int k, int64_t idx_offset, const hb::Scatter& sc) {
  const hb::PipeSlot& ps = b->pipe[slot];
  Exchange::Pending& pd = x->pending;
  pd.active = true;
  pd.bank = b;
  pd.slot = slot;
  pd.q = q;
  pd.Q = ps.Q;
  pd.q_pad = ps.q_pad;
  pd.idx_offset = idx_offset;
  pd.k = k;
  pd.kp = ps.kp;
  pd.cand = ps.cand;
  pd.sc = sc;
  pd.sc.phase = 2;
}

static int issue_phase2(Exchange* x, cudaStream_t st) {
  Exchange::Pending& pd = x->pending;
  if (!pd.active) return HB_OK;
  pd.active = false;
  Bank* b = pd.bank;
  int rc = hb::rerank_launch(b, pd.q, pd.Q, pd.k, pd.kp, 1, pd.q_pad, pd.cand, pd.idx_offset, nullptr, nullptr, &pd.sc,
                             nullptr, st);
  if (rc != HB_OK) return rc;
  // the slot's candidate buffer is free for the next search only now
  hb::PipeSlot& ps = b->pipe[pd.slot];
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) (void)cudaGetLastError();
  if (cap == cudaStreamCaptureStatusNone) {
    if (ps.done == nullptr) HB_CHECK_CUDA(cudaEventCreateWithFlags(&ps.done, cudaEventDisableTiming));
    HB_CHECK_CUDA(cudaEventRecord(ps.done, st));
  }
  b->last_launches += 2;
  return HB_OK;
}

int hb_search_scatter(hb_bank_t* bank, hb_exchange_t* xchg, const float* q_dev, int64_t Q, int k,
                      int k_prime, int64_t idx_offset, const int64_t* qsplit_host,
                      float* out_qnorm_dev, void* stream) {
  HB_REQUIRE(bank && xchg, "hb_search_scatter: NULL handle");
  Bank* b = reinterpret_cast<Bank*>(bank);
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  HB_REQUIRE(q_dev != nullptr, "hb_search_scatter: NULL pointer");
  // the step counter advances only once the launches are in the stream: a call that fails before
  // that leaves this end of the exchange as it was
  const uint32_t step = x->step + 1;
  hb::Scatter sc;
  int rc = prepare_scatter(b, x, Q, k, k_prime, qsplit_host, step, &sc, "hb_search_scatter");
  if (rc != HB_OK) return rc;
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  x->pending.active = false;  // a phase 2 that was never issued belongs to an abandoned batch (error recovery)
  rc = hb::search_impl(b, q_dev, Q, k, k_prime, idx_offset, nullptr, nullptr, out_qnorm_dev, nullptr, 0,
                       static_cast<cudaStream_t>(stream), &sc, nullptr);
  if (rc != HB_OK) return rc;
  if (sc.phase == 1) remember_phase2(x, b, 0, q_dev, k, idx_offset, sc);
  x->step = step;
  x->last_rows = qsplit_host[x->rank + 1] - qsplit_host[x->rank];
  x->last_k = k;
  return HB_OK;
}

int hb_search_finish_scatter(hb_bank_t* bank, hb_exchange_t* xchg, int slot, const float* q_dev, int k,
                             int64_t idx_offset, const int64_t* qsplit_host, void* stream) {
  HB_REQUIRE(bank && xchg, "hb_search_finish_scatter: NULL handle");
  Bank* b = reinterpret_cast<Bank*>(bank);
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  HB_REQUIRE(slot == 0 || slot == 1, "hb_search_finish_scatter: slot=%d not in {0, 1}", slot);
  HB_REQUIRE(q_dev != nullptr, "hb_search_finish_scatter: NULL pointer");
  if (!b->pipe[slot].begun) {
    hb::set_error("hb_search_finish_scatter: pipeline slot %d holds no begun search (call hb_search_begin first)", slot);
    return HB_ERR_STATE;
  }
  const uint32_t step = x->step + 1;
  hb::Scatter sc;
  int rc = prepare_scatter(b, x, b->pipe[slot].Q, k, b->pipe[slot].kp, qsplit_host, step, &sc, "hb_search_finish_scatter");
  if (rc != HB_OK) return rc;
  HB_CHECK_CUDA(cudaSetDevice(b->device));
  x->pending.active = false;  // a phase 2 that was never issued belongs to an abandoned batch (error recovery)
  rc = hb::search_finish_impl(b, slot, q_dev, k, idx_offset, nullptr, nullptr, &sc, nullptr, static_cast<cudaStream_t>(stream));
  if (rc != HB_OK) return rc;
  if (sc.phase == 1) remember_phase2(x, b, slot, q_dev, k, idx_offset, sc);
  x->step = step;
  x->last_rows = qsplit_host[x->rank + 1] - qsplit_host[x->rank];
  x->last_k = k;
  return HB_OK;
}

int hb_exchange_merge(hb_exchange_t* xchg, float* out_scores_dev, int64_t* out_idx_dev, void* stream) {
  HB_REQUIRE(xchg != nullptr, "hb_exchange_merge: exchange is NULL");
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  if (x->step == 0 || x->last_k == 0) {
    hb::set_error("hb_exchange_merge: no scatter to merge (call hb_search_scatter first)");
    return HB_ERR_STATE;
  }
  HB_REQUIRE(x->last_rows == 0 || (out_scores_dev && out_idx_dev), "hb_exchange_merge: NULL output");
  HB_CHECK_CUDA(cudaSetDevice(x->device));
  int rc = issue_phase2(x, static_cast<cudaStream_t>(stream));
  if (rc != HB_OK) return rc;
  const int k = x->last_k;
  x->last_k = 0;  // one merge per scatter
  return hb::merge_window_launch(x, x->step, x->last_rows, k, out_scores_dev, out_idx_dev, nullptr,
                                 static_cast<cudaStream_t>(stream));
}

int hb_exchange_config(hb_exchange_t* xchg, int mode) {
  HB_REQUIRE(xchg != nullptr, "hb_exchange_config: exchange is NULL");
  HB_REQUIRE(mode == 0 || mode == 1, "hb_exchange_config: mode=%d not in {0, 1}", mode);
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  if (x->pending.active) {
    hb::set_error("hb_exchange_config: an exchange is in flight");
    return HB_ERR_STATE;
  }
  x->mode = mode;
  return HB_OK;
}

int hb_exchange_rerank(hb_exchange_t* xchg, void* stream) {
  HB_REQUIRE(xchg != nullptr, "hb_exchange_rerank: exchange is NULL");
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  HB_CHECK_CUDA(cudaSetDevice(x->device));
  return issue_phase2(x, static_cast<cudaStream_t>(stream));
}

int hb_exchange_merge_transfer(hb_exchange_t* xchg, const uint16_t* label_table_dev, int64_t table_rows, int C,
                               int patch_pixels, const float* qnorm_slice_dev, float beta,
                               float* out_scores_dev, int64_t* out_idx_dev, float* out_label_hat_dev,
                               void* stream) {
  HB_REQUIRE(xchg != nullptr, "hb_exchange_merge_transfer: exchange is NULL");
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  if (x->step == 0 || x->last_k == 0) {
    hb::set_error("hb_exchange_merge_transfer: no scatter to merge (call hb_search_scatter first)");
    return HB_ERR_STATE;
  }
  HB_REQUIRE(C >= 1 && C <= 256 && patch_pixels >= 1 && beta > 0.f, "hb_exchange_merge_transfer: bad C/patch_pixels/beta");
  HB_REQUIRE((out_scores_dev == nullptr) == (out_idx_dev == nullptr), "hb_exchange_merge_transfer: give both or neither of out_scores/out_idx");
  HB_REQUIRE(x->last_rows == 0 || (label_table_dev && qnorm_slice_dev && out_label_hat_dev), "hb_exchange_merge_transfer: NULL pointer");
  HB_CHECK_CUDA(cudaSetDevice(x->device));
  hb::LabelOut lo;
  lo.table = label_table_dev;
  lo.table_rows = table_rows;
  lo.C = C;
  lo.pp = patch_pixels;
  lo.beta = beta;
  lo.qnorm = qnorm_slice_dev;
  lo.out = out_label_hat_dev;
  int rc = issue_phase2(x, static_cast<cudaStream_t>(stream));
  if (rc != HB_OK) return rc;
  const int k = x->last_k;
  x->last_k = 0;  // one merge per scatter
  return hb::merge_window_launch(x, x->step, x->last_rows, k, out_scores_dev, out_idx_dev,
                                 x->last_rows > 0 ? &lo : nullptr, static_cast<cudaStream_t>(stream));
}

int hb_exchange_set_timeout(hb_exchange_t* xchg, int64_t timeout_ms) {
  HB_REQUIRE(xchg != nullptr, "hb_exchange_set_timeout: exchange is NULL");
  HB_REQUIRE(timeout_ms >= 1, "hb_exchange_set_timeout: timeout_ms=%lld must be positive", (long long)timeout_ms);
  reinterpret_cast<Exchange*>(xchg)->timeout_ms = static_cast<unsigned long long>(timeout_ms);
  return HB_OK;
}

int hb_exchange_status(hb_exchange_t* xchg, void* stream) {
  HB_REQUIRE(xchg != nullptr, "hb_exchange_status: exchange is NULL");
  Exchange* x = reinterpret_cast<Exchange*>(xchg);
  HB_CHECK_CUDA(cudaSetDevice(x->device));
  unsigned int flag = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HB_CHECK_CUDA(cudaMemcpyAsync(&flag, x->timeout_flag, sizeof(flag), cudaMemcpyDeviceToHost, st));
  HB_CHECK_CUDA(cudaStreamSynchronize(st));
  if (flag != 0u) {
    HB_CHECK_CUDA(cudaMemsetAsync(x->timeout_flag, 0, sizeof(flag), st));  // reported once
    hb::set_error("shard exchange: rank %u did not publish its results within %llu ms (step %u); the merged "
                  "results of that step were not written", flag & 0x7fffffffu, x->timeout_ms, x->step);
    return HB_ERR_STATE;
  }
  return HB_OK;
}

int64_t hb_exchange_slice_rows(const hb_exchange_t* xchg) {
  return xchg ? reinterpret_cast<const Exchange*>(xchg)->last_rows : -1;
}

}  // extern "C"
