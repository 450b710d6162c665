/*
 * hbird_b200.h — C-ABI of the B200-native dense nearest-neighbour evaluation path.
 *
 * This is the drop-in boundary for ONE hot path of vpariza/open-hummingbird-eval:
 * memory-bank construction -> exact inner-product kNN -> soft label transfer ->
 * upsample/argmax -> mIoU confusion matrix.  Every entry point names the reference
 * interface it replaces (file:line into the reference tree).  Only plain pointers,
 * sizes and an opaque handle cross this boundary: no torch types.
 *
 * Conventions
 *   - every `*_dev` pointer is a DEVICE pointer on the bank's CUDA device;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - all entry points return an `hb_status` (0 = ok, negative = error); the message of
 *     the last error on the calling thread is returned by hb_last_error();
 *   - all launches are asynchronous on `stream`; no entry point synchronises the device
 *     unless stated;
 *   - there is NO CPU fallback: without an sm_100 device every compute call fails with
 *     HB_ERR_UNSUPPORTED.
 *   - a handle is not thread-safe (the reference drives its backend from one Python
 *     thread, hbird_eval.py:214-246).
 */
#ifndef HBIRD_B200_H_
#define HBIRD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 2

typedef enum hb_status {
  HB_OK = 0,
  HB_ERR_INVALID = -1,     /* bad argument (maps to ValueError, search_faiss.py:25,48) */
  HB_ERR_CUDA = -2,        /* CUDA runtime/driver error (maps to RuntimeError)          */
  HB_ERR_OOM = -3,         /* device allocation failed (maps to MemoryError)            */
  HB_ERR_UNSUPPORTED = -4, /* no sm_100 device (maps to RuntimeError, search_faiss.py:15-16) */
  HB_ERR_STATE = -5        /* call order violated, e.g. search before finalize          */
} hb_status;

/* hb_bank_create flags */
#define HB_BANK_KEEP_F32 1u   /* keep an fp32 copy of the normalised rows for the exact re-rank */
#define HB_BANK_L2 2u         /* squared-L2 metric (GpuIndexFlatL2, search_faiss.py:45-46): hb_search
                                 returns ||q-x||^2 ascending; rows are NOT expected to be unit-norm  */

typedef struct hb_bank hb_bank_t; /* opaque: one HBM-resident shard of the memory bank */
typedef struct hb_exchange hb_exchange_t; /* opaque: one rank's end of the fused shard exchange */
#define HB_EXCHANGE_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t) */

/* ---- library ------------------------------------------------------------------- */

int hb_abi_version(void);
/* Message of the last failing call on this thread ("" if none). */
const char* hb_last_error(void);
/* HB_OK iff `device` exists and is compute capability 10.x (replaces the
 * faiss.get_num_gpus() < 1 check, search_faiss.py:14-16).  Writes the SM count. */
int hb_device_check(int device, int* num_sms_out);

/* ---- K1: memory-bank construction ------------------------------------------------
 * Replaces hbird_eval.py:309-329 (mask decode, _patchify_gt :554-573, one_hot+mean
 * :319-320, L2 normalise :324, append :328-329) and the faiss index build + index.add
 * (search_faiss.py:50-81).  The bank owns its HBM: bf16 rows (row pitch padded to a
 * multiple of 64 elements, zero filled), optionally an fp32 copy, and one label record
 * per row: the per-patch class histogram as uint16[num_classes] (the reference's soft
 * label is histogram / patch_pixels, hbird_eval.py:319-320). */
int hb_bank_create(int device, int d, int num_classes, int patch_pixels,
                   int64_t capacity_rows, unsigned flags, hb_bank_t** bank_out);
int hb_bank_destroy(hb_bank_t* bank);

/* Append n rows.  feats_dev: fp32 (src_rows, d) raw ViT features of B images with an
 * S x S patch grid (src_rows = B*S*S), NOT normalised.  mask_dev: uint8 (B, S*ps, S*ps)
 * class ids already decoded (hb_decode_mask) — patch (b, py, px) covers pixels
 * [py*ps, (py+1)*ps) x [px*ps, (px+1)*ps).  sel_dev: optional int32 (n,) source-row
 * indices (the bounded-memory sampler's picks, hbird_eval.py:332-355); NULL = all
 * src_rows rows in order (then n must equal B*S*S).  Row i of the bank receives
 * feats[src]/||feats[src]||_2 (no epsilon, hbird_eval.py:324). */
int hb_bank_append(hb_bank_t* bank, const float* feats_dev, const uint8_t* mask_dev,
                   int B, int S, int ps, const int32_t* sel_dev, int64_t n, void* stream);

/* Append n rows whose soft labels are already known as fp32 (n, num_classes) rows that
 * are multiples of 1/patch_pixels (load_memory path, hbird_eval.py:380-400): counts are
 * recovered as rint(label*patch_pixels).  normalise != 0 re-normalises the features. */
int hb_bank_append_soft(hb_bank_t* bank, const float* feats_dev, const float* soft_dev,
                        int64_t n, int normalise, void* stream);

/* Bounded-memory sampler (replaces _sample_features, hbird_eval.py:447-517): for each of the B
 * images pick the K patches with the smallest score*U, score = sum over the classes present in the
 * patch of the number of patches of the image containing that class (1e6 for an empty patch).
 * mask_dev: uint8 (B, S*ps, S*ps) decoded class ids; uniform_dev: fp32 (B, S*S) = the reference's CPU
 * RNG stream (torch.rand, image order) uploaded by the caller, so a seeded run picks the same
 * patches; sel_out_dev: int32 (B, K) flat source rows b*S*S + patch, ascending score — ready to be
 * passed to hb_bank_append as sel_dev. */
int hb_sample_patches(const uint8_t* mask_dev, int B, int S, int ps, int C, const float* uniform_dev,
                      int K, int32_t* sel_out_dev, void* stream);

/* Freeze the bank (builds the TMA tensor map).  Must precede hb_search. */
int hb_bank_finalize(hb_bank_t* bank);
int64_t hb_bank_rows(const hb_bank_t* bank);
int64_t hb_bank_capacity(const hb_bank_t* bank);
/* Device pointer to the uint16 (rows, num_classes) label histogram table. */
const uint16_t* hb_bank_label_table(const hb_bank_t* bank);
/* Export rows [row0, row0+n) as the reference's tensors: feats_out_dev fp32 (n, d)
 * unit-norm rows (fp32 copy if kept, else widened bf16) and/or labels_out_dev fp32
 * (n, num_classes) soft labels (feature_memory / label_memory, hbird_eval.py:357-366).
 * Either output may be NULL. */
int hb_bank_export(const hb_bank_t* bank, int64_t row0, int64_t n, float* feats_out_dev,
                   float* labels_out_dev, void* stream);

/* ---- K2 + K2b: search ----------------------------------------------------------------
 * Replaces NearestNeighborSearchFaiss.find_nearest_neighbors (search_faiss.py:83-90),
 * i.e. GpuIndexFlatIP.search: exact top-k by inner product of the raw (un-normalised)
 * queries against the unit-norm bank rows, sorted by descending score.
 * q_dev: fp32 (Q, d).  k <= k_prime, k_prime in {32, 64, 128}: the tcgen05 bf16 pass keeps
 * candidates per query, the fp32 pass re-scores them exactly and keeps k.
 * out_scores_dev fp32 (Q, k); out_idx_dev int64 (Q, k) = row index + idx_offset (the
 * shard's first global row); out_qnorm_dev fp32 (Q,) = ||q||_2 or NULL.  If the bank
 * holds fewer than k rows the tail is (-inf, -1), as faiss pads.
 *
 * What the bf16 pass guarantees (csrc/search.cu).  A query's bank rows are scanned by L >= 4
 * candidate lists of k_prime/2 entries each (two per bank chunk, at least two chunks; which of
 * a chunk's two lists sees a given 32-row group changes pseudo-randomly from tile to tile).
 * A list drops a row only (a) below its own (k_prime/2)-th best score or (b) below a shared
 * threshold x for which k_prime rows scoring >= x are known to exist (2 lists holding k_prime/2
 * each, or 4 lists holding k_prime/4 each).  Hence:
 *   - the query's best k_prime/2 rows by bf16 score are ALWAYS among the candidates;
 *   - a row ranked r <= k_prime is missing only if k_prime/2 better rows share its list; for
 *     rows spread over the lists independently of their score that is P[Bin(r-1, 1/L) >=
 *     k_prime/2]: for k_prime = 64, L = 4: 6e-10 at r = 48 and 1e-5 at r = 64; L >= 6: < 1e-9;
 *   - banks of at most 16384 rows always run with k_prime = 128 lists (64 entries each).
 * The fp32 top-k is a subset of the bf16 top-k_prime only statistically (bf16 rounding moves a
 * row by a few ranks; SURVEY.md H1 measured the fp32 top-30 inside the bf16 top-64 always and
 * inside the bf16 top-30 for 99.6 % of the neighbours).  Callers that need the bf16 top-64
 * strictly (k up to 64, or adversarial row placement) pass k_prime = 128. */
int hb_search(hb_bank_t* bank, const float* q_dev, int64_t Q, int k, int k_prime,
              int64_t idx_offset, float* out_scores_dev, int64_t* out_idx_dev,
              float* out_qnorm_dev, void* stream);

/* K2 + K2b with K4a fused into the re-rank warp: as hb_search, and the warp that holds a query's k
 * exact neighbours also writes label_hat[q] (see hb_label_transfer) — the neighbour list is never
 * re-read from HBM.  label_table_dev: uint16 (table_rows, C) indexed by row + idx_offset (NULL =
 * the bank's own table, table_rows ignored); out_scores_dev / out_idx_dev may both be NULL when
 * only label_hat is wanted; out_label_hat_dev fp32 (Q, C).  Inner-product banks only. */
int hb_search_transfer(hb_bank_t* bank, const uint16_t* label_table_dev, int64_t table_rows,
                       const float* q_dev, int64_t Q, int k, int k_prime, int64_t idx_offset,
                       float beta, float* out_scores_dev, int64_t* out_idx_dev,
                       float* out_qnorm_dev, float* out_label_hat_dev, void* stream);

/* Split form of hb_search / hb_search_transfer for software pipelining across batches.  begin =
 * query prep + K2 (the tensor-core pass) into pipeline slot `slot` (0 or 1) on `stream`; finish = K2b
 * (exact re-rank, outputs as hb_search / hb_search_transfer: give out_scores/out_idx, out_label_hat,
 * or all three) from that slot on ANY stream that is ordered after the begin (event).  With the
 * finish of batch i on a second stream and the begin of batch i+1 on a higher-priority one, the
 * post-processing fills the SMs the search CTAs vacate as they finish and the launch gaps between two
 * search kernels (it is not co-resident with the default search build: hb_coresidency_config).  A
 * slot is reused only after its finish has run (enforced with an event inside the library); q_dev
 * must stay valid until the finish has executed; the scratch block grows in stream order.  prepared_event: optional `cudaEvent_t` recorded between the query prep and
 * K2 — a pipelined caller makes the previous batch's finish wait for it, so that the small kernels
 * are released together with the search kernel (whose stream should have the higher priority)
 * instead of flooding the SMs in the gap before it. */
int hb_search_begin(hb_bank_t* bank, const float* q_dev, int64_t Q, int k_prime, int slot,
                    float* out_qnorm_dev, void* prepared_event, void* stream);
int hb_search_finish(hb_bank_t* bank, int slot, const float* q_dev, int k, int64_t idx_offset,
                     const uint16_t* label_table_dev, int64_t table_rows, float beta,
                     float* out_scores_dev, int64_t* out_idx_dev, float* out_label_hat_dev,
                     void* stream);

/* Forget searches that were begun and will not be finished (error recovery of a pipelined caller):
 * both slots become free again.  Host-side state only; work already in the streams runs to its end. */
int hb_search_abort(hb_bank_t* bank);

/* One validation batch through the whole path in 4 launches (query prep, K2, K2b+K4a, fused tail):
 * replaces hbird_eval.py:217-252 for a bank that is not row-sharded.  q_dev fp32 (B*S*S, d) raw
 * features; y_dev fp32 (B, H, W) = class id / 255 (loader contract, :219); label_hat_dev fp32
 * (B*S*S, C) scratch that holds label_hat on return; conf_dev int64 (C, C) accumulated in place;
 * out_pred_dev uint8 (B, H, W) or NULL; out_scores_dev / out_idx_dev (B*S*S, k) or both NULL.
 * Capturable in a CUDA graph once the bank's scratch has been sized by a first call. */
int hb_eval_step(hb_bank_t* bank, const uint16_t* label_table_dev, int64_t table_rows,
                 const float* q_dev, int B, int S, int H, int W, const float* y_dev, int k,
                 int k_prime, int64_t idx_offset, float beta, int ignore_index,
                 float* label_hat_dev, int64_t* conf_dev, uint8_t* out_pred_dev,
                 float* out_scores_dev, int64_t* out_idx_dev, void* stream);

/* Tuning/diagnostics for hb_search: cta_group (1 or 2; 0 = library default),
 * max_chunks (bank split per query block; 0 = auto). */
int hb_search_config(hb_bank_t* bank, int cta_group, int max_chunks);
/* L2 prefetch distance (in 256-row bank tiles) of the search kernel's TMA producer; 0 = off,
 * -1 = library default.
 * ablate is a MEASUREMENT-ONLY switch (results are wrong when it is non-zero): 1 = the epilogue
 * releases accumulators unread (pure GEMM pipeline), 2 = it scans but never inserts. */
int hb_search_tune(hb_bank_t* bank, int prefetch_tiles, int ablate);
/* Co-residency of the post-processing with a RUNNING search kernel (hb_search_begin / _finish on
 * two streams).  A CTA becomes resident on an SM only if every one of its warps finds registers on
 * the sub-partition it is assigned to (warp id mod 4), and the search CTA's ten 168-register warps
 * fill two of the four sub-partitions.  lean_search != 0 selects a 128-register build of the search
 * kernel (k' = 64, CTA pairs), which leaves 4096 registers free on every sub-partition;
 * rerank_warps_per_cta = queries per K2b CTA (1, 2 or 4; 0 = default 4); rerank_shared_carveout =
 * preferred shared-memory carve-out of K2b in per cent (-1 = driver default; 100 = the search
 * kernel's own configuration, so that an SM need not drain to switch). */
int hb_coresidency_config(hb_bank_t* bank, int lean_search, int rerank_warps_per_cta, int rerank_shared_carveout);
/* Diagnostics: when stats_dev (device, zeroed, 148*8*8 uint64) is non-NULL, hb_search runs an
 * instrumented build that accumulates per-epilogue-warp cycle counters {wait, tmem load, slow path,
 * post-release folds, slow-path entries, folds, tiles, -}.  NULL switches it off. */
int hb_search_stats(hb_bank_t* bank, unsigned long long* stats_dev);
/* L2 pacing window of the search kernel (default on): CTAs streaming the same bank chunk stay within
 * ~100 tiles of each other so each tile is read from HBM once per wave. */
int hb_search_pacing(hb_bank_t* bank, int enable);
/* Number of kernel launches the last hb_search on this bank issued. */
int hb_search_last_launches(const hb_bank_t* bank);

/* Kernel timing for roofline reports: when enabled, every hb_search brackets its tcgen05
 * GEMM+top-k kernel with CUDA events on the launching stream (a ring of 64 pairs).
 * hb_search_kernel_time synchronises those events and returns the mean duration in ms of the
 * kernels recorded since timing was (re-)enabled, and how many there were. */
int hb_search_timing(hb_bank_t* bank, int enable);
int hb_search_kernel_time(hb_bank_t* bank, float* mean_ms_out, int* count_out);
/* Same, for the kernel that follows it in the same searches: K2b (exact re-rank, with whatever is
 * fused into it: label transfer, scatter into the exchange windows). */
int hb_search_rerank_time(hb_bank_t* bank, float* mean_ms_out, int* count_out);

/* Host-only (no GPU needed): the work decomposition hb_search would use for a bank of `rows`
 * rows and Q queries on `num_sms` SMs.  out4 = {n_tiles, n_qblocks, n_chunks, n_units}; chunk c
 * covers tiles [n_tiles*c/n_chunks, n_tiles*(c+1)/n_chunks) of 256 bank rows. */
int hb_plan_search(int64_t rows, int64_t Q, int cta_group, int num_sms, int max_chunks, int* out4);

/* Debug/validation: full bf16-input fp32-accumulate score matrix of the tcgen05 pass,
 * out_dev fp32 (Q, rows).  Small problems only (Q*rows*4 bytes are written). */
int hb_search_dump_scores(hb_bank_t* bank, const float* q_dev, int64_t Q, float* out_dev,
                          int cta_group, void* stream);

/* ---- K3: cross-shard merge -------------------------------------------------------------
 * Replaces the host-side merge of faiss.IndexShards (search_faiss.py:53-63,89).
 * shard_scores_dev fp32 (G, Q, k) and shard_idx_dev int64 (G, Q, k) are the all-gathered
 * per-shard results (each sorted descending, global indices).  Writes the top-k of the
 * union per query, sorted descending (ties: smaller index first). */
int hb_merge_topk(const float* shard_scores_dev, const int64_t* shard_idx_dev, int G,
                  int64_t Q, int k, float* out_scores_dev, int64_t* out_idx_dev,
                  void* stream);

/* K3 with K4a fused (see hb_search_transfer): qnorm_dev fp32 (Q,), out_label_hat_dev fp32 (Q, C);
 * out_scores_dev / out_idx_dev may both be NULL.  Inner-product results only (the merge keeps the
 * largest scores). */
int hb_merge_topk_transfer(const float* shard_scores_dev, const int64_t* shard_idx_dev, int G,
                           int64_t Q, int k, const uint16_t* label_table_dev, int64_t table_rows,
                           int C, int patch_pixels, const float* qnorm_dev, float beta,
                           float* out_scores_dev, int64_t* out_idx_dev,
                           float* out_label_hat_dev, void* stream);

/* ---- K3x: fused shard exchange over NVLink peer memory -------------------------------------
 * The B200 form of faiss.IndexShards' search (search_faiss.py:53-63,89) for one process per GPU:
 * every rank searches all Q queries against its row shard; rank p post-processes the queries
 * [qsplit[p], qsplit[p+1]).  hb_search_scatter runs K2 + K2b and stores each query's shard top-k
 * DIRECTLY into the owner rank's window (peer memory mapped through CUDA IPC, NVLink stores), then
 * publishes a step counter there; hb_exchange_merge waits for the counters of all ranks and merges
 * the G lists of the local slice.  No NCCL call and no host synchronisation on the data path.
 * All ranks must make the same sequence of scatter/merge calls (one merge per scatter).
 *
 * Set-up: each rank creates its window (slice_capacity = most rows any rank will own per call,
 * max_k = largest k), exports a 64-byte handle, the host side all-gathers the handles (any
 * transport; hbird_b200 uses torch.distributed) and every rank connects.  world == 1 needs no
 * connect.  HB_ERR_UNSUPPORTED from connect = no peer access / no CUDA IPC: use the NCCL
 * all-gather + hb_merge_topk path instead.  hb_exchange_connect_local wires exchanges created in
 * ONE process on one device to each other by pointer (tests; the ranks then run one after the
 * other on a stream: all scatters of a step before its merges). */
int hb_exchange_create(int device, int rank, int world, int64_t slice_capacity, int max_k,
                       hb_exchange_t** out);
int hb_exchange_destroy(hb_exchange_t* xchg);
int hb_exchange_handle(hb_exchange_t* xchg, void* handle_out, int handle_bytes);
int hb_exchange_connect(hb_exchange_t* xchg, const void* handles, int n_handles);
int hb_exchange_connect_local(hb_exchange_t* xchg, hb_exchange_t* const* peers, int n_peers);
/* Orderly teardown across processes: every rank disconnects (unmaps its peers' windows), the host
 * side runs a barrier, then every rank destroys (frees its own window).  hb_exchange_destroy alone
 * disconnects first, which is enough when the peers are already gone. */
int hb_exchange_disconnect(hb_exchange_t* xchg);
/* qsplit_host: world+1 ascending int64 on the HOST, qsplit[0] = 0, qsplit[world] = Q.
 * out_qnorm_dev: optional fp32 (Q,) query norms, as hb_search. */
int hb_search_scatter(hb_bank_t* bank, hb_exchange_t* xchg, const float* q_dev, int64_t Q, int k,
                      int k_prime, int64_t idx_offset, const int64_t* qsplit_host,
                      float* out_qnorm_dev, void* stream);
/* hb_search_scatter's second half for a search started with hb_search_begin (see there): K2b of
 * pipeline slot `slot` with the scatter into the owner ranks' windows, on `stream`. */
int hb_search_finish_scatter(hb_bank_t* bank, hb_exchange_t* xchg, int slot, const float* q_dev, int k,
                             int64_t idx_offset, const int64_t* qsplit_host, void* stream);
/* Outputs: fp32 / int64 (rows, k) for this rank's slice of the last scatter, sorted descending,
 * global indices; rows = hb_exchange_slice_rows().  A peer that never arrives makes the kernel give
 * up after the exchange's timeout (see hb_exchange_status) instead of hanging the GPU. */
int hb_exchange_merge(hb_exchange_t* xchg, float* out_scores_dev, int64_t* out_idx_dev, void* stream);
/* hb_exchange_merge with K4a fused into the merging warp: qnorm_slice_dev fp32 (rows,) are the norms
 * of this rank's query slice, out_label_hat_dev fp32 (rows, C); out_scores_dev / out_idx_dev may
 * both be NULL.  Counts as the one merge of the last scatter. */
int hb_exchange_merge_transfer(hb_exchange_t* xchg, const uint16_t* label_table_dev, int64_t table_rows,
                               int C, int patch_pixels, const float* qnorm_slice_dev, float beta,
                               float* out_scores_dev, int64_t* out_idx_dev,
                               float* out_label_hat_dev, void* stream);
int64_t hb_exchange_slice_rows(const hb_exchange_t* xchg);
/* Failure handling of the exchange.  A merge kernel waits for its peers at most `timeout_ms` (default
 * 600 000); when a peer does not arrive it records that rank and returns without writing its outputs
 * — no trap, the CUDA context stays usable.  hb_exchange_status synchronises `stream` and returns
 * HB_ERR_STATE (message: which rank, which step) if a merge since the last call timed out, HB_OK
 * otherwise. */
int hb_exchange_set_timeout(hb_exchange_t* xchg, int64_t timeout_ms);
/* How much every shard re-ranks.  mode 0 (default): each shard re-ranks its own whole bf16 top-k'
 * (what faiss.IndexShards' per-shard exact search amounts to), one exchange hop.  mode 1, the
 * THRESHOLD EXCHANGE: hb_search_scatter / hb_search_finish_scatter stop after a first kernel that
 * merges a query's chunk lists into its sorted bf16 top-k' and stores four order statistics of it
 * (the scores at ranks k', k'/2, k'/4, k'/8) into every rank's window; once all peers' statistics
 * have arrived (second flag array, same bounded wait), every shard derives the SAME lower bound of
 * the global k'-th best bf16 score — if j shards hold k'/j candidates >= x each, k' rows score >= x —
 * and gathers fp32 rows only for its candidates at or above it (on average ~k'/G + a few instead of
 * k' per query and shard: the G-fold redundant row gather of mode 0 disappears, and the candidate
 * set is a superset of the unsharded search's).  That second phase is issued by hb_exchange_rerank,
 * or implicitly by the next hb_exchange_merge*, on the stream of THAT call (which must be the
 * scatter's stream or ordered after it); the bank and the query buffer given to the scatter must stay
 * alive until then.  All ranks must use the same mode; with world = 1 the mode is moot. */
int hb_exchange_config(hb_exchange_t* xchg, int mode);
/* mode 1 only: issue phase 2 (wait for the peers' statistics, re-rank the survivors, scatter the
 * results) of the exchange begun by the last scatter call.  No-op when nothing is pending. */
int hb_exchange_rerank(hb_exchange_t* xchg, void* stream);
int hb_exchange_status(hb_exchange_t* xchg, void* stream);

/* ---- K4: label transfer ------------------------------------------------------------------
 * Replaces the neighbour gather (hbird_eval.py:611-637) and _cross_attention
 * (hbird_eval.py:575-609): label_hat[q] = sum_j softmax_j(cos(q, m_j)/beta) * soft_label[idx_j]
 * with cos(q, m_j) = score_j / ||q|| (bank rows are unit norm), soft_label = hist/patch_pixels.
 * label_table_dev: uint16 (table_rows, C) indexed by the (global) indices in idx_dev.
 * Entries with idx < 0 are skipped.  out_label_hat_dev fp32 (Q, C). */
int hb_label_transfer(const uint16_t* label_table_dev, int64_t table_rows, int C,
                      int patch_pixels, const float* scores_dev, const int64_t* idx_dev,
                      const float* qnorm_dev, int64_t Q, int k, float beta,
                      float* out_label_hat_dev, void* stream);

/* Replaces hbird_eval.py:235-243: label_hat (B, S*S, C) viewed as (B, C, S, S),
 * F.interpolate(size=(H, W), mode="bilinear", align_corners=False), argmax over C
 * (first maximum wins).  out_pred_dev uint8 (B, H, W). */
int hb_upsample_argmax(const float* label_hat_dev, int B, int S, int C, int H, int W,
                       uint8_t* out_pred_dev, void* stream);

/* Fused tail — replaces hbird_eval.py:219 (mask decode), :235-243 (upsample + argmax) and
 * PredsmIoU.update (eval_metrics.py:73-109) in one pass over the output pixels: per (image, band of
 * rows) the label_hat cells are staged in shared memory, every pixel is interpolated with torch's
 * arithmetic, its argmax is scored against the ground truth in a shared-memory histogram, and
 * neither the decoded mask nor the prediction map has to be written.  Ground truth: y_dev fp32
 * (B, H, W) = id/255, or gt_dev uint8 (B, H, W) already decoded (y_dev wins; both NULL only with
 * conf_dev NULL).  conf_dev int64 (C, C) accumulated in place or NULL; out_pred_dev uint8
 * (B, H, W) or NULL. */
int hb_predict_score(const float* label_hat_dev, int B, int S, int C, int H, int W,
                     const float* y_dev, const uint8_t* gt_dev, int ignore_index,
                     int64_t* conf_dev, uint8_t* out_pred_dev, void* stream);

/* ---- K5: scoring --------------------------------------------------------------------------
 * Replaces the loader-contract decode `(y*255).long()` (hbird_eval.py:219,309-310):
 * out[i] = (uint8) trunc(y[i]*255); remap_255_to_0 applies `y[y==255]=0` (bank side only). */
int hb_decode_mask(const float* y_dev, int64_t n, int remap_255_to_0, uint8_t* out_dev,
                   void* stream);

/* Replaces PredsmIoU.update (eval_metrics.py:73-109): for every pixel with
 * gt != ignore_index (ignore_index < 0: none), gt < C_gt and pred < C_pred,
 * conf[gt*C_pred + pred] += 1.  conf_dev: int64 (C_gt, C_pred), accumulated in place. */
int hb_confusion_accumulate(const uint8_t* gt_dev, const uint8_t* pred_dev, int64_t n,
                            int C_gt, int C_pred, int ignore_index, int64_t* conf_dev,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HBIRD_B200_H_ */
