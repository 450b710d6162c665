// Work decomposition of the search: pure integer host logic (also compiled by g++ in the CPU tests).
//
// The (queries x bank) rectangle is cut into work items = (query block of 128*CG rows) x (bank chunk
// of consecutive 256-row tiles).  Items are dealt round-robin to the persistent CTAs / CTA pairs,
// chunk-major, so CTAs that run concurrently stream the SAME bank tiles (they hit in the 126 MB L2)
// against different query blocks.  n_chunks is chosen to fill whole waves of the machine.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define HB_HD __host__ __device__
#else
#define HB_HD
#endif

namespace hb {

struct SearchPlan {
  int n_tiles;    // ceil(rows / 256)
  int n_qblocks;  // ceil(Q / (128 * cta_group))
  int n_chunks;   // bank chunks per query block
  int n_units;    // persistent CTAs (cta_group 1) or CTA pairs (cta_group 2)
};

// First tile of chunk c when n_tiles tiles are split into n_chunks balanced chunks.
HB_HD inline int chunk_tile_begin(int n_tiles, int n_chunks, int c) {
  return static_cast<int>((static_cast<int64_t>(n_tiles) * c) / n_chunks);
}

inline SearchPlan plan_search(int64_t rows, int64_t Q, int cta_group, int num_sms, int max_chunks) {
  SearchPlan p;
  p.n_tiles = static_cast<int>((rows + 255) / 256);
  if (p.n_tiles < 1) p.n_tiles = 1;
  const int64_t qrows = 128 * static_cast<int64_t>(cta_group);
  p.n_qblocks = static_cast<int>((Q + qrows - 1) / qrows);
  if (p.n_qblocks < 1) p.n_qblocks = 1;
  p.n_units = num_sms / cta_group;
  if (p.n_units < 1) p.n_units = 1;
  int cap = max_chunks > 0 ? max_chunks : 64;
  if (cap > p.n_tiles) cap = p.n_tiles;
  // Prefer few, long chunks (tighter running thresholds, fewer candidate lists); take more only
  // when that fills the last wave noticeably better.  At least two chunks (= four candidate lists
  // per query) unless the caller caps it: the top-k' bound of the search rests on the query's best
  // rows being spread over several lists (search.cu).
  int first = (cap >= 2 && max_chunks != 1) ? 2 : 1;
  int best = first;
  double best_eff = 0.0;
  for (int c = first; c <= cap; ++c) {
    const int64_t items = static_cast<int64_t>(p.n_qblocks) * c;
    const int64_t waves = (items + p.n_units - 1) / p.n_units;
    // each item costs ~n_tiles/c tiles; the job takes waves * ceil(n_tiles/c) tile-times
    const int64_t tiles_per_item = (p.n_tiles + c - 1) / c;
    const double ideal = static_cast<double>(p.n_qblocks) * p.n_tiles / p.n_units;
    const double eff = ideal / static_cast<double>(waves * tiles_per_item);
    if (eff > best_eff * 1.02) {
      best_eff = eff;
      best = c;
    }
  }
  p.n_chunks = best;
  return p;
}

}  // namespace hb
