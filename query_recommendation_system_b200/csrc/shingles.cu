// f-1 — set construction (Recommender.compute_shingles, recommender.py:68-103): the answer set of every
// conjunctive query `attr = value AND ...` over the relational table, as the sets x elements CSR the
// signature kernel consumes (set = query, element = tuple index).
//
// The host dictionary-encodes every column once (value -> int32 code); a query is a row of codes with
// -1 for "no condition on this attribute" (the reference's empty string, recommender.py:83) and -2 for a
// value that does not occur in the column.  One warp per query scans the column-major table, 32 tuples
// per step (coalesced 128-byte reads per attribute), first to count, then to write the ascending tuple ids.
#include "common.cuh"

namespace qrs {

constexpr int kMaxFeatures = 32;

template <bool FILL>
__global__ void __launch_bounds__(256)
answer_sets_kernel(const int *__restrict__ table_cm, int64_t drows, int F, const int *__restrict__ queries,
                   int64_t n_queries, uint32_t *__restrict__ counts, const uint64_t *__restrict__ rowptr,
                   int *__restrict__ colidx) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const unsigned lt = lanemask_lt();
  for (int64_t q = warp_global; q < n_queries; q += nwarps) {
    // the query's conditions, one per lane (F <= 32)
    const int want_mine = lane < F ? queries[q * F + lane] : -1;
    const unsigned cond_mask = __ballot_sync(kFull, want_mine != -1);
    int64_t written = FILL ? int64_t(rowptr[q]) : 0;
    uint32_t total = 0;
    for (int64_t r0 = 0; r0 < drows; r0 += 32) {
      const int64_t r = r0 + lane;
      bool ok = r < drows;
      for (unsigned m = cond_mask; m != 0 && __any_sync(kFull, ok); m &= m - 1) {
        const int f = __ffs(m) - 1;
        const int want = __shfl_sync(kFull, want_mine, f);
        if (ok) ok = table_cm[int64_t(f) * drows + r] == want;
      }
      const unsigned hit = __ballot_sync(kFull, ok);
      if (FILL) {
        if (ok) colidx[written + __popc(hit & lt)] = int(r);
        written += __popc(hit);
      } else {
        total += __popc(hit);
      }
    }
    if (!FILL && lane == 0) counts[q] = total;
  }
}

}  // namespace qrs

using namespace qrs;

extern "C" {

size_t qrs_answer_sets_workspace_bytes(int64_t n_queries) {
  if (n_queries < 0) return 0;
  return round_up<size_t>(size_t(n_queries > 0 ? n_queries : 1) * 4, 256) + scan_workspace_bytes(n_queries);
}

int qrs_answer_sets_count(const int32_t *table_cm, int64_t drows, int32_t F, const int32_t *queries,
                          int64_t n_queries, int64_t *rowptr, void *ws, size_t ws_bytes, void *stream) {
  if (drows < 0 || F <= 0 || n_queries < 0 || !rowptr || (n_queries > 0 && (!queries || (drows > 0 && !table_cm))))
    return QRS_ERR_INVALID_ARG;
  if (F > kMaxFeatures || drows >= (int64_t(1) << 31)) return QRS_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < qrs_answer_sets_workspace_bytes(n_queries)) return QRS_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  uint32_t *counts = static_cast<uint32_t *>(ws);
  void *scan_ws = static_cast<char *>(ws) + round_up<size_t>(size_t(n_queries > 0 ? n_queries : 1) * 4, 256);
  if (n_queries > 0) {
    int64_t blocks = ceil_div<int64_t>(n_queries, 8);
    const int64_t cap = int64_t(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    answer_sets_kernel<false><<<unsigned(blocks), 256, 0, st>>>(table_cm, drows, F, queries, n_queries, counts, nullptr,
                                                                nullptr);
    QRS_LAUNCH_CHECK();
  }
  return exclusive_scan_u32_to_u64(counts, reinterpret_cast<uint64_t *>(rowptr), n_queries, scan_ws,
                                   scan_workspace_bytes(n_queries), st);
}

int qrs_answer_sets_fill(const int32_t *table_cm, int64_t drows, int32_t F, const int32_t *queries, int64_t n_queries,
                         const int64_t *rowptr, int32_t *colidx, void *stream) {
  if (drows < 0 || F <= 0 || n_queries < 0 || !rowptr || (n_queries > 0 && (!queries || (drows > 0 && !table_cm))))
    return QRS_ERR_INVALID_ARG;
  if (F > kMaxFeatures || drows >= (int64_t(1) << 31)) return QRS_ERR_UNSUPPORTED;
  if (n_queries == 0 || drows == 0) return QRS_OK;
  if (!colidx) return QRS_ERR_INVALID_ARG;
  int64_t blocks = ceil_div<int64_t>(n_queries, 8);
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  answer_sets_kernel<true><<<unsigned(blocks), 256, 0, as_stream(stream)>>>(
      table_cm, drows, F, queries, n_queries, nullptr, reinterpret_cast<const uint64_t *>(rowptr), colidx);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // extern "C"
