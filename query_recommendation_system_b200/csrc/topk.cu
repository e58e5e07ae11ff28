// a7 + a9 — neighbour tables (recommender.py:183-210): adjacency lists built from the scored
// candidate pairs by counting + scan + fill, then one warp per row keeps the k best
// (similarity descending, neighbour id ascending) in registers with a 64-wide bitonic
// network — two keys per lane, never touching shared or local memory.
#include "common.cuh"
#include "bitonic.cuh"

namespace qrs {

__device__ __forceinline__ uint64_t adj_key(int milli, uint32_t nb) {
  return (uint64_t(uint32_t(1000 - milli)) << 32) | uint64_t(nb);
}

__global__ void __launch_bounds__(256)
degree_kernel(const uint64_t *__restrict__ pairs, int64_t n_pairs, int64_t row_base, int64_t n_rows,
              uint32_t *__restrict__ deg) {
  for (int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < n_pairs; p += int64_t(gridDim.x) * blockDim.x) {
    const uint64_t pr = ld_stream_u64(pairs + p);
    const int64_t i = int64_t(pr >> 32) - row_base, j = int64_t(pr & 0xFFFFFFFFu) - row_base;
    if (i >= 0 && i < n_rows) atomicAdd(deg + i, 1u);
    if (j >= 0 && j < n_rows) atomicAdd(deg + j, 1u);
  }
}

__global__ void __launch_bounds__(256)
fill_kernel(const uint64_t *__restrict__ pairs, const int *__restrict__ sim_milli, int64_t n_pairs,
            int64_t row_base, int64_t n_rows, const uint64_t *__restrict__ row_off, uint32_t *__restrict__ cursor,
            uint64_t *__restrict__ adj) {
  for (int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < n_pairs; p += int64_t(gridDim.x) * blockDim.x) {
    const uint64_t pr = ld_stream_u64(pairs + p);
    const int milli = sim_milli[p];
    const uint32_t gi = uint32_t(pr >> 32), gj = uint32_t(pr);
    const int64_t i = int64_t(gi) - row_base, j = int64_t(gj) - row_base;
    if (i >= 0 && i < n_rows) adj[row_off[i] + atomicAdd(cursor + i, 1u)] = adj_key(milli, gj);
    if (j >= 0 && j < n_rows) adj[row_off[j] + atomicAdd(cursor + j, 1u)] = adj_key(milli, gi);
  }
}

__global__ void __launch_bounds__(256)
select_kernel(const uint64_t *__restrict__ adj, const uint64_t *__restrict__ row_off, int64_t n_rows, int k,
              int *__restrict__ idx, double *__restrict__ val, int *__restrict__ milli, int *__restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const uint64_t INF = ~uint64_t(0);
  for (int64_t row = warp_global; row < n_rows; row += nwarps) {
    const uint64_t lo = row_off[row], hi = row_off[row + 1];
    uint64_t h0 = INF, h1 = INF;
    for (uint64_t base = lo; base < hi; base += 64) {
      uint64_t n0 = base + lane < hi ? adj[base + lane] : INF;
      uint64_t n1 = base + 32 + lane < hi ? adj[base + 32 + lane] : INF;
      bitonic_keep_lowest64(h0, h1, n0, n1, lane);
    }
    const uint64_t deg = hi - lo;
    const int c = deg < uint64_t(k) ? int(deg) : k;
    if (lane == 0) cnt[row] = c;
    if (lane < k) {
      const bool ok = lane < c;
      idx[row * k + lane] = ok ? int(uint32_t(h0)) : -1;
      val[row * k + lane] = ok ? double(1000 - int(uint32_t(h0 >> 32))) / 1000.0 : 0.0;
      if (milli) milli[row * k + lane] = ok ? 1000 - int(uint32_t(h0 >> 32)) : 0;
    }
    if (lane + 32 < k) {
      const bool ok = lane + 32 < c;
      idx[row * k + lane + 32] = ok ? int(uint32_t(h1)) : -1;
      val[row * k + lane + 32] = ok ? double(1000 - int(uint32_t(h1 >> 32))) / 1000.0 : 0.0;
      if (milli) milli[row * k + lane + 32] = ok ? 1000 - int(uint32_t(h1 >> 32)) : 0;
    }
  }
}

}  // namespace qrs

using namespace qrs;

extern "C" {

size_t qrs_topk_workspace_bytes(int64_t n_rows) {
  if (n_rows < 0) return 0;
  const size_t nr = size_t(n_rows > 0 ? n_rows : 1);
  return 2 * round_up<size_t>(nr * 4, 256) + round_up<size_t>((nr + 1) * 8, 256) + scan_workspace_bytes(n_rows);
}

int qrs_topk_tables(const uint64_t *pairs, const int32_t *sim_milli, int64_t n_pairs, int64_t row_base,
                    int64_t n_rows, int32_t k, uint64_t *adj, void *ws, size_t ws_bytes, int32_t *idx, double *val,
                    int32_t *milli, int32_t *cnt, void *stream) {
  if (n_pairs < 0 || n_rows < 0 || k <= 0 || !idx || !val || !cnt) return QRS_ERR_INVALID_ARG;
  if (n_pairs > 0 && (!pairs || !sim_milli || !adj)) return QRS_ERR_INVALID_ARG;
  if (k > QRS_MAX_TOPK) return QRS_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < qrs_topk_workspace_bytes(n_rows)) return QRS_ERR_WORKSPACE;
  if (n_rows == 0) return QRS_OK;
  cudaStream_t st = as_stream(stream);
  const size_t nr = size_t(n_rows);
  char *w = static_cast<char *>(ws);
  uint32_t *deg = reinterpret_cast<uint32_t *>(w);
  w += round_up<size_t>(nr * 4, 256);
  uint32_t *cursor = reinterpret_cast<uint32_t *>(w);
  w += round_up<size_t>(nr * 4, 256);
  uint64_t *row_off = reinterpret_cast<uint64_t *>(w);
  w += round_up<size_t>((nr + 1) * 8, 256);
  void *scan_ws = w;

  QRS_CUDA(cudaMemsetAsync(deg, 0, 2 * round_up<size_t>(nr * 4, 256), st));   // deg and cursor are adjacent
  const int64_t cap = int64_t(sm_count()) * 8;
  if (n_pairs > 0) {
    int64_t blocks = ceil_div<int64_t>(n_pairs, 256);
    if (blocks > cap) blocks = cap;
    degree_kernel<<<unsigned(blocks), 256, 0, st>>>(pairs, n_pairs, row_base, n_rows, deg);
    QRS_LAUNCH_CHECK();
  }
  int rc = exclusive_scan_u32_to_u64(deg, row_off, n_rows, scan_ws, scan_workspace_bytes(n_rows), st);
  if (rc != QRS_OK) return rc;
  if (n_pairs > 0) {
    int64_t blocks = ceil_div<int64_t>(n_pairs, 256);
    if (blocks > cap) blocks = cap;
    fill_kernel<<<unsigned(blocks), 256, 0, st>>>(pairs, sim_milli, n_pairs, row_base, n_rows, row_off, cursor, adj);
    QRS_LAUNCH_CHECK();
  }
  int64_t blocks = ceil_div<int64_t>(n_rows, 8);
  if (blocks > cap) blocks = cap;
  select_kernel<<<unsigned(blocks), 256, 0, st>>>(adj, row_off, n_rows, k, idx, val, milli, cnt);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ f-2: per-user top-N recommendations
// Replaces what Recommender.top_k_queries serves (recommender.py:345-374: among the just-predicted cells of a
// user with prediction != 0, the k largest predictions) without materialising a dense users x queries matrix:
// one warp per user walks the user's sparse prediction slices and keeps the best 64 (prediction descending,
// column ascending — the reference's order inside ties comes from an unstable argsort and is unspecified).
namespace qrs {

__global__ void __launch_bounds__(256)
topn_kernel(const int64_t *__restrict__ row_start, const int *__restrict__ row_cnt, const void *__restrict__ cols,
            int cols_u16, const uint8_t *__restrict__ preds, int64_t n_users, int nt, int N,
            int *__restrict__ out_idx, uint8_t *__restrict__ out_val, int *__restrict__ out_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const uint64_t INF = ~uint64_t(0);
  for (int64_t u = warp_global; u < n_users; u += nwarps) {
    uint64_t h0 = INF, h1 = INF;
    int64_t seen = 0;
    for (int t = 0; t < nt; ++t) {
      const int64_t lo = row_start[u * nt + t];
      const int n = row_cnt[u * nt + t];
      for (int base = 0; base < n; base += 64) {
        uint64_t k0 = INF, k1 = INF;
        {
          const int x = base + lane;
          if (x < n) {
            const unsigned p = preds[lo + x];
            const unsigned c = cols_u16 ? unsigned(static_cast<const uint16_t *>(cols)[lo + x])
                                        : unsigned(static_cast<const int *>(cols)[lo + x]);
            if (p != 0) { k0 = (uint64_t(255u - p) << 32) | c; }
          }
          const int y = x + 32;
          if (y < n) {
            const unsigned p = preds[lo + y];
            const unsigned c = cols_u16 ? unsigned(static_cast<const uint16_t *>(cols)[lo + y])
                                        : unsigned(static_cast<const int *>(cols)[lo + y]);
            if (p != 0) { k1 = (uint64_t(255u - p) << 32) | c; }
          }
        }
        seen += __popc(__ballot_sync(kFull, k0 != INF)) + __popc(__ballot_sync(kFull, k1 != INF));
        // only keys better than the 64th best kept so far matter: a few are inserted one by one, many (the
        // first chunks, while the list fills) go through the sort + merge network
        const uint64_t worst = __shfl_sync(kFull, h1, 31);
        unsigned b0 = __ballot_sync(kFull, k0 < worst), b1 = __ballot_sync(kFull, k1 < worst);
        if ((b0 | b1) == 0u) continue;
        if (__popc(b0) + __popc(b1) > 8) {
          bitonic_keep_lowest64(h0, h1, k0, k1, lane);
        } else {
          for (; b0 != 0u; b0 &= b0 - 1u) {
            const uint64_t key = __shfl_sync(kFull, k0, __ffs(b0) - 1);
            if (key < __shfl_sync(kFull, h1, 31)) sorted_insert64(h0, h1, key, lane);
          }
          for (; b1 != 0u; b1 &= b1 - 1u) {
            const uint64_t key = __shfl_sync(kFull, k1, __ffs(b1) - 1);
            if (key < __shfl_sync(kFull, h1, 31)) sorted_insert64(h0, h1, key, lane);
          }
        }
      }
    }
    const int c = seen < N ? int(seen) : N;
    if (lane == 0) out_cnt[u] = c;
    if (lane < N) {
      const bool ok = lane < c;
      out_idx[u * N + lane] = ok ? int(uint32_t(h0)) : -1;
      out_val[u * N + lane] = ok ? uint8_t(255u - unsigned(h0 >> 32)) : 0;
    }
    if (lane + 32 < N) {
      const bool ok = lane + 32 < c;
      out_idx[u * N + lane + 32] = ok ? int(uint32_t(h1)) : -1;
      out_val[u * N + lane + 32] = ok ? uint8_t(255u - unsigned(h1 >> 32)) : 0;
    }
  }
}

}  // namespace qrs

extern "C" int qrs_recommend_topn(const int64_t *row_start, const int32_t *row_cnt, const void *cols, int32_t cols_mode,
                                  const uint8_t *preds, int64_t n_users, int32_t nt, int32_t N, int32_t *out_idx,
                                  uint8_t *out_val, int32_t *out_cnt, void *stream) {
  if (n_users < 0 || nt <= 0 || N <= 0 || !row_start || !row_cnt || !out_idx || !out_val || !out_cnt)
    return QRS_ERR_INVALID_ARG;
  if (N > QRS_MAX_TOPK || (cols_mode != QRS_COLS_I32 && cols_mode != QRS_COLS_U16)) return QRS_ERR_UNSUPPORTED;
  if (n_users == 0) return QRS_OK;
  int64_t blocks = qrs::ceil_div<int64_t>(n_users, 8);
  const int64_t cap = int64_t(qrs::sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  qrs::topn_kernel<<<unsigned(blocks), 256, 0, qrs::as_stream(stream)>>>(row_start, row_cnt, cols, cols_mode == QRS_COLS_U16,
                                                                       preds, n_users, nt, N, out_idx, out_val, out_cnt);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}
