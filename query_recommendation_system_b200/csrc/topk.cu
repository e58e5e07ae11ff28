// a7 + a9 — neighbour tables (recommender.py:183-210): adjacency lists built from the scored
// candidate pairs by counting + scan + fill, then one warp per row keeps the k best
// (similarity descending, neighbour id ascending) in registers with a 64-wide bitonic
// network — two keys per lane, never touching shared or local memory.
#include "common.cuh"

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

// compare-exchange step of the 64-key network (key e = lane + 32*slot)
__device__ __forceinline__ void cx_shfl(uint64_t &x, int j, bool ascending, int lane) {
  const uint64_t other = __shfl_xor_sync(kFull, x, j);
  const bool lower = (lane & j) == 0;
  const uint64_t mn = x < other ? x : other, mx = x < other ? other : x;
  x = (lower == ascending) ? mn : mx;
}
__device__ __forceinline__ void cx_local(uint64_t &a, uint64_t &b, bool ascending) {
  const uint64_t mn = a < b ? a : b, mx = a < b ? b : a;
  a = ascending ? mn : mx;
  b = ascending ? mx : mn;
}
__device__ __forceinline__ void bitonic_sort64(uint64_t &x0, uint64_t &x1, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      // slot 0 holds e = lane, slot 1 holds e = lane + 32; (e & k) is the same for both when k <= 16,
      // and for k == 32 slot 1 runs descending
      const bool asc0 = (lane & k) == 0;
      const bool asc1 = k == 32 ? false : asc0;
      cx_shfl(x0, j, k == 32 ? true : asc0, lane);
      cx_shfl(x1, j, asc1, lane);
    }
  }
  // k = 64: everything ascending; j = 32 is the in-thread exchange
  cx_local(x0, x1, true);
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    cx_shfl(x0, j, true, lane);
    cx_shfl(x1, j, true, lane);
  }
}
// sorts a bitonic 64-sequence ascending
__device__ __forceinline__ void bitonic_merge64(uint64_t &x0, uint64_t &x1, int lane) {
  cx_local(x0, x1, true);
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    cx_shfl(x0, j, true, lane);
    cx_shfl(x1, j, true, lane);
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
      bitonic_sort64(n0, n1, lane);
      if (base == lo) {
        h0 = n0; h1 = n1;
      } else {
        // lowest 64 of two ascending lists: min(h[e], n[63 - e]) is bitonic
        const uint64_t r1 = __shfl_sync(kFull, n1, 31 - lane);   // n[63 - lane]
        const uint64_t r0 = __shfl_sync(kFull, n0, 31 - lane);   // n[31 - lane] = n[63 - (lane + 32)]
        h0 = h0 < r1 ? h0 : r1;
        h1 = h1 < r0 ? h1 : r0;
        bitonic_merge64(h0, h1, lane);
      }
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
