// Multi-GPU plumbing of the path: rows are sharded across the GPUs of one box, bucket records go to
// the rank their bucket hash names, scored pairs go to the owners of their two ends, and the prediction gathers
// neighbour rows (recommender.py:320) wherever they live.
//
//   * peer buffers: a rank's prepared rating rows sit in a buffer every other rank maps (CUDA IPC) and the
//     predictor READS NEIGHBOUR ROWS IN PLACE over NVLink — no request/response exchange, no copy of remote rows;
//   * fixed-capacity exchanges: both all-to-alls move G equal slots per rank, padded with an all-ones sentinel that
//     every consumer ignores (sentinel bucket hash / row id out of range), so split sizes never travel to the host
//     and nothing synchronises with the CPU between the collectives.  A slot that would overflow raises a device
//     flag that the step reads once, at its end, together with the other counters.
#include "common.cuh"

namespace qrs {

// records are sorted by their top `lg` bits (destination rank): slot d of `out` (cap entries) gets group d
__global__ void __launch_bounds__(256)
group_pad_kernel(const uint64_t *__restrict__ sorted, int64_t n, int lg, int64_t cap, uint64_t *__restrict__ out,
                 unsigned long long *__restrict__ counts, unsigned *__restrict__ overflow) {
  __shared__ int64_t bound[65];
  const int G = 1 << lg;
  if (threadIdx.x <= G) {       // first record whose destination is >= threadIdx.x
    const uint64_t d = threadIdx.x;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((sorted[mid] >> (64 - lg)) < d) lo = mid + 1; else hi = mid;
    }
    bound[threadIdx.x] = threadIdx.x == G ? n : lo;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < G) {
    const int64_t c = bound[threadIdx.x + 1] - bound[threadIdx.x];
    counts[threadIdx.x] = (unsigned long long)c;
    if (c > cap) atomicOr(overflow, 1u);
  }
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const uint64_t rec = ld_stream_u64(sorted + i);
    const int d = int(rec >> (64 - lg));
    const int64_t pos = i - bound[d];
    if (pos < cap) out[int64_t(d) * cap + pos] = rec;
  }
}

// scored pair -> one 64-bit word (i << 37 | j << 11 | sim + 1000; ids < 2^26) appended to the slot of the owner
// of i and, if different, of j.  Appends are warp-aggregated per destination.
__global__ void __launch_bounds__(256)
route_pairs_kernel(const uint64_t *__restrict__ pairs, const int *__restrict__ sims,
                   const unsigned long long *__restrict__ n_pairs_dev, int64_t capacity_in, int64_t rows_per_rank,
                   int G, int64_t cap, uint64_t *__restrict__ out, unsigned long long *__restrict__ counts,
                   unsigned *__restrict__ overflow) {
  int64_t n = int64_t(*n_pairs_dev);
  if (n > capacity_in) n = capacity_in;
  const int lane = threadIdx.x & 31;
  const int64_t total = round_up<int64_t>(n, 32);
  for (int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += int64_t(gridDim.x) * blockDim.x) {
    int d0 = -1, d1 = -1;
    uint64_t word = 0;
    if (p < n) {
      const uint64_t pr = ld_stream_u64(pairs + p);
      const uint64_t i = pr >> 32, j = pr & 0xFFFFFFFFu;
      d0 = int(i / uint64_t(rows_per_rank));
      d1 = int(j / uint64_t(rows_per_rank));
      if (d1 == d0) d1 = -1;
      word = (i << 37) | (j << 11) | uint64_t(sims[p] + 1000);
    }
#pragma unroll
    for (int leg = 0; leg < 2; ++leg) {
      const int d = leg == 0 ? d0 : d1;
      const unsigned peers = __match_any_sync(kFull, d);
      if (d >= 0) {
        const int leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(counts + d, (unsigned long long)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        const int64_t pos = int64_t(base) + __popc(peers & lanemask_lt());
        if (pos < cap) out[int64_t(d) * cap + pos] = word;
        else atomicOr(overflow, 2u);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
unpack_routed_kernel(const uint64_t *__restrict__ words, int64_t n, uint64_t *__restrict__ pairs,
                     int *__restrict__ sims) {
  for (int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < n; p += int64_t(gridDim.x) * blockDim.x) {
    const uint64_t w = ld_stream_u64(words + p);
    if (w == ~uint64_t(0)) {                 // empty slot entry: a pair no rank owns
      pairs[p] = ~uint64_t(0);
      sims[p] = 0;
    } else {
      pairs[p] = ((w >> 37) << 32) | ((w >> 11) & 0x3FFFFFFu);
      sims[p] = int(w & 0x7FFu) - 1000;
    }
  }
}

}  // namespace qrs

using namespace qrs;

extern "C" {

int qrs_group_pad(const uint64_t *sorted, int64_t n, int32_t lg, int64_t cap, uint64_t *out,
                  unsigned long long *counts, uint32_t *overflow, void *stream) {
  if (n < 0 || lg < 1 || lg > 6 || cap < 0 || !out || !counts || !overflow || (n > 0 && !sorted))
    return QRS_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  QRS_CUDA(cudaMemsetAsync(out, 0xFF, size_t(cap) * size_t(1 << lg) * 8, st));
  int64_t blocks = ceil_div<int64_t>(n > 0 ? n : 1, 256 * 4);
  const int64_t capb = int64_t(sm_count()) * 8;
  if (blocks > capb) blocks = capb;
  group_pad_kernel<<<unsigned(blocks), 256, 0, st>>>(sorted, n, lg, cap, out, counts, overflow);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

int qrs_route_pairs(const uint64_t *pairs, const int32_t *sims, const unsigned long long *n_pairs_dev,
                    int64_t capacity_in, int64_t rows_per_rank, int32_t n_ranks, int64_t cap, uint64_t *out,
                    unsigned long long *counts, uint32_t *overflow, void *stream) {
  if (!pairs || !sims || !n_pairs_dev || capacity_in < 0 || rows_per_rank <= 0 || n_ranks < 1 || n_ranks > 64 ||
      cap < 0 || !out || !counts || !overflow)
    return QRS_ERR_INVALID_ARG;
  if (rows_per_rank * n_ranks >= (int64_t(1) << 26)) return QRS_ERR_UNSUPPORTED;   // ids are packed into 26 bits
  cudaStream_t st = as_stream(stream);
  QRS_CUDA(cudaMemsetAsync(out, 0xFF, size_t(cap) * size_t(n_ranks) * 8, st));
  QRS_CUDA(cudaMemsetAsync(counts, 0, size_t(n_ranks) * 8, st));
  if (capacity_in == 0) return QRS_OK;
  int64_t blocks = ceil_div<int64_t>(capacity_in, 256 * 4);
  const int64_t capb = int64_t(sm_count()) * 8;
  if (blocks > capb) blocks = capb;
  route_pairs_kernel<<<unsigned(blocks), 256, 0, st>>>(pairs, sims, n_pairs_dev, capacity_in, rows_per_rank, n_ranks,
                                                        cap, out, counts, overflow);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

int qrs_unpack_routed(const uint64_t *words, int64_t n, uint64_t *pairs, int32_t *sims, void *stream) {
  if (n < 0 || (n > 0 && (!words || !pairs || !sims))) return QRS_ERR_INVALID_ARG;
  if (n == 0) return QRS_OK;
  int64_t blocks = ceil_div<int64_t>(n, 256 * 4);
  const int64_t capb = int64_t(sm_count()) * 8;
  if (blocks > capb) blocks = capb;
  unpack_routed_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(words, n, pairs, sims);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

/* ------------------------------------------------------------------ peer buffers (CUDA IPC, one process per GPU) */
int qrs_peer_alloc(size_t bytes, void **ptr) {
  if (!ptr || bytes == 0) return QRS_ERR_INVALID_ARG;
  QRS_CUDA(cudaMalloc(ptr, bytes));
  return QRS_OK;
}

int qrs_peer_free(void *ptr) {
  if (!ptr) return QRS_OK;
  QRS_CUDA(cudaFree(ptr));
  return QRS_OK;
}

int qrs_peer_export(void *ptr, unsigned char *handle64) {
  if (!ptr || !handle64) return QRS_ERR_INVALID_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI carries the handle as 64 bytes");
  cudaIpcMemHandle_t h;
  QRS_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return QRS_OK;
}

int qrs_peer_open(const unsigned char *handle64, void **ptr) {
  if (!handle64 || !ptr) return QRS_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  QRS_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return QRS_OK;
}

int qrs_peer_close(void *ptr) {
  if (!ptr) return QRS_OK;
  QRS_CUDA(cudaIpcCloseMemHandle(ptr));
  return QRS_OK;
}

}  // extern "C"
