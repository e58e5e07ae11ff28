// a3/a4 — bucket grouping (lsh.py:35-37 inserts every (band key → id) into a dict; here the
// records are grouped by a stable LSD radix sort, 8 bits per pass, over `segs` independent
// segments).  Per pass: per-tile digit histograms → one device-wide exclusive scan over the
// (segment, digit, tile) counts, which directly yields global destinations because every
// segment holds exactly n_per_seg keys → stable scatter with warp-match ranking.
#include "common.cuh"

namespace qrs {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;                            // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;      // 4096 keys per CTA
constexpr int kRadix = 256;

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint64_t *__restrict__ keys, int64_t n_per_seg, int tiles, int shift, uint32_t mask,
                  uint32_t *__restrict__ counts) {
  __shared__ uint32_t hist[kRadix];
  const int seg = blockIdx.y, tile = blockIdx.x;
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t *src = keys + int64_t(seg) * n_per_seg;
  const int64_t base = int64_t(tile) * kSortTile;
#pragma unroll 4
  for (int it = 0; it < kSortItems; ++it) {
    int64_t p = base + it * kSortThreads + threadIdx.x;
    if (p < n_per_seg) atomicAdd(&hist[uint32_t(src[p] >> shift) & mask], 1u);
  }
  __syncthreads();
  counts[(int64_t(seg) * kRadix + threadIdx.x) * tiles + tile] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t *__restrict__ keys, uint64_t *__restrict__ out, int64_t n_per_seg, int tiles,
                     int shift, uint32_t mask, const uint64_t *__restrict__ offsets) {
  __shared__ uint32_t whist[kSortWarps][kRadix];
  __shared__ uint64_t gbase[kRadix];
  const int seg = blockIdx.y, tile = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = threadIdx.x; t < kSortWarps * kRadix; t += kSortThreads) (&whist[0][0])[t] = 0;
  gbase[threadIdx.x] = offsets[(int64_t(seg) * kRadix + threadIdx.x) * tiles + tile];
  __syncthreads();

  const uint64_t *src = keys + int64_t(seg) * n_per_seg;
  const int64_t base = int64_t(tile) * kSortTile + int64_t(warp) * (32 * kSortItems);
  uint64_t key[kSortItems];
  uint16_t rank[kSortItems];
  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const int64_t p = base + it * 32 + lane;
    const bool valid = p < n_per_seg;
    key[it] = valid ? src[p] : 0;
    const uint32_t d = valid ? (uint32_t(key[it] >> shift) & mask) : kRadix;   // invalid lanes match only each other
    const unsigned peers = __match_any_sync(kFull, d);
    uint32_t prev = 0;
    if (valid) prev = whist[warp][d];
    __syncwarp();
    if (valid && (peers & lt) == 0) whist[warp][d] = prev + __popc(peers);
    __syncwarp();
    rank[it] = uint16_t(prev + __popc(peers & lt));
  }
  __syncthreads();
  {
    // exclusive scan of each digit's count across the warps of this tile
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      uint32_t c = whist[w][d];
      whist[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const int64_t p = base + it * 32 + lane;
    if (p < n_per_seg) {
      const uint32_t d = uint32_t(key[it] >> shift) & mask;
      out[gbase[d] + whist[warp][d] + rank[it]] = key[it];
    }
  }
}

static int64_t sort_tiles(int64_t n_per_seg) { return ceil_div<int64_t>(n_per_seg > 0 ? n_per_seg : 1, kSortTile); }

}  // namespace qrs

using namespace qrs;

extern "C" {

size_t qrs_sort_workspace_bytes(int64_t n_per_seg, int32_t segs) {
  if (n_per_seg < 0 || segs <= 0) return 0;
  const int64_t entries = int64_t(segs) * kRadix * sort_tiles(n_per_seg);
  return round_up<size_t>(size_t(entries) * 4, 256) + round_up<size_t>(size_t(entries + 1) * 8, 256) +
         scan_workspace_bytes(entries);
}

int qrs_sort_u64(uint64_t *keys, uint64_t *tmp, int64_t n_per_seg, int32_t segs, int32_t bit_lo, int32_t bit_hi,
                 void *ws, size_t ws_bytes, void *stream) {
  if (n_per_seg < 0 || segs <= 0 || bit_lo < 0 || bit_hi > 64 || bit_lo > bit_hi) return QRS_ERR_INVALID_ARG;
  if (segs > 65535) return QRS_ERR_UNSUPPORTED;
  if (n_per_seg == 0 || bit_lo == bit_hi) return QRS_OK;       // nothing to sort: empty buffers may be null
  if (!keys || !tmp) return QRS_ERR_INVALID_ARG;
  if (!ws || ws_bytes < qrs_sort_workspace_bytes(n_per_seg, segs)) return QRS_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int64_t tiles = sort_tiles(n_per_seg);
  if (tiles > 0x7FFFFFFF) return QRS_ERR_UNSUPPORTED;
  const int64_t entries = int64_t(segs) * kRadix * tiles;
  char *w = static_cast<char *>(ws);
  uint32_t *counts = reinterpret_cast<uint32_t *>(w);
  w += round_up<size_t>(size_t(entries) * 4, 256);
  uint64_t *offsets = reinterpret_cast<uint64_t *>(w);
  w += round_up<size_t>(size_t(entries + 1) * 8, 256);
  void *scan_ws = w;
  const size_t scan_ws_bytes = scan_workspace_bytes(entries);

  uint64_t *src = keys, *dst = tmp;
  const dim3 grid((unsigned)tiles, (unsigned)segs, 1u);
  for (int shift = bit_lo; shift < bit_hi; shift += 8) {
    const int nb = bit_hi - shift < 8 ? bit_hi - shift : 8;
    const uint32_t mask = (1u << nb) - 1u;
    radix_hist_kernel<<<grid, kSortThreads, 0, st>>>(src, n_per_seg, int(tiles), shift, mask, counts);
    QRS_LAUNCH_CHECK();
    int rc = exclusive_scan_u32_to_u64(counts, offsets, entries, scan_ws, scan_ws_bytes, st);
    if (rc != QRS_OK) return rc;
    radix_scatter_kernel<<<grid, kSortThreads, 0, st>>>(src, dst, n_per_seg, int(tiles), shift, mask, offsets);
    QRS_LAUNCH_CHECK();
    uint64_t *t = src; src = dst; dst = t;
  }
  if (src != keys)
    QRS_CUDA(cudaMemcpyAsync(keys, src, size_t(n_per_seg) * segs * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  return QRS_OK;
}

}  // extern "C"
