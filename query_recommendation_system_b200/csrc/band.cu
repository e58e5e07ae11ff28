// a2/a3 — banding of a signature matrix (lsh.py:17-38): exact packed bucket ids for the
// LSH object's introspection API, and bucket records for grouping when the signatures
// did not come from qrs_minhash_perm's fused epilogue.
#include <stdlib.h>

#include "common.cuh"

namespace qrs {

// thread per (set, band)
__global__ void __launch_bounds__(256)
band_keys_kernel(const int *__restrict__ sig, int64_t n, int K, int b, uint64_t *__restrict__ keys,
                 uint64_t *__restrict__ keys_hi, uint8_t *__restrict__ skipped) {
  const int r = K / b;
  const int64_t total = n * b;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t s = t / b;
    const int band = int(t - s * b);
    const int *lanes = sig + s * K + band * r;
    uint64_t lo = 0, hi = 0;
    bool all_neg = true;
    for (int c = 0; c < r; ++c) {
      uint64_t l16 = uint32_t(lanes[c]) & 0xFFFFu;
      all_neg = all_neg && (l16 == 0xFFFFu);
      hi = (hi << 16) | (lo >> 48);
      lo = (lo << 16) | l16;
    }
    keys[t] = lo;
    if (keys_hi) keys_hi[t] = hi;
    if (skipped) skipped[t] = all_neg ? 1 : 0;
  }
}

// warp per set: stage nothing, lanes stride over bands; norm2 by a warp reduction
__global__ void __launch_bounds__(256)
band_records_kernel(const int *__restrict__ sig, int64_t n, int K, int b, int64_t id_base,
                    uint64_t *__restrict__ records, int64_t rec_stride, long long *__restrict__ norm2) {
  const int r = K / b;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t s = warp_global; s < n; s += nwarps) {
    const int *row = sig + s * K;
    if (records) {
      for (int band = lane; band < b; band += 32)
        records[int64_t(band) * rec_stride + s] = band_record(row + band * r, r, uint32_t(band), uint32_t(id_base + s));
    }
    if (norm2) {
      long long acc = 0;
      for (int t = lane; t < K; t += 32) {
        long long x = row[t];
        acc += x * x;
      }
      acc = warp_sum_ll(acc);
      if (lane == 0) norm2[s] = acc;
    }
  }
}


// ------------------------------------------------------------------ tile form: TMA-staged signature tiles
// kBandTileSets consecutive signature rows are ONE contiguous block of sig (kBandTileSets * K ints): a single
// thread moves the block into shared memory with a bulk asynchronous copy (cp.async.bulk, the TMA unit's 1-D
// form) that completes on an mbarrier, kBandStages tiles deep, so the copies of the next tiles run while this
// one is hashed and no signature value passes through a register on its way in.  Hashing: thread <-> (set, band)
// in row-major order, i.e. consecutive threads read consecutive r-value groups of the staged block (one 16-byte
// shared load for r = 4); the records go through a small transposed staging array (pitch kBandTileSets + 1
// records: the 8-byte stores of a half-warp fall into 16 different bank pairs) and leave as runs of
// kBandTileSets consecutive records per band — 512-byte coalesced stores instead of one 8-byte store per band
// array.  The squared norms are taken from the same staged rows, one warp per row.
constexpr int kBandStages = 3;
constexpr int kBandThreads = 256;
constexpr size_t kBandTileMaxSmem = 200 * 1024;

__host__ __device__ constexpr size_t band_tile_smem_bytes(int K, int b, int sets) {
  return size_t(kBandStages) * sets * K * 4 + size_t(b) * (sets + 1) * 8 + 64;
}

template <int R>
__device__ __forceinline__ uint64_t band_record_staged(const int *lanes, int r, uint32_t band, uint32_t id) {
  if constexpr (R == 4) {
    const int4 v = *reinterpret_cast<const int4 *>(lanes);
    const uint32_t w0 = (uint32_t(v.x) & 0xFFFFu) | (uint32_t(v.y) << 16), w1 = (uint32_t(v.z) & 0xFFFFu) | (uint32_t(v.w) << 16);
    uint32_t h = band_hash_word(band_hash_word(band_hash_init(band), w0), w1);
    return make_record(band_hash_finish(h, (w0 & w1) == 0xFFFFFFFFu), band, id);
  } else if constexpr (R == 2) {
    const int2 v = *reinterpret_cast<const int2 *>(lanes);
    const uint32_t w0 = (uint32_t(v.x) & 0xFFFFu) | (uint32_t(v.y) << 16);
    return make_record(band_hash_finish(band_hash_word(band_hash_init(band), w0), w0 == 0xFFFFFFFFu), band, id);
  } else {
    return band_record(lanes, r, band, id);
  }
}

template <int R, int kBandTileSets>
__global__ void __launch_bounds__(kBandThreads)
band_records_tile_kernel(const int *__restrict__ sig, int64_t n, int K, int b, int64_t id_base,
                         uint64_t *__restrict__ records, int64_t rec_stride, long long *__restrict__ norm2) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int r = K / b;
  const size_t stage_bytes = size_t(kBandTileSets) * K * 4;
  uint64_t *outrec = reinterpret_cast<uint64_t *>(smem_raw + kBandStages * stage_bytes);
  const unsigned bars = unsigned(__cvta_generic_to_shared(outrec + size_t(b) * (kBandTileSets + 1)));
  const unsigned stage_s = unsigned(__cvta_generic_to_shared(smem_raw));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tiles = ceil_div<int64_t>(n, kBandTileSets);
  const int64_t my_tiles = tiles > blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  auto issue = [&](int64_t it) {          // one thread: tile `it` of this CTA into stage it % kBandStages
    const int64_t s0 = (int64_t(blockIdx.x) + it * gridDim.x) * kBandTileSets;
    const unsigned bytes = unsigned(min(int64_t(kBandTileSets), n - s0) * K * 4);
    const int st = int(it % kBandStages);
    mbar_arrive_expect_tx(bars + 8u * st, bytes);
    bulk_copy_g2s(stage_s + unsigned(st * stage_bytes), sig + s0 * K, bytes, bars + 8u * st);
  };
  if (tid == 0) {
    for (int st = 0; st < kBandStages; ++st) mbar_init(bars + 8u * st, 1u);
    mbar_init_fence();
    for (int64_t it = 0; it < my_tiles && it < kBandStages; ++it) issue(it);
  }
  __syncthreads();

  for (int64_t it = 0; it < my_tiles; ++it) {
    const int st = int(it % kBandStages);
    const int64_t s0 = (int64_t(blockIdx.x) + it * gridDim.x) * kBandTileSets;
    const int nsets = int(min(int64_t(kBandTileSets), n - s0));
    const int *tile = reinterpret_cast<const int *>(smem_raw + st * stage_bytes);
    mbar_wait(bars + 8u * st, unsigned(it / kBandStages) & 1u);
    if (records != nullptr) {
      const int total = nsets * b;
      for (int x = tid; x < total; x += kBandThreads) {
        const int set = x / b, band = x - set * b;
        outrec[band * (kBandTileSets + 1) + set] =
            band_record_staged<R>(tile + set * K + band * r, r, uint32_t(band), uint32_t(id_base + s0 + set));
      }
    }
    if (norm2 != nullptr) {
      for (int set = warp; set < nsets; set += kBandThreads / 32) {
        long long acc = 0;
        for (int t = lane; t < K; t += 32) {
          const long long x = tile[set * K + t];
          acc += x * x;
        }
        acc = warp_sum_ll(acc);
        if (lane == 0) norm2[s0 + set] = acc;
      }
    }
    __syncthreads();          // the stage has been read by everybody, the records are staged
    if (tid == 0 && it + kBandStages < my_tiles) {
      fence_proxy_async_smem();
      issue(it + kBandStages);
    }
    if (records != nullptr) {
      for (int band = warp; band < b; band += kBandThreads / 32) {
        uint64_t *dst = records + int64_t(band) * rec_stride + s0;
        const uint64_t *src = outrec + band * (kBandTileSets + 1);
        for (int j = lane; j < nsets; j += 32) dst[j] = src[j];
      }
      __syncthreads();        // outrec is rewritten by the next tile
    }
  }
}

template <int R, int kBandTileSets>
static int launch_band_records_tile(const int *sig, int64_t n, int K, int b, int64_t id_base, uint64_t *records,
                                    int64_t rec_stride, long long *norm2, cudaStream_t st) {
  auto kern = band_records_tile_kernel<R, kBandTileSets>;
  const size_t smem = band_tile_smem_bytes(K, b, kBandTileSets);
  QRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  QRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBandThreads, smem));
  if (per_sm < 1) return QRS_ERR_UNSUPPORTED;
  int64_t grid = int64_t(sm_count()) * per_sm;
  const int64_t tiles = ceil_div<int64_t>(n, kBandTileSets);
  if (grid > tiles) grid = tiles;
  kern<<<unsigned(grid), kBandThreads, smem, st>>>(sig, n, K, b, id_base, records, rec_stride, norm2);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

// sets per tile: the largest of 64 / 32 / 16 whose stages leave room for two CTAs on an SM (measured: profiles/r2_band_probe.jsonl; the hashing of one
// CTA then overlaps the stores of another); 0 = no tile fits
static int band_tile_sets(int K, int b) {
  static const int forced = [] { const char *e = getenv("QRS_BAND_TILE_SETS"); return e ? atoi(e) : 0; }();
  for (int sets = 64; sets >= 16; sets >>= 1) {
    if (forced && forced != sets) continue;
    const size_t bytes = band_tile_smem_bytes(K, b, sets);
    if (bytes <= (forced ? kBandTileMaxSmem : size_t(113 * 1024))) return sets;
  }
  return band_tile_smem_bytes(K, b, 16) <= kBandTileMaxSmem ? 16 : 0;
}

template <int R>
static int launch_band_records_tile_r(int sets, const int *sig, int64_t n, int K, int b, int64_t id_base,
                                      uint64_t *records, int64_t rec_stride, long long *norm2, cudaStream_t st) {
  if (sets == 64) return launch_band_records_tile<R, 64>(sig, n, K, b, id_base, records, rec_stride, norm2, st);
  if (sets == 32) return launch_band_records_tile<R, 32>(sig, n, K, b, id_base, records, rec_stride, norm2, st);
  return launch_band_records_tile<R, 16>(sig, n, K, b, id_base, records, rec_stride, norm2, st);
}

}  // namespace qrs

using namespace qrs;

extern "C" {

int qrs_band_keys(const int32_t *sig, int64_t n, int32_t K, int32_t b, uint64_t *keys, uint64_t *keys_hi,
                  uint8_t *skipped, void *stream) {
  if (!sig || !keys || n < 0 || K <= 0 || b <= 0 || K % b != 0) return QRS_ERR_INVALID_ARG;
  const int r = K / b;
  if (r > 8 || (r > 4 && !keys_hi)) return QRS_ERR_UNSUPPORTED;
  if (n == 0) return QRS_OK;
  int64_t blocks = ceil_div<int64_t>(n * b, 256);
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  band_keys_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(sig, n, K, b, keys, keys_hi, skipped);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

int qrs_band_records(const int32_t *sig, int64_t n, int32_t K, int32_t b, int64_t id_base, uint64_t *records,
                     int64_t rec_stride, int64_t *norm2, void *stream) {
  if (!sig || n < 0 || K <= 0) return QRS_ERR_INVALID_ARG;
  if (records) {
    if (b <= 0 || b > QRS_MAX_BANDS || K % b != 0 || rec_stride < n) return QRS_ERR_INVALID_ARG;
    if (id_base < 0 || id_base + n > (int64_t(1) << 32)) return QRS_ERR_UNSUPPORTED;
  } else {
    b = 1;
  }
  if (n == 0) return QRS_OK;
  // tile form (bulk-copy staged signature tiles) whenever a tile is a 16-byte multiple of a 16-byte aligned matrix
  // and the stages fit shared memory; QRS_BAND_PLAIN=1 keeps the per-row kernel (measurements)
  static const bool plain = [] { const char *e = getenv("QRS_BAND_PLAIN"); return e && atoi(e) != 0; }();
  const int tile_sets = (!plain && K % 4 == 0 && (reinterpret_cast<uintptr_t>(sig) & 15) == 0) ? band_tile_sets(K, b) : 0;
  if (tile_sets != 0) {
    const int r = K / b;
    long long *n2 = reinterpret_cast<long long *>(norm2);
    cudaStream_t st = as_stream(stream);
    if (records && r == 4) return launch_band_records_tile_r<4>(tile_sets, sig, n, K, b, id_base, records, rec_stride, n2, st);
    if (records && r == 2) return launch_band_records_tile_r<2>(tile_sets, sig, n, K, b, id_base, records, rec_stride, n2, st);
    return launch_band_records_tile_r<0>(tile_sets, sig, n, K, b, id_base, records, rec_stride, n2, st);
  }
  int64_t blocks = ceil_div<int64_t>(n, 8);
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  band_records_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(
      sig, n, K, b, id_base, records, rec_stride, reinterpret_cast<long long *>(norm2));
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // extern "C"
