// a2/a3 — banding of a signature matrix (lsh.py:17-38): exact packed bucket ids for the
// LSH object's introspection API, and bucket records for grouping when the signatures
// did not come from qrs_minhash_perm's fused epilogue.
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
  int64_t blocks = ceil_div<int64_t>(n, 8);
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  band_records_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(
      sig, n, K, b, id_base, records, rec_stride, reinterpret_cast<long long *>(norm2));
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // extern "C"
