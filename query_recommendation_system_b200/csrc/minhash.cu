// a1 — MinHash signatures (recommender.py:105-143) with the band-record / norm2 epilogue fused.
//
// sig[s][i] = min_{e in set(s)} perm_i[e]  ==  the first rank r (in increasing order) whose
// element order_i[r] belongs to set(s) — the reference's own first-touch walk
// (recommender.py:125-129).  The first-touch kernel therefore keeps a membership map of the
// set in shared memory and, per hash, lets the 32 lanes of a warp test 32 consecutive ranks
// at once; the expected number of rounds is m / (32 * |set|) + 1, i.e. ~|set|^2 / m times
// fewer table reads than the per-nonzero gather-min ("direct") formulation, which remains
// for short sets and as the fallback when a hash is not resolved within the R-rank head.
#include "common.cuh"

namespace qrs {

constexpr int kByteMapMaxElems = 16384;           // one byte per element up to here, one bit beyond
constexpr int kMaxMapBytes = 200 * 1024;          // shared-memory budget for the membership map
constexpr int64_t kMaxElemsFirstTouch = int64_t(kMaxMapBytes) * 8;

__global__ void perm_head_invert_kernel(const int *__restrict__ perm_tab, int m, int K, int R,
                                        int *__restrict__ order_head) {
  const int64_t total = int64_t(m) * K;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += int64_t(gridDim.x) * blockDim.x) {
    int r = perm_tab[t];
    if (r < R) {
      int e = int(t / K), i = int(t - int64_t(e) * K);
      order_head[int64_t(i) * R + r] = e;
    }
  }
}

// min over the row of perm_tab[e*K + i] with the lanes striding over the row (rare fallback)
__device__ __forceinline__ int direct_min_one_hash(const int *__restrict__ colidx, int64_t lo, int len,
                                                   const int *__restrict__ perm_tab, int K, int i, int lane) {
  int v = 0x7FFFFFFF;
  for (int t = lane; t < len; t += 32) {
    int e = __ldg(colidx + lo + t);
    v = min(v, __ldg(perm_tab + int64_t(e) * K + i));
  }
  return __reduce_min_sync(kFull, v);
}

template <bool BYTEMAP>
__device__ __forceinline__ bool map_test(const unsigned char *map, int e) {
  if (BYTEMAP) return map[e] != 0;
  return (reinterpret_cast<const unsigned *>(map)[e >> 5] >> (e & 31)) & 1u;
}

// One CTA per set at a time (persistent grid-stride over sets).
//   dynamic smem: int sigrow[Kpad] | long long wsum[32] | membership map
template <bool BYTEMAP, int THREADS>
__global__ void __launch_bounds__(THREADS)
minhash_kernel(const int64_t *__restrict__ rowptr, const int *__restrict__ colidx, int64_t n, int m,
               const int *__restrict__ perm_tab, const int *__restrict__ order_head, int R, int K, int b,
               int64_t id_base, int *__restrict__ sig, long long *__restrict__ norm2,
               uint64_t *__restrict__ records, int64_t rec_stride, int mode, int map_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Kpad = round_up(K, 4);
  int *sigrow = reinterpret_cast<int *>(smem_raw);
  long long *wsum = reinterpret_cast<long long *>(sigrow + Kpad);
  unsigned char *map = reinterpret_cast<unsigned char *>(wsum + 32);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = THREADS / 32;
  const int r_band = b > 0 ? K / b : 0;

  if (map_bytes > 0) {
    for (int t = tid * 16; t < map_bytes; t += THREADS * 16)
      *reinterpret_cast<uint4 *>(map + t) = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();

  for (int64_t s = blockIdx.x; s < n; s += gridDim.x) {
    const int64_t lo = rowptr[s];
    const int len = int(rowptr[s + 1] - lo);
    const bool use_ft = map_bytes > 0 && R >= 32 && len > 0 &&
                        (mode == QRS_MINHASH_FIRST_TOUCH || (mode == QRS_MINHASH_AUTO && int64_t(len) * len >= m));
    if (use_ft) {
      for (int t = tid; t < len; t += THREADS) {
        int e = ld_stream_s32(colidx + lo + t);
        if (BYTEMAP) map[e] = 1;
        else atomicOr(reinterpret_cast<unsigned *>(map) + (e >> 5), 1u << (e & 31));
      }
      __syncthreads();
    }

    for (int i0 = warp * 32; i0 < K; i0 += NW * 32) {
      int myval = -1;
      if (len > 0 && use_ft) {
        const int nh = min(32, K - i0);
        for (int jj = 0; jj < nh; jj += 4) {
          int e[4];
          unsigned bal[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            int i = min(i0 + jj + u, K - 1);
            e[u] = __ldg(order_head + int64_t(i) * R + lane);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) bal[u] = __ballot_sync(kFull, map_test<BYTEMAP>(map, e[u]));
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + jj + u;
            if (i >= K) break;   // warp-uniform
            int r;
            if (bal[u]) {
              r = __ffs(bal[u]) - 1;
            } else {
              r = -1;
              for (int r0 = 32; r0 < R; r0 += 32) {
                int ee = __ldg(order_head + int64_t(i) * R + r0 + lane);
                unsigned bb = __ballot_sync(kFull, map_test<BYTEMAP>(map, ee));
                if (bb) { r = r0 + __ffs(bb) - 1; break; }
              }
              if (r < 0) r = direct_min_one_hash(colidx, lo, len, perm_tab, K, i, lane);
            }
            if (lane == jj + u) myval = r;
          }
        }
      } else if (len > 0) {
        // direct: lane owns hash i0 + lane; elements are broadcast 32 at a time
        const int i = min(i0 + lane, K - 1);
        int v = 0x7FFFFFFF;
        for (int t0 = 0; t0 < len; t0 += 32) {
          int mine = (t0 + lane < len) ? __ldg(colidx + lo + t0 + lane) : -1;
          const int cnt = min(32, len - t0);
          int t = 0;
          for (; t + 4 <= cnt; t += 4) {
            int e0 = __shfl_sync(kFull, mine, t), e1 = __shfl_sync(kFull, mine, t + 1);
            int e2 = __shfl_sync(kFull, mine, t + 2), e3 = __shfl_sync(kFull, mine, t + 3);
            int v0 = __ldg(perm_tab + int64_t(e0) * K + i), v1 = __ldg(perm_tab + int64_t(e1) * K + i);
            int v2 = __ldg(perm_tab + int64_t(e2) * K + i), v3 = __ldg(perm_tab + int64_t(e3) * K + i);
            v = min(min(v, min(v0, v1)), min(v2, v3));
          }
          for (; t < cnt; ++t) {
            int e0 = __shfl_sync(kFull, mine, t);
            v = min(v, __ldg(perm_tab + int64_t(e0) * K + i));
          }
        }
        myval = v;
      }
      if (i0 + lane < K) {
        sigrow[i0 + lane] = myval;
        sig[s * K + i0 + lane] = myval;
      }
    }
    __syncthreads();

    // ---- epilogue: squared norm (a8) and bucket records (a2/a3) from the staged row
    if (norm2 != nullptr) {
      long long acc = 0;
      for (int t = tid; t < K; t += THREADS) {
        long long x = sigrow[t];
        acc += x * x;
      }
      acc = warp_sum_ll(acc);
      if (lane == 0) wsum[warp] = acc;
    }
    if (records != nullptr) {
      for (int band = tid; band < b; band += THREADS)
        records[int64_t(band) * rec_stride + s] =
            band_record(sigrow + band * r_band, r_band, uint32_t(band), uint32_t(id_base + s));
    }
    if (use_ft) {
      for (int t = tid; t < len; t += THREADS) {
        int e = __ldg(colidx + lo + t);
        if (BYTEMAP) map[e] = 0;
        else reinterpret_cast<unsigned *>(map)[e >> 5] = 0u;
      }
    }
    __syncthreads();
    if (norm2 != nullptr && tid == 0) {
      long long tot = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) tot += wsum[w];
      norm2[s] = tot;
    }
    // the next iteration's first write to wsum/sigrow happens after its own barriers
  }
}

template <bool BYTEMAP, int THREADS>
static int launch_minhash(const int64_t *rowptr, const int *colidx, int64_t n, int m, const int *perm_tab,
                          const int *order_head, int R, int K, int b, int64_t id_base, int *sig,
                          long long *norm2, uint64_t *records, int64_t rec_stride, int mode, int map_bytes,
                          cudaStream_t stream) {
  auto kern = minhash_kernel<BYTEMAP, THREADS>;
  const size_t smem = size_t(round_up(K, 4)) * 4 + 32 * 8 + size_t(map_bytes);
  QRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  QRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
  if (per_sm < 1) return QRS_ERR_UNSUPPORTED;
  int64_t grid = int64_t(sm_count()) * per_sm;
  if (grid > n) grid = n;
  kern<<<unsigned(grid), THREADS, smem, stream>>>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base,
                                                  sig, norm2, records, rec_stride, mode, map_bytes);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // namespace qrs

using namespace qrs;

extern "C" {

int64_t qrs_minhash_max_elements(void) { return kMaxElemsFirstTouch; }

int qrs_perm_head_invert(const int32_t *perm_tab, int32_t m, int32_t K, int32_t R, int32_t *order_head,
                         void *stream) {
  if (!perm_tab || !order_head || m <= 0 || K <= 0 || R <= 0 || R > m) return QRS_ERR_INVALID_ARG;
  const int64_t total = int64_t(m) * K;
  int64_t blocks = ceil_div<int64_t>(total, 256);
  const int64_t cap = int64_t(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  perm_head_invert_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(perm_tab, m, K, R, order_head);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

int qrs_minhash_perm(const int64_t *rowptr, const int32_t *colidx, int64_t n, int32_t m,
                     const int32_t *perm_tab, const int32_t *order_head, int32_t R, int32_t K, int32_t b,
                     int64_t id_base, int32_t *sig, int64_t *norm2, uint64_t *records, int64_t rec_stride,
                     int32_t mode, void *stream) {
  if (n < 0 || m <= 0 || K <= 0 || !rowptr || !perm_tab || !sig) return QRS_ERR_INVALID_ARG;
  if (n > 0 && !colidx) return QRS_ERR_INVALID_ARG;
  if (mode < QRS_MINHASH_AUTO || mode > QRS_MINHASH_FIRST_TOUCH) return QRS_ERR_INVALID_ARG;
  if (records) {
    if (b <= 0 || b > QRS_MAX_BANDS || K % b != 0 || rec_stride < n) return QRS_ERR_INVALID_ARG;
    if (id_base < 0 || id_base + n > (int64_t(1) << 32)) return QRS_ERR_UNSUPPORTED;
  } else {
    b = 0;
  }
  if (K > 8192) return QRS_ERR_UNSUPPORTED;
  if (n == 0) return QRS_OK;
  if (!order_head || R < 32) { R = 0; order_head = nullptr; }
  if (R % 32 != 0 || R > m) return QRS_ERR_INVALID_ARG;
  bool ft_possible = R >= 32 && mode != QRS_MINHASH_DIRECT && m <= kMaxElemsFirstTouch;
  if (mode == QRS_MINHASH_FIRST_TOUCH && !ft_possible) return QRS_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  long long *n2 = reinterpret_cast<long long *>(norm2);
  if (!ft_possible)
    return launch_minhash<true, 128>(rowptr, colidx, n, m, perm_tab, order_head, 0, K, b, id_base, sig, n2,
                                     records, rec_stride, QRS_MINHASH_DIRECT, 0, st);
  if (m <= kByteMapMaxElems) {
    const int map_bytes = round_up(m, 16);
    return launch_minhash<true, 128>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, n2,
                                     records, rec_stride, mode, map_bytes, st);
  }
  const int map_bytes = round_up(ceil_div(m, 32) * 4, 16);
  if (map_bytes <= 24 * 1024)
    return launch_minhash<false, 128>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, n2,
                                      records, rec_stride, mode, map_bytes, st);
  return launch_minhash<false, 512>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, n2,
                                    records, rec_stride, mode, map_bytes, st);
}

}  // extern "C"
