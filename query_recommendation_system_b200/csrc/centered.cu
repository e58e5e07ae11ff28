// f-3: the user-similarity branch at scale.  The reference clusters users with
// PCA + BIRCH and scores every pair inside a cluster with the cosine of the mean-centred rating
// rows (recommender.py:256-269).  Here the candidate pairs come from the LSH stages (Binding B) and
// only the scoring is kept: the same centred rows, the same rounded, clamped cosine.
//
// The reference's centred block is an INTEGER array (recommender.py:263): `row[rated] - mean`
// assigned back into it truncates toward zero, so a centred value is the C quotient
// (r * n - S) / n for a row with n rated cells summing to S.  Dot products and squared norms are
// therefore exact integers; one fp64 sqrt and divide follow, as in qrs_pair_cosine.
#include "common.cuh"

namespace qrs {

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ int centered_milli(long long dot, long long n2a, long long n2b) {
  const double den = sqrt(double(n2a) * double(n2b));
  if (den == 0.0) return 0;                        // a zero row stays zero under sklearn's normalize
  const int milli = int(rint(double(dot) / den * 1000.0));
  return milli < 0 ? 0 : milli;                    // recommender.py:269
}

// warp per row: centred values of the rated cells and the squared norm of the centred row
__global__ void __launch_bounds__(256)
center_rows_kernel(const int64_t *__restrict__ rowptr, const uint8_t *__restrict__ val, int64_t n_rows,
                   int16_t *__restrict__ cval, long long *__restrict__ norm2) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = warp_global; row < n_rows; row += nwarps) {
    const int64_t a = rowptr[row], e = rowptr[row + 1];
    int cnt = 0;
    long long sum = 0;
    for (int64_t x = a + lane; x < e; x += 32) {
      const int r = val[x];
      cnt += r != 0;
      sum += r;
    }
    cnt = warp_sum_int(cnt);
    sum = warp_sum_ll(sum);
    long long sq = 0;
    const bool narrow = e - a < (int64_t(1) << 23);     // |r * n - S| < 2^31: 32-bit division
    for (int64_t x = a + lane; x < e; x += 32) {
      const int r = val[x];
      int c = 0;
      if (r != 0) c = narrow ? (r * cnt - int(sum)) / cnt : int(((long long)r * cnt - sum) / cnt);
      cval[x] = int16_t(c);
      sq += c * c;
    }
    sq = warp_sum_ll(sq);
    if (lane == 0) norm2[row] = sq;
  }
}

constexpr int kCcThreads = 256;
constexpr int kCcWarps = kCcThreads / 32;
constexpr int kCcChunk = 256;          // consecutive pairs one CTA takes at a time
constexpr int kCcStagedMaxM = 200 * 1024;

// Pairs grouped by their first id.  A CTA walks a chunk of pairs; for each run of one first id i it
// scatters row i into shared memory (the rating byte by column, plus the row's rating -> centred
// value table), and every warp then streams the second row of one pair against it.
__global__ void __launch_bounds__(kCcThreads)
pair_centered_staged_kernel(const int64_t *__restrict__ rowptr, const int *__restrict__ colidx,
                            const uint8_t *__restrict__ val, const int16_t *__restrict__ cval,
                            const long long *__restrict__ norm2, const uint64_t *__restrict__ pairs, int64_t n_pairs,
                            int m_pad, int *__restrict__ sim_milli) {
  extern __shared__ __align__(16) unsigned char s_row[];
  short *lut = reinterpret_cast<short *>(s_row + m_pad);          // [256] rating -> centred value of row i
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int x = tid; x < m_pad / 4; x += kCcThreads) reinterpret_cast<unsigned *>(s_row)[x] = 0u;
  if (tid < 128) reinterpret_cast<unsigned *>(lut)[tid] = 0u;
  __syncthreads();
  const int64_t chunks = ceil_div<int64_t>(n_pairs, kCcChunk);
  for (int64_t ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
    int64_t p = ch * kCcChunk;
    const int64_t end = min(n_pairs, p + kCcChunk);
    while (p < end) {
      const unsigned i = unsigned(pairs[p] >> 32);
      int64_t lo = p + 1, hi = end;                               // end of the run of first id i (block-uniform)
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (unsigned(pairs[mid] >> 32) == i) lo = mid + 1; else hi = mid;
      }
      const int64_t run_end = lo;
      const int64_t a = rowptr[i], e = rowptr[i + 1];
      for (int64_t x = a + tid; x < e; x += kCcThreads) {
        const int r = val[x];
        s_row[colidx[x]] = (unsigned char)r;
        lut[r] = cval[x];                                         // same value from every cell rated r
      }
      __syncthreads();
      const long long n2i = norm2[i];
      for (int64_t q = p + warp; q < run_end; q += kCcWarps) {
        const unsigned j = unsigned(pairs[q]);
        const int64_t b0 = rowptr[j], b1 = rowptr[j + 1];
        long long dot = 0;
        int64_t x = b0 + lane;
        for (; x + 96 < b1; x += 128) {                           // four independent gathers in flight
          int col[4], cj[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { col[u] = ld_stream_s32(colidx + x + 32 * u); cj[u] = cval[x + 32 * u]; }
          int part = 0;
#pragma unroll
          for (int u = 0; u < 4; ++u) part += int(lut[s_row[col[u]]]) * cj[u];
          dot += part;
        }
        for (; x < b1; x += 32) dot += int(lut[s_row[ld_stream_s32(colidx + x)]]) * int(cval[x]);
        dot = warp_sum_ll(dot);
        if (lane == 0) sim_milli[q] = centered_milli(dot, n2i, norm2[j]);
      }
      __syncthreads();
      for (int64_t x = a + tid; x < e; x += kCcThreads) s_row[colidx[x]] = 0;
      __syncthreads();
      p = run_end;
    }
  }
}

// Any pair order, any m: warp per pair, the shorter row's cells are looked up in the longer row.
__global__ void __launch_bounds__(256)
pair_centered_search_kernel(const int64_t *__restrict__ rowptr, const int *__restrict__ colidx,
                            const int16_t *__restrict__ cval, const long long *__restrict__ norm2,
                            const uint64_t *__restrict__ pairs, int64_t n_pairs, int *__restrict__ sim_milli) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t q = warp_global; q < n_pairs; q += nwarps) {
    const uint64_t pr = pairs[q];
    const unsigned i = unsigned(pr >> 32), j = unsigned(pr);
    int64_t a0 = rowptr[i], a1 = rowptr[i + 1], b0 = rowptr[j], b1 = rowptr[j + 1];
    if (a1 - a0 > b1 - b0) { int64_t t = a0; a0 = b0; b0 = t; t = a1; a1 = b1; b1 = t; }
    long long dot = 0;
    for (int64_t x = a0 + lane; x < a1; x += 32) {
      const int ca = cval[x];
      if (ca == 0) continue;
      const int col = colidx[x];
      int64_t lo = b0, hi = b1;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(colidx + mid) < col) lo = mid + 1; else hi = mid;
      }
      if (lo < b1 && __ldg(colidx + lo) == col) dot += ca * int(cval[lo]);
    }
    dot = warp_sum_ll(dot);
    if (lane == 0) sim_milli[q] = centered_milli(dot, norm2[i], norm2[j]);
  }
}

}  // namespace qrs

using namespace qrs;

extern "C" int qrs_center_rows(const int64_t *rowptr, const uint8_t *val, int64_t n_rows, int16_t *cval,
                               int64_t *norm2, void *stream) {
  if (n_rows < 0 || !rowptr || !norm2) return QRS_ERR_INVALID_ARG;
  if (n_rows == 0) return QRS_OK;
  int64_t blocks = ceil_div<int64_t>(n_rows, 8);
  const int64_t cap = int64_t(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  center_rows_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(rowptr, val, n_rows, cval,
                                                                     reinterpret_cast<long long *>(norm2));
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

extern "C" int qrs_pair_centered_cosine(const int64_t *rowptr, const int32_t *colidx, const uint8_t *val,
                                        const int16_t *cval, const int64_t *norm2, const uint64_t *pairs,
                                        int64_t n_pairs, int32_t m, int32_t flags, int32_t *sim_milli,
                                        void *stream) {
  if (n_pairs < 0 || m <= 0 || !rowptr || !norm2) return QRS_ERR_INVALID_ARG;
  if (n_pairs == 0) return QRS_OK;
  if (!colidx || !val || !cval || !pairs || !sim_milli) return QRS_ERR_INVALID_ARG;
  const bool grouped = (flags & QRS_PAIRS_GROUPED) != 0, force_search = (flags & QRS_PAIRS_FORCE_SEARCH) != 0;
  const long long *n2 = reinterpret_cast<const long long *>(norm2);
  if (grouped && !force_search && m <= kCcStagedMaxM) {
    const int m_pad = round_up(m, 16);
    const size_t smem = size_t(m_pad) + 512;
    QRS_CUDA(cudaFuncSetAttribute(pair_centered_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  int(smem)));
    int per_sm = 1;
    QRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pair_centered_staged_kernel, kCcThreads, smem));
    if (per_sm < 1) return QRS_ERR_UNSUPPORTED;
    int64_t blocks = ceil_div<int64_t>(n_pairs, kCcChunk);
    const int64_t cap = int64_t(sm_count()) * per_sm;
    if (blocks > cap) blocks = cap;
    pair_centered_staged_kernel<<<unsigned(blocks), kCcThreads, smem, as_stream(stream)>>>(
        rowptr, colidx, val, cval, n2, pairs, n_pairs, m_pad, sim_milli);
  } else {
    int64_t blocks = ceil_div<int64_t>(n_pairs, 8);
    const int64_t cap = int64_t(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    pair_centered_search_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(rowptr, colidx, cval, n2, pairs,
                                                                               n_pairs, sim_milli);
  }
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}
