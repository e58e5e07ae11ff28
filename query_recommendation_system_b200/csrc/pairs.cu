// a5 + a8 — candidate pairs out of the grouped bucket records (lsh.py:40-54), de-duplicated
// across bands without a global set/sort, and scored with the rounded cosine of the
// signature vectors (recommender.py:203-204) in the same pass.
//
//   pair_slots : record p gets one "slot" per earlier record of its (hash24, band) run;
//                an exclusive scan of the slot counts load-balances the quadratic part.
//   pairs_score: one warp per slot.  It reads both signature rows ONCE and derives from
//                them (1) exact int16-lane equality in the slot's band (hash collisions are
//                discarded), (2) whether an EARLIER band already puts the two sets in one
//                bucket (then this slot is a duplicate of the pair emitted there), and
//                (3) the int64 dot product for the cosine.
#include "common.cuh"

namespace qrs {

constexpr int kPairThreads = 256;
constexpr int kPairWarps = kPairThreads / 32;
constexpr int kMaxSigWords = 32;   // K <= 1024

__global__ void __launch_bounds__(256)
pair_slot_count_kernel(const uint64_t *__restrict__ records, int64_t n, uint32_t *__restrict__ cnt) {
  for (int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < n; p += int64_t(gridDim.x) * blockDim.x) {
    const uint32_t key = rec_key(records[p]);
    uint32_t t = 0;
    if ((key >> 8) != QRS_SENTINEL_HASH) {
      while (int64_t(t) < p && rec_key(records[p - 1 - t]) == key) ++t;
    }
    cnt[p] = t;
  }
}

__device__ __forceinline__ double cosine_of(long long dot, long long n2a, long long n2b) {
  const double den = sqrt(double(n2a) * double(n2b));
  return den == 0.0 ? 0.0 : double(dot) / den;
}
__device__ __forceinline__ int to_milli(double c) { return int(rint(c * 1000.0)); }

// r consecutive bits of a multi-word mask starting at bit `start` (r <= 32)
__device__ __forceinline__ uint32_t mask_bits(const uint32_t *m, int start, int r, int words) {
  const int w = start >> 5, o = start & 31;
  uint32_t v = m[w] >> o;
  if (o + r > 32 && w + 1 < words) v |= m[w + 1] << (32 - o);
  return r == 32 ? v : (v & ((1u << r) - 1u));
}

__global__ void __launch_bounds__(kPairThreads)
pairs_score_kernel(const uint64_t *__restrict__ records, int64_t n_records, const uint64_t *__restrict__ slot_off,
                   const int *__restrict__ sig, const long long *__restrict__ norm2, int64_t sig_id_base, int K,
                   int b, uint64_t *__restrict__ pairs, int *__restrict__ sim_milli, int64_t capacity,
                   unsigned long long *__restrict__ n_pairs) {
  __shared__ uint32_t s_eq[kPairWarps][kMaxSigWords];
  __shared__ uint32_t s_neg[kPairWarps][kMaxSigWords];
  __shared__ uint64_t s_pair[kPairWarps][32];
  __shared__ long long s_dot[kPairWarps][32];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int words = ceil_div(K, 32);
  const int r = K / b;
  const uint32_t full_r = r == 32 ? 0xFFFFFFFFu : ((1u << r) - 1u);
  const uint64_t total = slot_off[n_records];
  const uint64_t chunks = (total + 31) / 32;
  const uint64_t warp_global = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;

  for (uint64_t c = warp_global; c < chunks; c += nwarps) {
    // ---- each lane locates its own slot: record p = last index with slot_off[p] <= q
    const uint64_t q = c * 32 + lane;
    uint64_t recB = 0, recA = 0;
    const bool have = q < total;
    if (have) {
      int64_t lo = 0, hi = n_records;   // invariant: slot_off[lo] <= q < slot_off[hi]
      while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (slot_off[mid] <= q) lo = mid; else hi = mid;
      }
      const uint64_t tq = q - slot_off[lo];
      const uint64_t tp = slot_off[lo + 1] - slot_off[lo];
      recB = records[lo];
      recA = records[lo - int64_t(tp) + int64_t(tq)];
    }
    const unsigned have_mask = __ballot_sync(kFull, have);
    const int nslots = __popc(have_mask);
    int nout = 0;

    for (int sl = 0; sl < nslots; ++sl) {
      const uint64_t ra = __shfl_sync(kFull, recA, sl), rb = __shfl_sync(kFull, recB, sl);
      const int band = int(rec_band(rb));
      uint32_t ia = rec_id(ra), ib = rec_id(rb);
      if (ia > ib) { uint32_t t = ia; ia = ib; ib = t; }
      const int *rowa = sig + (int64_t(ia) - sig_id_base) * K;
      const int *rowb = sig + (int64_t(ib) - sig_id_base) * K;
      long long dot = 0;
      for (int w = 0; w < words; ++w) {
        const int t = w * 32 + lane;
        int va = 0, vb = 1;
        if (t < K) { va = __ldg(rowa + t); vb = __ldg(rowb + t); }
        const bool eq = ((va ^ vb) & 0xFFFF) == 0 && t < K;
        const uint32_t em = __ballot_sync(kFull, eq);
        const uint32_t nm = __ballot_sync(kFull, eq && (va & 0xFFFF) == 0xFFFF);
        if (lane == 0) { s_eq[warp][w] = em; s_neg[warp][w] = nm; }
        if (t < K) dot += (long long)va * (long long)vb;
      }
      __syncwarp();
      // band test by lane: bands lane, lane+32, ... up to and including `band`
      bool own = false, earlier = false;
      for (int b0 = 0; b0 <= band; b0 += 32) {
        const int bb = b0 + lane;
        bool match = false;
        if (bb <= band) {
          const uint32_t e = mask_bits(s_eq[warp], bb * r, r, words);
          const uint32_t ng = mask_bits(s_neg[warp], bb * r, r, words);
          match = (e == full_r) && (ng != full_r);
        }
        const uint32_t mm = __ballot_sync(kFull, match);
        const int own_lane = band - b0;   // lane holding `band` in this round, if any
        if (own_lane < 32) {
          own = (mm >> own_lane) & 1u;
          earlier = earlier || (mm & ((1u << own_lane) - 1u)) != 0;
        } else {
          earlier = earlier || mm != 0;
        }
      }
      __syncwarp();
      if (own && !earlier) {
        if (sim_milli != nullptr) dot = warp_sum_ll(dot);
        if (lane == 0) {
          s_pair[warp][nout] = (uint64_t(ia) << 32) | uint64_t(ib);
          s_dot[warp][nout] = dot;
        }
        ++nout;
      }
    }
    __syncwarp();
    if (nout > 0) {
      // the emitted pairs of this chunk: cosines in parallel (one lane each), one reservation, one store
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(n_pairs, (unsigned long long)nout);
      base = __shfl_sync(kFull, base, 0);
      if (lane < nout && int64_t(base + lane) < capacity) {
        const uint64_t pr = s_pair[warp][lane];
        pairs[base + lane] = pr;
        if (sim_milli != nullptr) {
          const int64_t a = int64_t(pr >> 32) - sig_id_base, bb = int64_t(pr & 0xFFFFFFFFu) - sig_id_base;
          sim_milli[base + lane] = to_milli(cosine_of(s_dot[warp][lane], norm2[a], norm2[bb]));
        }
      }
    }
    __syncwarp();
  }
}

// r == 4 and K a multiple of 128 (the headline configs 128/32x4 and 256/64x4): a lane's 16-byte load
// of row elements [128c + 4*lane, +4) IS band 32c + lane, so one ballot per 128 elements yields the
// whole band-match mask and no shared-memory mask staging is needed.
__global__ void __launch_bounds__(kPairThreads)
pairs_score_r4_kernel(const uint64_t *__restrict__ records, int64_t n_records, const uint64_t *__restrict__ slot_off,
                      const int *__restrict__ sig, const long long *__restrict__ norm2, int64_t sig_id_base, int K,
                      uint64_t *__restrict__ pairs, int *__restrict__ sim_milli, int64_t capacity,
                      unsigned long long *__restrict__ n_pairs) {
  __shared__ uint64_t s_pair[kPairWarps][32];
  __shared__ long long s_dot[kPairWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks128 = K >> 7;
  const uint64_t total = slot_off[n_records];
  const uint64_t chunks = (total + 31) / 32;
  const uint64_t warp_global = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;

  for (uint64_t c = warp_global; c < chunks; c += nwarps) {
    const uint64_t q = c * 32 + lane;
    uint64_t recB = 0, recA = 0;
    const bool have = q < total;
    if (have) {
      int64_t lo = 0, hi = n_records;   // invariant: slot_off[lo] <= q < slot_off[hi]
      while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (slot_off[mid] <= q) lo = mid; else hi = mid;
      }
      const uint64_t tq = q - slot_off[lo];
      const uint64_t tp = slot_off[lo + 1] - slot_off[lo];
      recB = records[lo];
      recA = records[lo - int64_t(tp) + int64_t(tq)];
    }
    const int nslots = __popc(__ballot_sync(kFull, have));
    int nout = 0;
    for (int sl = 0; sl < nslots; ++sl) {
      const uint64_t ra = __shfl_sync(kFull, recA, sl), rb = __shfl_sync(kFull, recB, sl);
      const int band = int(rec_band(rb));
      uint32_t ia = rec_id(ra), ib = rec_id(rb);
      if (ia > ib) { uint32_t t = ia; ia = ib; ib = t; }
      const int4 *rowa = reinterpret_cast<const int4 *>(sig + (int64_t(ia) - sig_id_base) * K) + lane;
      const int4 *rowb = reinterpret_cast<const int4 *>(sig + (int64_t(ib) - sig_id_base) * K) + lane;
      long long dot = 0;
      bool own = false, earlier = false;
      for (int ch = 0; ch < chunks128; ++ch) {
        const int4 a = __ldg(rowa + ch * 32), b4 = __ldg(rowb + ch * 32);
        const unsigned dx = unsigned(a.x ^ b4.x) | unsigned(a.y ^ b4.y) | unsigned(a.z ^ b4.z) | unsigned(a.w ^ b4.w);
        const unsigned an = unsigned(a.x & a.y & a.z & a.w);
        const bool match = (dx & 0xFFFFu) == 0u && (an & 0xFFFFu) != 0xFFFFu;   // equal int16 lanes, not all -1
        const unsigned mm = __ballot_sync(kFull, match);
        const int rel = band - ch * 32;                                          // band's lane in this chunk
        if (rel >= 32) earlier = earlier || mm != 0u;
        else if (rel >= 0) {
          own = (mm >> rel) & 1u;
          earlier = earlier || (mm & ((1u << rel) - 1u)) != 0u;
        }
        dot += (long long)a.x * b4.x + (long long)a.y * b4.y + (long long)a.z * b4.z + (long long)a.w * b4.w;
      }
      if (own && !earlier) {
        if (sim_milli != nullptr) dot = warp_sum_ll(dot);
        if (lane == 0) {
          s_pair[warp][nout] = (uint64_t(ia) << 32) | uint64_t(ib);
          s_dot[warp][nout] = dot;
        }
        ++nout;
      }
    }
    __syncwarp();
    if (nout > 0) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(n_pairs, (unsigned long long)nout);
      base = __shfl_sync(kFull, base, 0);
      if (lane < nout && int64_t(base + lane) < capacity) {
        const uint64_t pr = s_pair[warp][lane];
        pairs[base + lane] = pr;
        if (sim_milli != nullptr) {
          const int64_t a = int64_t(pr >> 32) - sig_id_base, bb = int64_t(pr & 0xFFFFFFFFu) - sig_id_base;
          sim_milli[base + lane] = to_milli(cosine_of(s_dot[warp][lane], norm2[a], norm2[bb]));
        }
      }
    }
    __syncwarp();
  }
}

// The same for K = 128 * CH with CH known at compile time (CH = 1, 2: the headline configs), restructured around
// what bound the kernel above — one slot at a time per warp, every slot waiting a full L2 round trip for two
// 4K-byte rows:
//   * the slots of record p (p against every earlier record of its run) are consecutive, so p's row is loaded
//     once and kept in registers while it repeats (runs of taste-group size make this most of the slots);
//   * the other row of slot sl + 1 is requested before slot sl is evaluated;
//   * the 64-bit dot is formed only for slots that emit a pair (own band, no earlier band), and summed with three
//     REDUX on 16-bit pieces when the lane partials stay below 2^48 (five 64-bit shuffle steps otherwise).
template <int CH>
__global__ void __launch_bounds__(kPairThreads)
pairs_score_r4c_kernel(const uint64_t *__restrict__ records, int64_t n_records, const uint64_t *__restrict__ slot_off,
                       const int *__restrict__ sig, const long long *__restrict__ norm2, int64_t sig_id_base,
                       uint64_t *__restrict__ pairs, int *__restrict__ sim_milli, int64_t capacity,
                       unsigned long long *__restrict__ n_pairs) {
  __shared__ uint64_t s_pair[kPairWarps][32];
  __shared__ long long s_dot[kPairWarps][32];
  constexpr int K = 128 * CH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t total = slot_off[n_records];
  const uint64_t chunks = (total + 31) / 32;
  const uint64_t warp_global = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  auto row_of = [&](uint64_t rec) {
    return reinterpret_cast<const int4 *>(sig + (int64_t(rec_id(rec)) - sig_id_base) * K) + lane;
  };

  for (uint64_t c = warp_global; c < chunks; c += nwarps) {
    const uint64_t q = c * 32 + lane;
    uint64_t recB = 0, recA = 0;
    const bool have = q < total;
    if (have) {
      int64_t lo = 0, hi = n_records;   // invariant: slot_off[lo] <= q < slot_off[hi]  (a warp-cooperative 32-ary
                                        // search for the chunk's ends was measured: no gain, the top levels stay cached)
      while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (slot_off[mid] <= q) lo = mid; else hi = mid;
      }
      const uint64_t tq = q - slot_off[lo];
      const uint64_t tp = slot_off[lo + 1] - slot_off[lo];
      recB = records[lo];
      recA = records[lo - int64_t(tp) + int64_t(tq)];
    }
    const int nslots = __popc(__ballot_sync(kFull, have));
    int nout = 0;
    constexpr int AHEAD = 1;            // rows of later slots in flight while a slot is evaluated (the kernel is
                                        // issue-bound, 72 % of the slots: deeper read-ahead measured no gain)
    int4 an[AHEAD][CH], bq[CH];
    uint64_t cached = ~0ull;            // the record whose row sits in bq
#pragma unroll
    for (int d = 0; d < AHEAD; ++d) {
      if (d < nslots) {                 // warp-uniform
        const int4 *r0 = row_of(__shfl_sync(kFull, recA, d));
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) an[d][ch] = __ldg(r0 + ch * 32);
      }
    }
#pragma unroll 2
    for (int sl = 0; sl < nslots; ++sl) {
      const uint64_t ra = __shfl_sync(kFull, recA, sl), rb = __shfl_sync(kFull, recB, sl);
      int4 a[CH];
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) a[ch] = an[0][ch];
#pragma unroll
      for (int d = 0; d + 1 < AHEAD; ++d)
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) an[d][ch] = an[d + 1][ch];
      if (rb != cached) {               // warp-uniform
        const int4 *rowb = row_of(rb);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) bq[ch] = __ldg(rowb + ch * 32);
        cached = rb;
      }
      if (sl + AHEAD < nslots) {        // warp-uniform: a later slot's other row, before this slot is evaluated
        const int4 *r1 = row_of(__shfl_sync(kFull, recA, sl + AHEAD));
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) an[AHEAD - 1][ch] = __ldg(r1 + ch * 32);
      }
      const int band = int(rec_band(rb));
      bool own = false, earlier = false;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        const int4 x = a[ch], y = bq[ch];
        const unsigned dx = unsigned(x.x ^ y.x) | unsigned(x.y ^ y.y) | unsigned(x.z ^ y.z) | unsigned(x.w ^ y.w);
        const unsigned an4 = unsigned(x.x & x.y & x.z & x.w);
        const bool match = (dx & 0xFFFFu) == 0u && (an4 & 0xFFFFu) != 0xFFFFu;   // equal int16 lanes, not all -1
        const unsigned mm = __ballot_sync(kFull, match);
        const int rel = band - ch * 32;                                          // band's lane in this chunk
        if (rel >= 32) earlier = earlier || mm != 0u;
        else if (rel >= 0) {
          own = (mm >> rel) & 1u;
          earlier = earlier || (mm & ((1u << rel) - 1u)) != 0u;
        }
      }
      if (own && !earlier) {            // warp-uniform
        long long dot = 0;
        if (sim_milli != nullptr) {
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            const int4 x = a[ch], y = bq[ch];
            dot += (long long)x.x * y.x + (long long)x.y * y.y + (long long)x.z * y.z + (long long)x.w * y.w;
          }
          // rows that reach a pair hold no -1 (an empty set has sentinel records only), so the partials are >= 0
          if (__any_sync(kFull, (unsigned long long)dot >> 48)) dot = warp_sum_ll(dot);
          else dot = (long long)warp_sum_u48((unsigned long long)dot);
        }
        if (lane == 0) {
          const uint32_t ia = rec_id(ra), ib = rec_id(rb);
          s_pair[warp][nout] = (uint64_t(min(ia, ib)) << 32) | uint64_t(max(ia, ib));
          s_dot[warp][nout] = dot;
        }
        ++nout;
      }
    }
    __syncwarp();
    if (nout > 0) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(n_pairs, (unsigned long long)nout);
      base = __shfl_sync(kFull, base, 0);
      if (lane < nout && int64_t(base + lane) < capacity) {
        const uint64_t pr = s_pair[warp][lane];
        pairs[base + lane] = pr;
        if (sim_milli != nullptr) {
          const int64_t a0 = int64_t(pr >> 32) - sig_id_base, b0 = int64_t(pr & 0xFFFFFFFFu) - sig_id_base;
          sim_milli[base + lane] = to_milli(cosine_of(s_dot[warp][lane], norm2[a0], norm2[b0]));
        }
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
pair_cosine_kernel(const int *__restrict__ sig, const long long *__restrict__ norm2, int K,
                   const uint64_t *__restrict__ pairs, int64_t n_pairs, int *__restrict__ sim_milli) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t p = warp_global; p < n_pairs; p += nwarps) {
    const uint64_t pr = pairs[p];
    const int64_t ia = int64_t(pr >> 32), ib = int64_t(pr & 0xFFFFFFFFu);
    const int *rowa = sig + ia * K, *rowb = sig + ib * K;
    long long dot = 0;
    for (int t = lane; t < K; t += 32) dot += (long long)__ldg(rowa + t) * (long long)__ldg(rowb + t);
    dot = warp_sum_ll(dot);
    if (lane == 0) sim_milli[p] = to_milli(cosine_of(dot, norm2[ia], norm2[ib]));
  }
}

}  // namespace qrs

using namespace qrs;

extern "C" {

size_t qrs_pair_slots_workspace_bytes(int64_t n_records) {
  if (n_records < 0) return 0;
  return round_up<size_t>(size_t(n_records > 0 ? n_records : 1) * 4, 256) + scan_workspace_bytes(n_records);
}

int qrs_pair_slots(const uint64_t *records, int64_t n_records, uint64_t *slot_off, void *ws, size_t ws_bytes,
                   void *stream) {
  if (n_records < 0 || !slot_off || (n_records > 0 && !records)) return QRS_ERR_INVALID_ARG;
  if (!ws || ws_bytes < qrs_pair_slots_workspace_bytes(n_records)) return QRS_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  uint32_t *cnt = static_cast<uint32_t *>(ws);
  void *scan_ws = static_cast<char *>(ws) + round_up<size_t>(size_t(n_records > 0 ? n_records : 1) * 4, 256);
  if (n_records > 0) {
    int64_t blocks = ceil_div<int64_t>(n_records, 256);
    const int64_t cap = int64_t(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    pair_slot_count_kernel<<<unsigned(blocks), 256, 0, st>>>(records, n_records, cnt);
    QRS_LAUNCH_CHECK();
  }
  return exclusive_scan_u32_to_u64(cnt, slot_off, n_records, scan_ws, scan_workspace_bytes(n_records), st);
}

int qrs_pairs_score(const uint64_t *records, int64_t n_records, const uint64_t *slot_off, const int32_t *sig,
                    const int64_t *norm2, int64_t sig_id_base, int32_t K, int32_t b, uint64_t *pairs,
                    int32_t *sim_milli, int64_t capacity, unsigned long long *n_pairs, void *stream) {
  if (n_records < 0 || !slot_off || !sig || !n_pairs || K <= 0 || b <= 0 || K % b != 0 || capacity < 0)
    return QRS_ERR_INVALID_ARG;
  if (capacity > 0 && !pairs) return QRS_ERR_INVALID_ARG;
  if (sim_milli && !norm2) return QRS_ERR_INVALID_ARG;
  if (K > kMaxSigWords * 32 || K / b > 32 || b > QRS_MAX_BANDS) return QRS_ERR_UNSUPPORTED;
  if (n_records == 0) return QRS_OK;
  const int grid = sm_count() * 8;
  // the r = 4 kernel reads signature rows with 16-byte loads: only for a 16-byte aligned matrix (rows are K * 4 bytes,
  // K % 128 == 0); anything else takes the generic kernel's 4-byte loads
  if (K / b == 4 && (K == 128 || K == 256) && (reinterpret_cast<uintptr_t>(sig) & 15) == 0) {
    const long long *n2 = reinterpret_cast<const long long *>(norm2);
    if (K == 128)
      pairs_score_r4c_kernel<1><<<grid, kPairThreads, 0, as_stream(stream)>>>(records, n_records, slot_off, sig, n2,
                                                                              sig_id_base, pairs, sim_milli, capacity, n_pairs);
    else
      pairs_score_r4c_kernel<2><<<grid, kPairThreads, 0, as_stream(stream)>>>(records, n_records, slot_off, sig, n2,
                                                                              sig_id_base, pairs, sim_milli, capacity, n_pairs);
    QRS_LAUNCH_CHECK();
    return QRS_OK;
  }
  if (K / b == 4 && K % 128 == 0 && (reinterpret_cast<uintptr_t>(sig) & 15) == 0) {
    pairs_score_r4_kernel<<<grid, kPairThreads, 0, as_stream(stream)>>>(
        records, n_records, slot_off, sig, reinterpret_cast<const long long *>(norm2), sig_id_base, K, pairs, sim_milli,
        capacity, n_pairs);
    QRS_LAUNCH_CHECK();
    return QRS_OK;
  }
  pairs_score_kernel<<<grid, kPairThreads, 0, as_stream(stream)>>>(
      records, n_records, slot_off, sig, reinterpret_cast<const long long *>(norm2), sig_id_base, K, b, pairs,
      sim_milli, capacity, n_pairs);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

int qrs_pair_cosine(const int32_t *sig, const int64_t *norm2, int32_t K, const uint64_t *pairs, int64_t n_pairs,
                    int32_t *sim_milli, void *stream) {
  if (!sig || !norm2 || K <= 0 || n_pairs < 0 || !sim_milli || (n_pairs > 0 && !pairs)) return QRS_ERR_INVALID_ARG;
  if (n_pairs == 0) return QRS_OK;
  int64_t blocks = ceil_div<int64_t>(n_pairs, 8);
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  pair_cosine_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(
      sig, reinterpret_cast<const long long *>(norm2), K, pairs, n_pairs, sim_milli);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // extern "C"
