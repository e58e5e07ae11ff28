// a1 — MinHash signatures (recommender.py:105-143) with the band-record / norm2 epilogue fused.
//
// sig[s][i] = min_{e in set(s)} perm_i[e]  ==  the first rank r (in increasing order) whose
// element order_i[r] belongs to set(s) — the reference's own first-touch walk
// (recommender.py:125-129).  The first-touch kernel therefore keeps a membership map of the
// set in shared memory and, per hash, lets the 32 lanes of a warp test 32 consecutive ranks
// at once; the expected number of rounds is m / (32 * |set|) + 1, i.e. ~|set|^2 / m times
// fewer table reads than the per-nonzero gather-min ("direct") formulation, which remains
// for short sets and as the fallback when a hash is not resolved within the R-rank head.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace qrs {

constexpr int kByteMapMaxElems = 16384;           // one byte per element up to here, one bit beyond
// Beyond the tile scan's reach the membership map of ONE set fills most of an SM's shared memory and the per-set scan
// runs at one CTA per SM; sets with |S|^2 < kDirectFactor * m then go to a separate, map-less launch of the direct
// gather (64 warps per SM, eight table rows in flight per lane): 10.4 against 264 ms for 131 072 sets of ~1 000 of
// 1 000 000 elements at K = 128 (profiles/r2_minhash_modes_c4.jsonl); the gather grows with |S|, the scan with m / |S|.
constexpr int64_t kDirectFactor = 16;
constexpr int kMaxMapBytes = 200 * 1024;          // shared-memory budget for the membership map
constexpr int64_t kMaxElemsFirstTouch = int64_t(kMaxMapBytes) * 8;

// order_head layout (chunk-major): the element of rank r of hash i sits at
//   ((r / 8) * Kp + i) * 8 + (r % 8),   Kp = round_up(K, 32)   (hashes K..Kp-1 are zero padding)
// so that a lane that owns hash i takes its next 8 ranks with two 16-byte loads and a warp's 32 lanes read one
// contiguous kilobyte; 32 consecutive ranks of one hash (the per-set scan) are four 32-byte sectors.
__host__ __device__ constexpr int head_pitch(int K) { return round_up(K, 32); }
__device__ __forceinline__ int64_t head_at(int Kp, int i, int r) {
  return ((int64_t(r >> 3) * Kp + i) << 3) + (r & 7);
}

__global__ void perm_head_invert_kernel(const int *__restrict__ perm_tab, int m, int K, int R,
                                        int *__restrict__ order_head) {
  const int64_t total = int64_t(m) * K;
  const int Kp = head_pitch(K);
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += int64_t(gridDim.x) * blockDim.x) {
    int r = perm_tab[t];
    if (r < R) {
      int e = int(t / K), i = int(t - int64_t(e) * K);
      order_head[head_at(Kp, i, r)] = e << 2;
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
  const int Kp_head = head_pitch(K);

  if (map_bytes > 0) {
    for (int t = tid * 16; t < map_bytes; t += THREADS * 16)
      *reinterpret_cast<uint4 *>(map + t) = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();

  // bits 9..10 of the mode word: row filter of a split launch (1: rows with |S|^2 < kDirectFactor * m only — and the
  // empty ones; 2: the others only)
  const int filt = (mode >> 9) & 3;
  mode &= 0xFF;
  for (int64_t s = blockIdx.x; s < n; s += gridDim.x) {
    const int64_t lo = rowptr[s];
    const int len = int(rowptr[s + 1] - lo);
    if (filt != 0) {                      // CTA-uniform
      const bool short_row = int64_t(len) * len < kDirectFactor * int64_t(m);
      if ((filt == 1) != short_row) continue;
    }
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
            e[u] = __ldg(order_head + head_at(Kp_head, i, lane)) >> 2;
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
                int ee = __ldg(order_head + head_at(Kp_head, i, r0 + lane)) >> 2;
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
          for (; t + 8 <= cnt; t += 8) {          // eight table rows in flight per lane (the table is far larger than L2
            int ev[8], vv[8];                     // at the shapes that take this path)
#pragma unroll
            for (int u = 0; u < 8; ++u) ev[u] = __shfl_sync(kFull, mine, t + u);
#pragma unroll
            for (int u = 0; u < 8; ++u) vv[u] = __ldg(perm_tab + int64_t(ev[u]) * K + i);
#pragma unroll
            for (int u = 0; u < 8; ++u) v = min(v, vv[u]);
          }
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

// ---------------------------------------------------------------------------------------------
// Tile scan (QRS_MINHASH_TILE): the same first-touch walk for SETS consecutive sets at once.
//
// Shared memory holds one SETS-bit word per element: bit j of word e says "element e belongs
// to set s0 + j".  A lane owns one hash and walks its permutation in rank order; each rank
// costs ONE word load that answers the membership question for all SETS sets.  The walk
// ends when every (set, hash) of the warp is resolved: about ln(32 * SETS) * m / |set| ranks.
//   * SETS = 32 (32-bit words, m <= 32768, 4 warps, 4 CTAs per SM at the bench shape), 16 (16-bit words,
//     8 warps, one CTA per SM up to m ~ 100k) or 8 (bytes, 8 warps, up to m ~ 200k);
//   * the word map is built without atomics: a byte of every word — 8 sets — belongs to a group of
//     G warps (G = 1 / 4 / 8), which add those sets one after the other with byte read-modify-writes,
//     the G warps taking alternate 256-entry steps of the row and meeting at a named barrier between sets;
//   * order_head is chunk-major (head_at): a lane takes the next 8 ranks of its hash with two
//     16-byte loads, four chunks ahead of their use — nothing is staged or transposed;
//   * hits are NOT recorded one by one.  The rank at which set j is first met is kept BIT-SLICED:
//     plane p of a lane is a word whose bit j is bit p of that rank.  A set is met exactly
//     once (the word is masked with `found`), so a rank is recorded by OR-ing the new bits into the
//     planes selected by the rank's binary digits — the six low planes by position inside the
//     unrolled 64-rank group (one LOP3 each: plane |= word & ~found), the high planes once per group
//     with a warp-uniform mask of the group index.  No branch, no divergence, ~6 instructions per
//     rank for 32 x SETS (set, hash) pairs; the planes are un-sliced once per pass (SETS = 32: two sets
//     per operation, bits j and j + 16 land in the two halves of one register);
//   * short sets (len*len < m, AUTO only) are left out of the map and take the direct
//     gather-min; a (set, hash) still open after R ranks takes direct_min_one_hash;
//   * the signature tile is staged 16 bits wide when m <= 32768 (ranks and values fit), else as ints.
//   dynamic smem: long long rp[40] | map (m words of SETS bits + a spare 16 bytes) | sig_t sigt[SETS][Kp]
constexpr int kTileMaxElems16 = 32768;         // 16-bit signature tile (and the 32-set form) up to here
constexpr int kTileMaxSmemBytes = 226 * 1024;
constexpr bool kAutoUsesTile = true;   // AUTO takes the tile scan whenever a word map fits

__device__ __forceinline__ void st_shared_b16(unsigned addr, unsigned v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"((unsigned short)(v)) : "memory");
}
__device__ __forceinline__ void st_shared_b32(unsigned addr, unsigned v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int4 ld_stream_v4(const int *p) {      // 16-byte streaming (read-once) load
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// membership word of the element whose head-table entry (4 * element) is `entry`
template <int SETS>
__device__ __forceinline__ unsigned ld_map(unsigned bm_s, int entry) {
  unsigned v;
  if (SETS == 32) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(bm_s + unsigned(entry)));
  else if (SETS == 16) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(bm_s + (unsigned(entry) >> 1)));
  else asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(bm_s + (unsigned(entry) >> 2)));
  return v;
}
__device__ __forceinline__ unsigned ld_map_u8(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_map_u8(unsigned addr, unsigned v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// plane | (word & ~found) in one LOP3
__device__ __forceinline__ unsigned or_andn(unsigned plane, unsigned word, unsigned found) {
  unsigned d;
  asm("lop3.b32 %0, %1, %2, %3, 0xF4;" : "=r"(d) : "r"(plane), "r"(word), "r"(found));
  return d;
}

template <int SETS>
struct TileShape {
  static constexpr int kWarps = SETS == 32 ? 4 : 8;
  static constexpr int kThreads = kWarps * 32;
  static constexpr int kGroup = kWarps / (SETS / 8);      // warps that share one byte of the map words
  static constexpr int kWordBytes = SETS / 8;
  static constexpr int kMinCtas = SETS == 32 ? 4 : (SETS == 16 ? 1 : 2);
  static constexpr int kSetsPerWarp = SETS / kWarps;
};

// signature-tile pitch in elements: an odd number of 32-bit words per row, so the rows start in different banks
__host__ __device__ constexpr int tile_sig_stride(int K, bool wide) { return round_up(K, 32) + (wide ? 1 : 2); }

// NH = number of high planes: ranks < 8 << NH.  WIDE: signature values need more than 15 bits (int tile).
template <int SETS, int NH, bool WIDE>
__global__ void __launch_bounds__(TileShape<SETS>::kThreads, TileShape<SETS>::kMinCtas)
minhash_tile_kernel(const int64_t *__restrict__ rowptr, const int *__restrict__ colidx, int64_t n, int m,
                    const int *__restrict__ perm_tab, const int *__restrict__ order_head, int R, int K, int b,
                    int64_t id_base, int *__restrict__ sig, long long *__restrict__ norm2,
                    uint64_t *__restrict__ records, int64_t rec_stride, int mode, int map_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Shape = TileShape<SETS>;
  using sig_t = typename std::conditional<WIDE, int, short>::type;
  constexpr int NW = Shape::kWarps, G = Shape::kGroup, WB = Shape::kWordBytes, SPW = Shape::kSetsPerWarp;
  constexpr int NP = 3 + NH;
  static_assert(NP <= 16, "two un-sliced ranks share one register");
  static_assert(NH >= 3, "the three lowest chunk digits are static");
  static_assert(SETS == 32 || WIDE, "the narrower maps exist for element counts beyond the 16-bit tile");
  const int Kp = tile_sig_stride(K, WIDE);
  long long *rp = reinterpret_cast<long long *>(smem_raw);
  unsigned char *bmb = reinterpret_cast<unsigned char *>(rp + 40);        // 16-byte aligned
  sig_t *sigt = reinterpret_cast<sig_t *>(bmb + map_bytes + 16);          // map_bytes % 16 == 0; the spare word

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool force_scalar = (mode & 0x100) != 0;
  mode &= 0xFF;
  const int r_band = b > 0 ? K / b : 0;
  const int64_t tiles = ceil_div<int64_t>(n, SETS);
  const int nchunks = R >> 3;
  const size_t chunk_stride = size_t(head_pitch(K)) << 1;          // int4 units per 8-rank chunk
  const unsigned sigt_s = unsigned(__cvta_generic_to_shared(sigt));
  const unsigned bm_s = unsigned(__cvta_generic_to_shared(bmb));
  const unsigned row_bytes = unsigned(Kp) * unsigned(sizeof(sig_t));

  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t s0 = tile * SETS;
    const int nsets = int(min(int64_t(SETS), n - s0));

    // ---- stage the row bounds, clear the map
    if (tid <= SETS) rp[tid] = rowptr[s0 + min(tid, nsets)];
    for (int t = tid * 16; t < map_bytes; t += Shape::kThreads * 16)
      *reinterpret_cast<uint4 *>(bmb + t) = make_uint4(0, 0, 0, 0);
    __syncthreads();

    // scanmask: sets that take part in the walk; the others are empty (-1 row), absent, or short
    unsigned scanmask = 0;
    {
      bool scan = false;
      if (lane < nsets) {
        const long long len = rp[lane + 1] - rp[lane];
        scan = len > 0 && (mode == QRS_MINHASH_TILE || len * len >= m);
      }
      scanmask = __ballot_sync(kFull, scan);
    }

    // ---- empty sets: a row of -1
    for (int jj = 0; jj < SPW; ++jj) {
      const int j = warp * SPW + jj;
      if (j < nsets && rp[j + 1] == rp[j])
        for (int i = lane; i < K; i += 32) sigt[j * Kp + i] = sig_t(-1);
    }

    // ---- build: byte `bidx` of the words (sets 8 * bidx ..) belongs to G warps, which add those sets one after
    //      the other, warp g of the group taking the 256-entry steps g, g + G, ... of a row; the NEXT step's
    //      column indices are loaded before this step's bytes are updated.  Entries past the end of a row point
    //      at a spare word behind the map, so the read-modify-writes carry no predicate, and a step that lies
    //      inside its row loads without bounds checks.  Every warp of a group takes at least one (possibly empty)
    //      step per set and the group meets at a barrier before the next set.
    {
      const int bidx = warp / G, g = warp % G;
      unsigned todo = (scanmask >> (bidx * 8)) & 0xFFu;
      int jj = todo ? __ffs(todo) - 1 : -1;
      int t0 = g * 256;
      int e[8];
      const int spare = map_bytes / WB;
      const bool vec = (reinterpret_cast<uintptr_t>(colidx) & 15) == 0 && !force_scalar;
      const int64_t nnz_all = rowptr[n];
      // 128-bit form (colidx 16-byte aligned): a row is walked from its start rounded DOWN to 16 bytes, a lane takes
      // entries 4 * lane .. + 3 and 128 + 4 * lane .. + 3 of a step with two 16-byte loads; what lies before the row
      // (up to 3 entries of the previous one), behind it, or behind the end of the array points at the spare word.
      auto row_head = [&](int set) -> int { return vec ? int(rp[bidx * 8 + set] & 3) : 0; };
      auto load8 = [&](int (&dst)[8], int set, int first) {
        const int64_t lo = rp[bidx * 8 + set];
        const int len = int(rp[bidx * 8 + set + 1] - lo);
        if (vec) {
          const int head = int(lo & 3), vlen = len + head;
          const int64_t vbase = lo - head;
          const bool interior = first >= head && first + 256 <= vlen;          // warp-uniform: every entry is the row's
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int vp = first + 128 * u + 4 * lane;
            int4 v = make_int4(spare, spare, spare, spare);
            if (interior || (vp < vlen && vbase + vp + 4 <= nnz_all)) {
              v = ld_stream_v4(colidx + vbase + vp);
            } else if (vp < vlen) {                                            // the last 16 bytes of the array
              if (vbase + vp < nnz_all) v.x = ld_stream_s32(colidx + vbase + vp);
              if (vbase + vp + 1 < nnz_all) v.y = ld_stream_s32(colidx + vbase + vp + 1);
              if (vbase + vp + 2 < nnz_all) v.z = ld_stream_s32(colidx + vbase + vp + 2);
            }
            if (!interior) {
              if (vp < head || vp >= vlen) v.x = spare;
              if (vp + 1 < head || vp + 1 >= vlen) v.y = spare;
              if (vp + 2 < head || vp + 2 >= vlen) v.z = spare;
              if (vp + 3 < head || vp + 3 >= vlen) v.w = spare;
            }
            dst[4 * u] = v.x; dst[4 * u + 1] = v.y; dst[4 * u + 2] = v.z; dst[4 * u + 3] = v.w;
          }
          return;
        }
        const int *src = colidx + lo + first + lane;
        if (first + 256 <= len) {          // warp-uniform
#pragma unroll
          for (int u = 0; u < 8; ++u) dst[u] = ld_stream_s32(src + 32 * u);
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) dst[u] = first + lane + 32 * u < len ? ld_stream_s32(src + 32 * u) : spare;
        }
      };
      // one step: the entries in `cur` (set jj, from t0) go into the map while the next step's are loaded into
      // `nxt`; the two register sets swap roles from step to step (no copies); false when the warp is done
      const unsigned bmw_s = bm_s + unsigned(bidx);
      auto step = [&](int (&cur)[8], int (&nxt)[8]) -> bool {
        const int len = int(rp[bidx * 8 + jj + 1] - rp[bidx * 8 + jj]) + row_head(jj);      // steps count from the aligned start
        int njj = jj, nt0 = t0 + G * 256;
        if (nt0 >= len) {
          todo &= todo - 1;
          njj = todo ? __ffs(todo) - 1 : -1;
          nt0 = g * 256;
        }
        if (njj >= 0) load8(nxt, njj, nt0);
        const unsigned bit = 1u << jj;
        unsigned old[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) old[u] = ld_map_u8(bmw_s + unsigned(WB) * unsigned(cur[u]));   // distinct elements
#pragma unroll                                                       // of one row: no two lanes share a byte (but the spare)
        for (int u = 0; u < 8; ++u) st_map_u8(bmw_s + unsigned(WB) * unsigned(cur[u]), old[u] | bit);
        if (njj != jj) {                   // the next set may touch the bytes other lanes / warps just wrote
          if (G == 1) __syncwarp();
          else asm volatile("bar.sync %0, %1;" ::"r"(1 + bidx), "n"(G * 32) : "memory");
        }
        jj = njj;
        t0 = nt0;
        return jj >= 0;
      };
      if (jj >= 0) {
        int en[8];
        load8(e, jj, t0);
        while (step(e, en) && step(en, e)) {
        }
      }
    }
    __syncthreads();

    // the next tile's column indices (one contiguous CSR span, read once) on their way into L2
    if (tile + gridDim.x < tiles) {
      const int64_t sn = (tile + gridDim.x) * SETS;
      const int64_t a = rowptr[sn], e = rowptr[min(sn + SETS, n)];
      const char *base = reinterpret_cast<const char *>(colidx + a);
      for (int64_t off = int64_t(tid) * 128; off < (e - a) * 4; off += Shape::kThreads * 128) prefetch_l2(base + off);
    }

    // ---- walk: lane = hash
    for (int i0 = warp * 32; i0 < K; i0 += NW * 32) {
      const int i = i0 + lane;             // i < head_pitch(K): the padding hashes read element 0 and write nothing
      unsigned found = (i < K) ? ~scanmask : kFull;
      unsigned P[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) P[p] = 0u;
      const int4 *hp = reinterpret_cast<const int4 *>(order_head) + (size_t(i) << 1);
      // four chunks in flight (R >= 32: they exist); buffer c % 4 is re-loaded with chunk c + 4 as soon as its
      // eight word loads have been issued
      int4 a0 = __ldg(hp), b0 = __ldg(hp + 1);
      int4 a1 = __ldg(hp + chunk_stride), b1 = __ldg(hp + chunk_stride + 1);
      int4 a2 = __ldg(hp + 2 * chunk_stride), b2 = __ldg(hp + 2 * chunk_stride + 1);
      int4 a3 = __ldg(hp + 3 * chunk_stride), b3 = __ldg(hp + 3 * chunk_stride + 1);
      const int4 *nx = hp + 4 * chunk_stride;                    // the chunk the next re-load takes
      // one 8-rank chunk, CC = its position inside a 64-rank group (compile time: the rank's six low digits are
      // static); true = stop (every pair of the warp resolved, or the head is exhausted).  Table entries are
      // byte offsets into a 32-bit word map (4 * element).
      auto chunk = [&](auto cc_tag, int c, int4 &qa, int4 &qb) -> bool {
        constexpr int CC = decltype(cc_tag)::value;
        if (c >= nchunks || __all_sync(kFull, found == kFull)) return true;
        const unsigned w0 = ld_map<SETS>(bm_s, qa.x), w1 = ld_map<SETS>(bm_s, qa.y);
        const unsigned w2 = ld_map<SETS>(bm_s, qa.z), w3 = ld_map<SETS>(bm_s, qa.w);
        const unsigned w4 = ld_map<SETS>(bm_s, qb.x), w5 = ld_map<SETS>(bm_s, qb.y);
        const unsigned w6 = ld_map<SETS>(bm_s, qb.z), w7 = ld_map<SETS>(bm_s, qb.w);
        if (c + 4 < nchunks) {             // warp-uniform
          qa = __ldg(nx);
          qb = __ldg(nx + 1);
        }
        nx += chunk_stride;
        const unsigned before = found;
        found |= w0;
        P[0] = or_andn(P[0], w1, found); found |= w1;
        P[1] = or_andn(P[1], w2, found); found |= w2;
        P[0] = or_andn(P[0], w3, found); P[1] = or_andn(P[1], w3, found); found |= w3;
        P[2] = or_andn(P[2], w4, found); found |= w4;
        P[0] = or_andn(P[0], w5, found); P[2] = or_andn(P[2], w5, found); found |= w5;
        P[1] = or_andn(P[1], w6, found); P[2] = or_andn(P[2], w6, found); found |= w6;
        P[0] = or_andn(P[0], w7, found); P[1] = or_andn(P[1], w7, found); P[2] = or_andn(P[2], w7, found);
        found |= w7;
        if (CC & 1) P[3] = or_andn(P[3], found, before);
        if (CC & 2) P[4] = or_andn(P[4], found, before);
        if (CC & 4) P[5] = or_andn(P[5], found, before);
        return false;
      };
      for (int c8 = 0;; c8 += 8) {
        const unsigned before8 = found;
        const bool stop = chunk(std::integral_constant<int, 0>{}, c8, a0, b0) ||
                          chunk(std::integral_constant<int, 1>{}, c8 + 1, a1, b1) ||
                          chunk(std::integral_constant<int, 2>{}, c8 + 2, a2, b2) ||
                          chunk(std::integral_constant<int, 3>{}, c8 + 3, a3, b3) ||
                          chunk(std::integral_constant<int, 4>{}, c8 + 4, a0, b0) ||
                          chunk(std::integral_constant<int, 5>{}, c8 + 5, a1, b1) ||
                          chunk(std::integral_constant<int, 6>{}, c8 + 6, a2, b2) ||
                          chunk(std::integral_constant<int, 7>{}, c8 + 7, a3, b3);
        // digits 6.. of the ranks met in this 64-rank group: one warp-uniform mask per plane
        const unsigned new8 = found & ~before8;
#pragma unroll
        for (int h = 3; h < NH; ++h) P[3 + h] |= new8 & (0u - ((unsigned(c8) >> h) & 1u));
        if (stop) break;
      }

      // ---- un-slice
      if (i < K) {
        const unsigned col_s = sigt_s + unsigned(i) * unsigned(sizeof(sig_t));
        if (SETS == 32 && !WIDE) {
          // sets j and j + 16 together: their bits of plane p go to bits p and p + 16
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const bool lo_on = (scanmask >> j) & 1u, hi_on = (scanmask >> (j + 16)) & 1u;     // warp-uniform
            if (lo_on || hi_on) {
              unsigned v2 = 0;
#pragma unroll
              for (int p = 0; p < NP; ++p) {
                const unsigned t = p >= j ? (P[p] << (p - j)) : (P[p] >> (j - p));
                v2 |= t & (0x00010001u << p);
              }
              if (lo_on) st_shared_b16(col_s + unsigned(j) * row_bytes, v2);
              if (hi_on) st_shared_b16(col_s + unsigned(j + 16) * row_bytes, v2 >> 16);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < SETS; ++j) {
            if ((scanmask >> j) & 1u) {                                                        // warp-uniform
              unsigned v = 0;
#pragma unroll
              for (int p = 0; p < NP; ++p) {
                const unsigned t = p >= j ? (P[p] << (p - j)) : (P[p] >> (j - p));
                v |= t & (1u << p);
              }
              if (WIDE) st_shared_b32(col_s + unsigned(j) * row_bytes, v);
              else st_shared_b16(col_s + unsigned(j) * row_bytes, v);
            }
          }
        }
      }
      __syncwarp();
      // still open after the head: gather-min per (set, hash), the whole warp on each
      unsigned miss = ~found;
      unsigned lanes_open = __ballot_sync(kFull, miss != 0);
      while (lanes_open) {
        const int src = __ffs(lanes_open) - 1;
        lanes_open &= lanes_open - 1;
        unsigned mm = __shfl_sync(kFull, miss, src);
        while (mm) {
          const int j = __ffs(mm) - 1;
          mm &= mm - 1;
          const int64_t lo = rp[j];
          const int v = direct_min_one_hash(colidx, lo, int(rp[j + 1] - lo), perm_tab, K, i0 + src, lane);
          if (lane == 0) sigt[j * Kp + i0 + src] = sig_t(v);
        }
      }
    }

    // ---- short sets (AUTO): direct gather-min, lane owns a hash, elements broadcast 32 at a time
    for (int jj = 0; jj < SPW; ++jj) {
      const int j = warp * SPW + jj;
      if (j >= nsets) break;
      const int64_t lo = rp[j];
      const int len = int(rp[j + 1] - lo);
      if (len == 0 || ((scanmask >> j) & 1u)) continue;
      for (int i0 = 0; i0 < K; i0 += 32) {
        const int i = min(i0 + lane, K - 1);
        int v = 0x7FFFFFFF;
        for (int t0 = 0; t0 < len; t0 += 32) {
          const int mine = (t0 + lane < len) ? __ldg(colidx + lo + t0 + lane) : -1;
          const int cnt = min(32, len - t0);
          for (int t = 0; t < cnt; ++t) {
            const int e0 = __shfl_sync(kFull, mine, t);
            v = min(v, __ldg(perm_tab + int64_t(e0) * K + i));
          }
        }
        if (i0 + lane < K) sigt[j * Kp + i0 + lane] = sig_t(v);
      }
    }
    __syncthreads();

    // ---- epilogue 1: the tile's rows to global (consecutive sets are one contiguous block of SETS * K ints) and
    //      their squared norms
    for (int j = warp; j < nsets; j += NW) {
      unsigned long long acc = 0;
      if (!WIDE && (K & 1) == 0) {
        int2 *dst = reinterpret_cast<int2 *>(sig + (s0 + j) * K);
        const unsigned *src = reinterpret_cast<const unsigned *>(sigt + j * Kp);     // Kp even: 4-byte aligned rows
        for (int q = lane; q < (K >> 1); q += 32) {
          const unsigned x = src[q];
          const int v0 = int(short(x & 0xFFFFu)), v1 = int(x) >> 16;
          dst[q] = make_int2(v0, v1);
          acc += unsigned(v0 * v0) + (unsigned long long)unsigned(v1 * v1);          // |v| <= 32768
        }
      } else {
        for (int i = lane; i < K; i += 32) {
          const long long v = sigt[j * Kp + i];
          sig[(s0 + j) * K + i] = int(v);
          acc += (unsigned long long)(v * v);
        }
      }
      if (norm2 != nullptr) {
        if (WIDE) acc = (unsigned long long)warp_sum_ll((long long)acc);
        else acc = warp_sum_u48(acc);      // K <= 8192 values below 2^30 per row
        if (lane == 0) norm2[s0 + j] = (long long)acc;
      }
    }
    // ---- epilogue 2: bucket records, lane = set (the rows' pitch, an odd number of words, keeps the lanes in
    //      different banks); a 16-bit tile with an even r reads the band's lanes two at a time, as the words the
    //      hash folds
    if (records != nullptr && lane < nsets) {
      const sig_t *row = sigt + lane * Kp;
      const uint32_t id = uint32_t(id_base + s0 + lane);
      if (!WIDE && (r_band & 1) == 0) {
        const unsigned *roww = reinterpret_cast<const unsigned *>(row);
        const int wpb = r_band >> 1;
        for (int band = warp; band < b; band += NW) {
          uint32_t h = band_hash_init(uint32_t(band));
          uint32_t all = 0xFFFFFFFFu;
          for (int c = 0; c < wpb; ++c) {
            const uint32_t w = roww[band * wpb + c];
            all &= w;
            h = band_hash_word(h, w);
          }
          records[int64_t(band) * rec_stride + s0 + lane] =
              make_record(band_hash_finish(h, all == 0xFFFFFFFFu), uint32_t(band), id);
        }
      } else {
        for (int band = warp; band < b; band += NW)
          records[int64_t(band) * rec_stride + s0 + lane] = band_record(row + band * r_band, r_band, uint32_t(band), id);
      }
    }
    __syncthreads();   // rp / sigt / the map are rewritten by the next tile
  }
}

// word-map bytes for m elements (a multiple of 16) and the whole dynamic shared memory of one CTA
static int tile_map_bytes(int m, int sets) { return round_up(m * (sets / 8), 16); }
static size_t tile_smem_bytes(int K, int m, int sets, bool wide) {
  return 40 * 8 + size_t(tile_map_bytes(m, sets)) + 16 + size_t(sets) * tile_sig_stride(K, wide) * (wide ? 4 : 2);
}

// sets per tile for this shape: 32 while a 16-bit tile can hold the values, then the widest map that fits; 0 = none
static int tile_sets_for(int m, int K) {
  if (m <= kTileMaxElems16 && tile_smem_bytes(K, m, 32, false) <= size_t(kTileMaxSmemBytes)) return 32;
  if (m <= kTileMaxElems16) return 0;
  if (tile_smem_bytes(K, m, 16, true) <= size_t(kTileMaxSmemBytes)) return 16;
  if (tile_smem_bytes(K, m, 8, true) <= size_t(kTileMaxSmemBytes)) return 8;
  return 0;
}

template <int SETS, int NH, bool WIDE>
static int launch_minhash_tile_t(const int64_t *rowptr, const int *colidx, int64_t n, int m, const int *perm_tab,
                                 const int *order_head, int R, int K, int b, int64_t id_base, int *sig,
                                 long long *norm2, uint64_t *records, int64_t rec_stride, int mode,
                                 cudaStream_t stream) {
  auto kern = minhash_tile_kernel<SETS, NH, WIDE>;
  const int map_bytes = tile_map_bytes(m, SETS);
  const size_t smem = tile_smem_bytes(K, m, SETS, WIDE);
  QRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  QRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TileShape<SETS>::kThreads, smem));
  if (per_sm < 1) return QRS_ERR_UNSUPPORTED;
  int64_t grid = int64_t(sm_count()) * per_sm;
  const int64_t tiles = ceil_div<int64_t>(n, SETS);
  if (grid > tiles) grid = tiles;
  kern<<<unsigned(grid), TileShape<SETS>::kThreads, smem, stream>>>(rowptr, colidx, n, m, perm_tab, order_head, R, K,
                                                                    b, id_base, sig, norm2, records, rec_stride, mode,
                                                                    map_bytes);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

template <int SETS, bool WIDE>
static int launch_minhash_tile_s(const int64_t *rowptr, const int *colidx, int64_t n, int m, const int *perm_tab,
                                 const int *order_head, int R, int K, int b, int64_t id_base, int *sig,
                                 long long *norm2, uint64_t *records, int64_t rec_stride, int mode,
                                 cudaStream_t stream) {
  if (R <= (8 << 6))
    return launch_minhash_tile_t<SETS, 6, WIDE>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig,
                                                norm2, records, rec_stride, mode, stream);
  if (R <= (8 << 9))
    return launch_minhash_tile_t<SETS, 9, WIDE>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig,
                                                norm2, records, rec_stride, mode, stream);
  if (R <= (8 << 12))
    return launch_minhash_tile_t<SETS, 12, WIDE>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig,
                                                 norm2, records, rec_stride, mode, stream);
  return QRS_ERR_UNSUPPORTED;
}

static int launch_minhash_tile(int sets, const int64_t *rowptr, const int *colidx, int64_t n, int m,
                               const int *perm_tab, const int *order_head, int R, int K, int b, int64_t id_base,
                               int *sig, long long *norm2, uint64_t *records, int64_t rec_stride, int mode,
                               cudaStream_t stream) {
  if (sets == 32)
    return launch_minhash_tile_s<32, false>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, norm2,
                                            records, rec_stride, mode, stream);
  if (sets == 16)
    return launch_minhash_tile_s<16, true>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, norm2,
                                           records, rec_stride, mode, stream);
  if (sets == 8)
    return launch_minhash_tile_s<8, true>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, norm2,
                                          records, rec_stride, mode, stream);
  return QRS_ERR_UNSUPPORTED;
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

int64_t qrs_perm_head_elems(int32_t K, int32_t R) {
  if (K <= 0 || R <= 0) return 0;
  return int64_t(head_pitch(K)) * R;
}

int qrs_perm_head_invert(const int32_t *perm_tab, int32_t m, int32_t K, int32_t R, int32_t *order_head,
                         void *stream) {
  if (!perm_tab || !order_head || m <= 0 || K <= 0 || R <= 0 || R > m) return QRS_ERR_INVALID_ARG;
  if (R % 8 != 0) return QRS_ERR_INVALID_ARG;
  const int64_t total = int64_t(m) * K;
  int64_t blocks = ceil_div<int64_t>(total, 256);
  const int64_t cap = int64_t(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  // the padding hashes K .. round_up(K, 32) - 1 read as element 0
  QRS_CUDA(cudaMemsetAsync(order_head, 0, size_t(qrs_perm_head_elems(K, R)) * 4, as_stream(stream)));
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
  if (mode < QRS_MINHASH_AUTO || mode > QRS_MINHASH_TILE) return QRS_ERR_INVALID_ARG;
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
  // sets per tile: the widest word map that fits shared memory (QRS_MINHASH_TILE_SETS=8|16 narrows it, for measurements)
  int tile_sets = (R >= 32 && R <= (8 << 12) && (reinterpret_cast<uintptr_t>(order_head) & 15) == 0) ? tile_sets_for(m, K) : 0;
  if (tile_sets == 16) {
    static const int forced = [] { const char *s = getenv("QRS_MINHASH_TILE_SETS"); return s ? atoi(s) : 0; }();
    if (forced == 8 && tile_smem_bytes(K, m, 8, true) <= size_t(kTileMaxSmemBytes)) tile_sets = 8;
  }
  if (mode == QRS_MINHASH_TILE && tile_sets == 0) return QRS_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  long long *n2 = reinterpret_cast<long long *>(norm2);
  if (tile_sets != 0 && (mode == QRS_MINHASH_TILE || (mode == QRS_MINHASH_AUTO && kAutoUsesTile)))
  {
    // The map build reads the column indices one 128-byte line per warp load (32 lanes x 4 bytes).  The 16-byte
    // form (a lane takes 4 consecutive entries, rows walked from their 16-byte aligned start) is kept behind
    // QRS_MINHASH_V4_CSR=1: measured SLOWER on B200 — 0.180 against 0.141 ms at C2, 2.57 against 2.25 ms per 200 k sets
    // at C3 (profiles/r2_minhash_csr_load_width.jsonl): the loads were never the limiter (the walk and the byte
    // read-modify-writes are), and masking the row ends per component costs more than the six saved load
    // instructions per step.  Bit 8 of the kernel's mode word = scalar form.
    static const bool scalar_csr = [] { const char *e = getenv("QRS_MINHASH_V4_CSR"); return !(e && atoi(e) != 0); }();
    return launch_minhash_tile(tile_sets, rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, n2,
                               records, rec_stride, mode | (scalar_csr ? 0x100 : 0), st);
  }
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
  if (mode == QRS_MINHASH_AUTO) {
    // split by row length (kDirectFactor): short and empty sets in a map-less direct launch at full occupancy, the
    // long ones in the scan; every row is written by exactly one of the two
    int rc = launch_minhash<true, 128>(rowptr, colidx, n, m, perm_tab, order_head, 0, K, b, id_base, sig, n2, records,
                                       rec_stride, QRS_MINHASH_DIRECT | (1 << 9), 0, st);
    if (rc != QRS_OK) return rc;
    return launch_minhash<false, 512>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, n2,
                                      records, rec_stride, QRS_MINHASH_FIRST_TOUCH | (2 << 9), map_bytes, st);
  }
  return launch_minhash<false, 512>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, n2,
                                    records, rec_stride, mode, map_bytes, st);
}

}  // extern "C"
