// a1 — MinHash signatures (recommender.py:105-143) with the band-record / norm2 epilogue fused.
//
// sig[s][i] = min_{e in set(s)} perm_i[e]  ==  the first rank r (in increasing order) whose
// element order_i[r] belongs to set(s) — the reference's own first-touch walk
// (recommender.py:125-129).  The first-touch kernel therefore keeps a membership map of the
// set in shared memory and, per hash, lets the 32 lanes of a warp test 32 consecutive ranks
// at once; the expected number of rounds is m / (32 * |set|) + 1, i.e. ~|set|^2 / m times
// fewer table reads than the per-nonzero gather-min ("direct") formulation, which remains
// for short sets and as the fallback when a hash is not resolved within the R-rank head.
#include <type_traits>

#include "common.cuh"

namespace qrs {

constexpr int kByteMapMaxElems = 16384;           // one byte per element up to here, one bit beyond
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
// Tile scan (QRS_MINHASH_TILE): the same first-touch walk for 32 consecutive sets at once.
//
// Shared memory holds one 32-bit word per element: bit j of word e says "element e belongs
// to set s0 + j".  A lane owns one hash and walks its permutation in rank order; each rank
// costs ONE word load that answers the membership question for all 32 sets.  The walk
// ends when every (set, hash) of the warp is resolved: about ln(1024) * m / |set| ranks.
//   * the word map is built without atomics: warp w owns the sets 8w .. 8w+7, i.e. byte w of
//     every word, and adds its sets one after the other with byte read-modify-writes;
//   * order_head is chunk-major (head_at): a lane takes the next 8 ranks of its hash with two
//     16-byte loads, three chunks ahead of their use — nothing is staged or transposed;
//   * hits are NOT recorded one by one.  The rank at which set j is first met is kept BIT-SLICED:
//     plane p of a lane is a 32-bit word whose bit j is bit p of that rank.  A set is met exactly
//     once (the word is masked with `found`), so a rank is recorded by OR-ing the new bits into the
//     planes selected by the rank's binary digits — the three low planes by position inside the
//     unrolled 8-rank chunk (one LOP3 each: plane |= word & ~found), the high planes once per chunk
//     with a warp-uniform mask of the chunk index.  No branch, no divergence, ~5 instructions per
//     rank for 32 x 32 (set, hash) pairs; the planes are un-sliced once per pass, two sets per
//     operation (bits j and j + 16 land in the two halves of one register).
//   * short sets (len*len < m, AUTO only) are left out of the map and take the direct
//     gather-min; a (set, hash) still open after R ranks takes direct_min_one_hash.
//   * the signature tile is staged 16 bits wide (m <= 32768 bounds ranks and signature values);
//     four CTAs share an SM at the bench shape.
//   dynamic smem: long long rp[40] | unsigned bm[m] | short sigt[32][Kp] | long long wsum[NW][32]
constexpr int kTileSets = 32;
constexpr int kTileThreads = 128;
constexpr int kTileMaxElems = 32768;           // one 32-bit word per element: 128 KB at most; ids fit 15 bits
constexpr int kTileMaxSmemBytes = 227 * 1024;
constexpr bool kAutoUsesTile = true;   // AUTO takes the tile scan whenever the word map fits

__device__ __forceinline__ void st_shared_b16(unsigned addr, unsigned v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"((unsigned short)(v)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ unsigned ld_map(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// plane | (word & ~found) in one LOP3
__device__ __forceinline__ unsigned or_andn(unsigned plane, unsigned word, unsigned found) {
  unsigned d;
  asm("lop3.b32 %0, %1, %2, %3, 0xF4;" : "=r"(d) : "r"(plane), "r"(word), "r"(found));
  return d;
}

// signature-tile pitch in shorts: (pitch / 2) odd, so the 32 rows start in 32 different banks
__host__ __device__ constexpr int tile_sig_stride(int K) { return round_up(K, 32) + 2; }

// NH = number of high planes: ranks < 8 << NH
template <int NH>
__global__ void __launch_bounds__(kTileThreads, 4)
minhash_tile_kernel(const int64_t *__restrict__ rowptr, const int *__restrict__ colidx, int64_t n, int m,
                    const int *__restrict__ perm_tab, const int *__restrict__ order_head, int R, int K, int b,
                    int64_t id_base, int *__restrict__ sig, long long *__restrict__ norm2,
                    uint64_t *__restrict__ records, int64_t rec_stride, int mode, int map_words) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NW = kTileThreads / 32;
  constexpr int NP = 3 + NH;
  static_assert(NW * 8 == kTileSets, "warp w owns byte w of every map word");
  static_assert(NP <= 16, "two un-sliced ranks share one register");
  static_assert(NH >= 3, "the three lowest chunk digits are static");
  const int Kp = tile_sig_stride(K);
  long long *rp = reinterpret_cast<long long *>(smem_raw);
  unsigned *bm = reinterpret_cast<unsigned *>(rp + 40);   // 16-byte aligned: map_words % 4 == 0
  short *sigt = reinterpret_cast<short *>(bm + map_words);
  long long *wsum = reinterpret_cast<long long *>(sigt + kTileSets * Kp);   // 64 * Kp bytes: 8-byte aligned
  unsigned char *bmb = reinterpret_cast<unsigned char *>(bm);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r_band = b > 0 ? K / b : 0;
  const int64_t tiles = ceil_div<int64_t>(n, kTileSets);
  const int nchunks = R >> 3;
  const size_t chunk_stride = size_t(head_pitch(K)) << 1;          // int4 units per 8-rank chunk
  const unsigned sigt_s = unsigned(__cvta_generic_to_shared(sigt));
  const unsigned bm_s = unsigned(__cvta_generic_to_shared(bm));
  const unsigned Kp2 = unsigned(Kp) * 2u;

  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t s0 = tile * kTileSets;
    const int nsets = int(min(int64_t(kTileSets), n - s0));

    // ---- stage the row bounds, clear the map
    if (tid <= kTileSets) rp[tid] = rowptr[s0 + min(tid, nsets)];
    for (int t = tid * 4; t < map_words; t += kTileThreads * 4)
      *reinterpret_cast<uint4 *>(bm + t) = make_uint4(0, 0, 0, 0);
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
    for (int jj = 0; jj < 8; ++jj) {
      const int j = warp * 8 + jj;
      if (j < nsets && rp[j + 1] == rp[j])
        for (int i = lane; i < K; i += 32) sigt[j * Kp + i] = -1;
    }

    // ---- build: warp w adds its (up to) 8 sets to byte w of the words, one set after the other, 256 entries
    //      per step; the NEXT step's column indices are loaded before this step's bytes are updated
    {
      unsigned todo = (scanmask >> (warp * 8)) & 0xFFu;
      int jj = todo ? __ffs(todo) - 1 : -1;
      int t0 = 0;
      int e[8];
      auto load8 = [&](int (&dst)[8], int set, int first) {
        const int64_t lo = rp[warp * 8 + set];
        const int len = int(rp[warp * 8 + set + 1] - lo);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int t = first + lane + 32 * u;
          dst[u] = t < len ? ld_stream_s32(colidx + lo + t) : -1;
        }
      };
      if (jj >= 0) load8(e, jj, 0);
      while (jj >= 0) {
        const int len = int(rp[warp * 8 + jj + 1] - rp[warp * 8 + jj]);
        int njj = jj, nt0 = t0 + 256;
        if (nt0 >= len) {
          todo &= todo - 1;
          njj = todo ? __ffs(todo) - 1 : -1;
          nt0 = 0;
        }
        int en[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
        if (njj >= 0) load8(en, njj, nt0);
        const unsigned char bit = (unsigned char)(1u << jj);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (e[u] >= 0) {                 // distinct elements of one row: no two lanes share a byte
            unsigned char *p = bmb + e[u] * 4 + warp;
            *p = *p | bit;
          }
        if (njj != jj) __syncwarp();       // the next set may touch the bytes other lanes just wrote
#pragma unroll
        for (int u = 0; u < 8; ++u) e[u] = en[u];
        jj = njj;
        t0 = nt0;
      }
    }
    __syncthreads();

    // the next tile's column indices (one contiguous CSR span, read once) on their way into L2
    if (tile + gridDim.x < tiles) {
      const int64_t sn = (tile + gridDim.x) * kTileSets;
      const int64_t a = rowptr[sn], e = rowptr[min(sn + kTileSets, n)];
      const char *base = reinterpret_cast<const char *>(colidx + a);
      for (int64_t off = int64_t(tid) * 128; off < (e - a) * 4; off += kTileThreads * 128) prefetch_l2(base + off);
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
      // one 8-rank chunk, CC = its position inside a 64-rank group (compile time: the rank's six low digits are
      // static); true = stop (every pair of the warp resolved, or the head is exhausted).  Table entries are
      // byte offsets into the word map (4 * element).
      auto chunk = [&](auto cc_tag, int c, int4 &qa, int4 &qb) -> bool {
        constexpr int CC = decltype(cc_tag)::value;
        if (c >= nchunks || __all_sync(kFull, found == kFull)) return true;
        const unsigned w0 = ld_map(bm_s + unsigned(qa.x)), w1 = ld_map(bm_s + unsigned(qa.y));
        const unsigned w2 = ld_map(bm_s + unsigned(qa.z)), w3 = ld_map(bm_s + unsigned(qa.w));
        const unsigned w4 = ld_map(bm_s + unsigned(qb.x)), w5 = ld_map(bm_s + unsigned(qb.y));
        const unsigned w6 = ld_map(bm_s + unsigned(qb.z)), w7 = ld_map(bm_s + unsigned(qb.w));
        if (c + 4 < nchunks) {             // warp-uniform
          const int4 *nx = hp + size_t(c + 4) * chunk_stride;
          qa = __ldg(nx);
          qb = __ldg(nx + 1);
        }
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

      // ---- un-slice: sets j and j + 16 together (their bits of plane p go to bits p and p + 16)
      if (i < K) {
        const unsigned col_s = sigt_s + unsigned(i) * 2u;
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
            if (lo_on) st_shared_b16(col_s + unsigned(j) * Kp2, v2);
            if (hi_on) st_shared_b16(col_s + unsigned(j + 16) * Kp2, v2 >> 16);
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
          if (lane == 0) sigt[j * Kp + i0 + src] = short(v);
        }
      }
    }

    // ---- short sets (AUTO): direct gather-min, lane owns a hash, elements broadcast 32 at a time
    for (int jj = 0; jj < 8; ++jj) {
      const int j = warp * 8 + jj;
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
        if (i0 + lane < K) sigt[j * Kp + i0 + lane] = short(v);
      }
    }
    __syncthreads();

    // ---- epilogue 1: the tile's rows to global — consecutive sets are one contiguous block of 32 * K ints
    if ((K & 1) == 0) {
      const int halfK = K >> 1;
      for (int j = warp; j < nsets; j += NW) {
        int2 *dst = reinterpret_cast<int2 *>(sig + (s0 + j) * K);
        const unsigned *src = reinterpret_cast<const unsigned *>(sigt + j * Kp);     // Kp even: 4-byte aligned rows
        for (int q = lane; q < halfK; q += 32) {
          const unsigned x = src[q];
          dst[q] = make_int2(int(short(x & 0xFFFFu)), int(x) >> 16);
        }
      }
    } else {
      for (int j = warp; j < nsets; j += NW)
        for (int i = lane; i < K; i += 32) sig[(s0 + j) * K + i] = sigt[j * Kp + i];
    }
    // ---- epilogue 2: lane = set.  Bucket records of the warp's bands and, from the same values, its share of
    //      the squared norm; the rows' pitch (odd number of words) keeps the 32 lanes in 32 banks
    if (records != nullptr || norm2 != nullptr) {
      long long acc = 0;
      const short *row = sigt + lane * Kp;
      if (lane < nsets) {
        if (records != nullptr) {
          for (int band = warp; band < b; band += NW) {
            uint32_t h = band_hash_init(uint32_t(band));
            bool all_neg = true;
            for (int c = 0; c < r_band; ++c) {
              const int x = row[band * r_band + c];
              acc += (long long)x * x;
              const uint32_t l16 = uint32_t(x) & 0xFFFFu;
              all_neg = all_neg && (l16 == 0xFFFFu);
              h = band_hash_step(h, l16);
            }
            records[int64_t(band) * rec_stride + s0 + lane] =
                make_record(band_hash_finish(h, all_neg), uint32_t(band), uint32_t(id_base + s0 + lane));
          }
        } else {
          for (int i = warp; i < K; i += NW) {
            const int x = row[i];
            acc += (long long)x * x;
          }
        }
      }
      if (norm2 != nullptr) {
        wsum[warp * 32 + lane] = acc;
        __syncthreads();
        if (warp == 0 && lane < nsets) {
          long long tot = 0;
#pragma unroll
          for (int w = 0; w < NW; ++w) tot += wsum[w * 32 + lane];
          norm2[s0 + lane] = tot;
        }
      }
    }
    __syncthreads();   // rp / sigt / bm / wsum are rewritten by the next tile
  }
}

static size_t tile_smem_bytes(int K, int map_words) {
  return 40 * 8 + size_t(map_words) * 4 + size_t(kTileSets) * tile_sig_stride(K) * 2 +
         size_t(kTileThreads / 32) * 32 * 8;
}

template <int NH>
static int launch_minhash_tile_nh(const int64_t *rowptr, const int *colidx, int64_t n, int m, const int *perm_tab,
                                  const int *order_head, int R, int K, int b, int64_t id_base, int *sig,
                                  long long *norm2, uint64_t *records, int64_t rec_stride, int mode,
                                  cudaStream_t stream) {
  auto kern = minhash_tile_kernel<NH>;
  const int map_words = round_up(m, 4);
  const size_t smem = tile_smem_bytes(K, map_words);
  QRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  QRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, smem));
  if (per_sm < 1) return QRS_ERR_UNSUPPORTED;
  int64_t grid = int64_t(sm_count()) * per_sm;
  const int64_t tiles = ceil_div<int64_t>(n, kTileSets);
  if (grid > tiles) grid = tiles;
  kern<<<unsigned(grid), kTileThreads, smem, stream>>>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base,
                                                       sig, norm2, records, rec_stride, mode, map_words);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

static int launch_minhash_tile(const int64_t *rowptr, const int *colidx, int64_t n, int m, const int *perm_tab,
                               const int *order_head, int R, int K, int b, int64_t id_base, int *sig,
                               long long *norm2, uint64_t *records, int64_t rec_stride, int mode,
                               cudaStream_t stream) {
  if (R <= (8 << 6))
    return launch_minhash_tile_nh<6>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, norm2,
                                     records, rec_stride, mode, stream);
  if (R <= (8 << 9))
    return launch_minhash_tile_nh<9>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, norm2,
                                     records, rec_stride, mode, stream);
  return launch_minhash_tile_nh<12>(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, norm2,
                                    records, rec_stride, mode, stream);
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
  const bool tile_possible = R >= 32 && m <= kTileMaxElems &&
                             (reinterpret_cast<uintptr_t>(order_head) & 15) == 0 &&
                             tile_smem_bytes(K, round_up(m, 4)) <= size_t(kTileMaxSmemBytes);
  if (mode == QRS_MINHASH_TILE && !tile_possible) return QRS_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  long long *n2 = reinterpret_cast<long long *>(norm2);
  if (tile_possible && (mode == QRS_MINHASH_TILE || (mode == QRS_MINHASH_AUTO && kAutoUsesTile)))
    return launch_minhash_tile(rowptr, colidx, n, m, perm_tab, order_head, R, K, b, id_base, sig, n2, records,
                               rec_stride, mode, st);
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
