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

// ---------------------------------------------------------------------------------------------
// Tile scan (QRS_MINHASH_TILE): the same first-touch walk for 32 consecutive sets at once.
//
// Shared memory holds one 32-bit word per element: bit j of word e says "element e belongs
// to set s0 + j".  A lane owns one hash and walks its permutation in rank order; each rank
// costs ONE word load that answers the membership question for all 32 sets, and the sets
// that meet their first element at this rank get the rank as their signature value.  The walk
// ends when every (set, hash) of the warp is resolved: about ln(1024) * m / |set| ranks.  Measured
// at 100k x 10k, 5 %, K = 128: ~10 warp instructions per (set, hash) instead of ~43 (0.25 ms
// against 0.68 ms for the per-set scan above).
//   * the word map is built without atomics: warp w owns the sets 8w .. 8w+7, i.e. byte w of
//     every word, and adds its sets one after the other with byte read-modify-writes;
//   * order_head is hash-major (32 ranks of one hash = one 128-byte line), the walk wants
//     lane = hash: every warp stages 16 ranks x 32 hashes through a padded shared tile;
//   * a hit is recorded per bit (the set's signature value = the rank), except that a lane which
//     meets many of its sets in one step hands them to the warp: lane j writes its own bit's rank;
//   * short sets (len*len < m, AUTO only) are left out of the map and take the direct
//     gather-min; a (set, hash) still open after R ranks takes direct_min_one_hash.
//   * everything staged in shared memory is 16 bits wide (m <= 32768 bounds element ids, ranks and
//     signature values), which is what lets four CTAs share an SM at the bench shape.
//   dynamic smem: long long rp[40] | unsigned bm[m] | short sigt[32][Kp] | u16 xpose[NW][32][24]
constexpr int kTileSets = 32;
constexpr int kTileThreads = 128;
constexpr int kTileMaxElems = 32768;           // one 32-bit word per element: 128 KB at most; ids fit 15 bits
constexpr int kTileMaxSmemBytes = 227 * 1024;
constexpr bool kAutoUsesTile = true;   // AUTO takes the tile scan whenever the word map fits
constexpr int kTileCoopBits = 6;       // new sets of one lane in one 8-rank step from which the warp records them
                                       // (measured 3 / 6 / 12: 0.262 / 0.246 / 0.247 ms)
constexpr int kTileChunk = 16;         // ranks staged per round
constexpr int kTilePitch = 24;         // u16 per staged hash row: 48-byte rows keep the 16-byte reads conflict-free

__device__ __forceinline__ unsigned bfind(unsigned x) {   // index of the most significant set bit
  unsigned r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
__device__ __forceinline__ void st_shared_s16(unsigned addr, int v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(short(v)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// signature-tile pitch in shorts: (pitch / 2) odd, so the 32 rows start in 32 different banks
__host__ __device__ constexpr int tile_sig_stride(int K) { return round_up(K, 32) + 2; }

__global__ void __launch_bounds__(kTileThreads)
minhash_tile_kernel(const int64_t *__restrict__ rowptr, const int *__restrict__ colidx, int64_t n, int m,
                    const int *__restrict__ perm_tab, const int *__restrict__ order_head, int R, int K, int b,
                    int64_t id_base, int *__restrict__ sig, long long *__restrict__ norm2,
                    uint64_t *__restrict__ records, int64_t rec_stride, int mode, int map_words) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NW = kTileThreads / 32;
  static_assert(NW * 8 == kTileSets, "warp w owns byte w of every map word");
  const int Kp = tile_sig_stride(K);
  long long *rp = reinterpret_cast<long long *>(smem_raw);
  unsigned *bm = reinterpret_cast<unsigned *>(rp + 40);   // 16-byte aligned: map_words % 4 == 0
  short *sigt = reinterpret_cast<short *>(bm + map_words);
  unsigned short *xpose = reinterpret_cast<unsigned short *>(sigt + kTileSets * Kp);   // Kp % 8 == 2: see tile_smem_bytes
  unsigned char *bmb = reinterpret_cast<unsigned char *>(bm);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r_band = b > 0 ? K / b : 0;
  const int64_t tiles = ceil_div<int64_t>(n, kTileSets);
  unsigned short *xp = xpose + warp * (32 * kTilePitch);

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

    // ---- walk: lane = hash, 16 ranks staged per round
    for (int i0 = warp * 32; i0 < K; i0 += NW * 32) {
      const int i = i0 + lane;
      unsigned found = (i < K) ? ~scanmask : kFull;
      // column i of the signature tile as a shared-window byte address (only lanes with i < K write)
      const unsigned srow_s = unsigned(__cvta_generic_to_shared(sigt + i));
      const unsigned Kp2 = unsigned(Kp) * 2u;
      for (int r0 = 0; r0 < R; r0 += kTileChunk) {
        if (__all_sync(kFull, found == kFull)) break;
        // 16-byte loads: 4 lanes cover the 16 ranks of one hash, a warp load covers 8 hashes; ids packed to u16
#pragma unroll
        for (int s8 = 0; s8 < 32; s8 += 8) {
          const int hh = s8 + (lane >> 2), rq4 = (lane & 3) * 4;
          const int ih = min(i0 + hh, K - 1);
          const int4 v = __ldg(reinterpret_cast<const int4 *>(order_head + int64_t(ih) * R + r0 + rq4));
          *reinterpret_cast<uint2 *>(xp + hh * kTilePitch + rq4) =
              make_uint2(unsigned(v.x) | (unsigned(v.y) << 16), unsigned(v.z) | (unsigned(v.w) << 16));
        }
        __syncwarp();
        // 8 ranks per step; the second step's words are loaded before the first step's hits are recorded
        unsigned w[8], wn[8];
        {
          const uint4 q = *reinterpret_cast<const uint4 *>(xp + lane * kTilePitch);
          w[0] = bm[q.x & 0xFFFFu]; w[1] = bm[q.x >> 16]; w[2] = bm[q.y & 0xFFFFu]; w[3] = bm[q.y >> 16];
          w[4] = bm[q.z & 0xFFFFu]; w[5] = bm[q.z >> 16]; w[6] = bm[q.w & 0xFFFFu]; w[7] = bm[q.w >> 16];
        }
#pragma unroll
        for (int rq = 0; rq < kTileChunk; rq += 8) {
          if (rq == 0) {
            const uint4 q = *reinterpret_cast<const uint4 *>(xp + lane * kTilePitch + 8);
            wn[0] = bm[q.x & 0xFFFFu]; wn[1] = bm[q.x >> 16]; wn[2] = bm[q.y & 0xFFFFu]; wn[3] = bm[q.y >> 16];
            wn[4] = bm[q.z & 0xFFFFu]; wn[5] = bm[q.z >> 16]; wn[6] = bm[q.w & 0xFFFFu]; wn[7] = bm[q.w >> 16];
          }
          unsigned nw[8], all_new = 0;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            nw[u] = w[u] & ~found;
            found |= w[u];
            all_new |= nw[u];
          }
          // a lane that meets many of its sets in this step (a shared element of a taste group) hands the
          // recording to the warp: lane j looks for its own bit and writes one value, no per-bit loop
          unsigned big = __ballot_sync(kFull, __popc(all_new) >= kTileCoopBits);
          if (big) {
            if ((big >> lane) & 1u) all_new = 0;
            do {
              const int src = __ffs(big) - 1;
              big &= big - 1;
              int rk = -1;
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if ((__shfl_sync(kFull, nw[u], src) >> lane) & 1u) rk = u;
              if (rk >= 0) sigt[lane * Kp + i0 + src] = short(r0 + rq + rk);
            } while (big);
          }
          if (__any_sync(kFull, all_new != 0)) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              unsigned x = nw[u] & all_new;
              while (x) {
                const unsigned j = bfind(x);
                x ^= 1u << j;
                st_shared_s16(srow_s + j * Kp2, r0 + rq + u);
              }
            }
          }
          if (__all_sync(kFull, found == kFull)) break;
#pragma unroll
          for (int u = 0; u < 8; ++u) w[u] = wn[u];
        }
        __syncwarp();   // xp is rewritten by the next round
      }
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

    // ---- epilogue: rows to global (+ norm2), bucket records with lane = set
    for (int j = warp; j < nsets; j += NW) {
      long long acc = 0;
      for (int i = lane; i < K; i += 32) {
        const int x = sigt[j * Kp + i];
        sig[(s0 + j) * K + i] = x;
        acc += (long long)x * x;
      }
      if (norm2 != nullptr) {
        acc = warp_sum_ll(acc);
        if (lane == 0) norm2[s0 + j] = acc;
      }
    }
    if (records != nullptr) {
      for (int idx = tid; idx < b * kTileSets; idx += kTileThreads) {
        const int band = idx >> 5, j = idx & 31;
        if (j < nsets)
          records[int64_t(band) * rec_stride + s0 + j] =
              band_record(sigt + j * Kp + band * r_band, r_band, uint32_t(band), uint32_t(id_base + s0 + j));
      }
    }
    __syncthreads();   // rp / sigt / bm are rewritten by the next tile
  }
}

static size_t tile_smem_bytes(int K, int map_words) {
  // sigt: 32 * Kp shorts with Kp % 8 == 2 -> 64 * Kp bytes, a multiple of 16: xpose stays 16-byte aligned
  return 40 * 8 + size_t(map_words) * 4 + size_t(kTileSets) * tile_sig_stride(K) * 2 +
         size_t(kTileThreads / 32) * 32 * kTilePitch * 2;
}

static int launch_minhash_tile(const int64_t *rowptr, const int *colidx, int64_t n, int m, const int *perm_tab,
                               const int *order_head, int R, int K, int b, int64_t id_base, int *sig,
                               long long *norm2, uint64_t *records, int64_t rec_stride, int mode,
                               cudaStream_t stream) {
  const int map_words = round_up(m, 4);
  const size_t smem = tile_smem_bytes(K, map_words);
  QRS_CUDA(cudaFuncSetAttribute(minhash_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  QRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, minhash_tile_kernel, kTileThreads, smem));
  if (per_sm < 1) return QRS_ERR_UNSUPPORTED;
  int64_t grid = int64_t(sm_count()) * per_sm;
  const int64_t tiles = ceil_div<int64_t>(n, kTileSets);
  if (grid > tiles) grid = tiles;
  minhash_tile_kernel<<<unsigned(grid), kTileThreads, smem, stream>>>(rowptr, colidx, n, m, perm_tab, order_head, R,
                                                                      K, b, id_base, sig, norm2, records, rec_stride,
                                                                      mode, map_words);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
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
