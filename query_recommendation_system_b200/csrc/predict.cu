// a10 + a11 — similarity-weighted rating prediction (recommender.py:36-47) and the hybrid
// rule of compute_scores (recommender.py:311-331).
//
// The reference visits every zero cell (i, j) and gathers column j at user i's neighbours.
// Here the loop is turned inside out: for user i the neighbours' CSR rows are streamed in
// table order and scattered into per-column fp64 accumulators (num += r*s, den += s), which
// reproduces, per cell, exactly the reference's left-to-right fp64 sums — cells no
// neighbour rated keep den == 0 and are "missed" (prediction 0) without being visited.
//
//   sparse form : one CTA per (user, column tile); accumulators for the tile live in shared
//                 memory (exact integers when the similarities are multiples of 0.001, fp64
//                 otherwise), neighbour row segments are located through a per-row tile index
//                 and loaded a group of neighbours ahead of the accumulation; only the columns
//                 a neighbour touched are swept, predicted and written.
//   dense form  : one CTA per user at reference scale, adds the query branch.
#include "common.cuh"

namespace qrs {

// Python evaluates these in IEEE double; so does the host compiler (recommender.py:30-34, 324-331).
constexpr double kQW = 0.6, kUW = 0.4, kDM = 60.0;
constexpr double kOnlyQ_mul = kQW + (kUW * 0.5), kOnlyQ_add = kDM * (kUW * 0.5);
constexpr double kOnlyU_mul = kUW + (kQW * 0.5), kOnlyU_add = kDM * (kQW * 0.5);

__device__ __forceinline__ int hybrid_rule(double q, double u) {
  if (u == 0.0 && q == 0.0) return 0;
  double x;
  if (u == 0.0) x = __dadd_rn(__dmul_rn(q, kOnlyQ_mul), kOnlyQ_add);
  else if (q == 0.0) x = __dadd_rn(__dmul_rn(u, kOnlyU_mul), kOnlyU_add);
  else x = __dadd_rn(__dmul_rn(q, kQW), __dmul_rn(u, kUW));
  return int(rint(x));
}

// ------------------------------------------------------------------ tile index
// tilepos[v*(nt+1) + t] = rowptr[v] + number of entries of row v with col < t*tile (t = 0..nt):
// the absolute CSR position where tile t of row v starts; tile t ends where tile t+1 starts.
__global__ void __launch_bounds__(256)
tile_index_kernel(const int64_t *__restrict__ rowptr, const int *__restrict__ colidx, int64_t U, int tile, int nt,
                  int64_t *__restrict__ tilepos) {
  const int64_t total = U * (nt + 1);
  for (int64_t x = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; x < total; x += int64_t(gridDim.x) * blockDim.x) {
    const int64_t v = x / (nt + 1);
    const int t = int(x - v * (nt + 1));
    const int64_t lo = rowptr[v];
    const int len = int(rowptr[v + 1] - lo);
    const int64_t bound = int64_t(t) * tile;
    int a = 0, e = len;   // first entry with col >= bound
    while (a < e) {
      const int mid = (a + e) >> 1;
      if (int64_t(__ldg(colidx + lo + mid)) < bound) a = mid + 1; else e = mid;
    }
    tilepos[x] = lo + a;
  }
}

// ------------------------------------------------------------------ sparse form
struct __align__(16) Acc { double num, den; };

// Two accumulator flavours.
//  F64  : {num, den} in fp64, updated in table order — the reference's arithmetic verbatim.
//  MILLI: similarities are multiples of 0.001 (np.around(., 3), recommender.py:204/270), so
//         N = sum r*milli, D = sum milli and the rater count are exact integers packed in one
//         u64 (r << 56 | N << 32 | cnt << 24 | D).  round(0.7*N/D + 18) is then decided in exact integer
//         arithmetic; it can differ from the reference's fp64 evaluation only when
//         (7N + 180D) / (10D) is EXACTLY half-way between two integers (the fp64 rounding noise
//         of ~1e-12 is far below the next possible distance 1/(20D) >= 7e-7), and exactly those
//         cells are re-evaluated with the reference's sequential fp64 arithmetic.
template <bool MILLI>
struct AccOf { using type = Acc; };
template <>
struct AccOf<true> { using type = unsigned long long; };

// CTA shape of the sparse predictor.  Shipped: 384 threads (12 warps: two entries per thread cover a 768-entry
// neighbour segment), two CTAs per SM.  Both are build-time parameters so that other shapes (e.g. 192 threads x 4
// CTAs over half-width column tiles) can be measured as separate libraries (QRS_B200_LIB); the host side derives
// its tile choice from qrs_predict_segment_span() / qrs_predict_max_tile(), never from these numbers.
#ifndef QRS_SP_THREADS
#define QRS_SP_THREADS 384
#endif
#ifndef QRS_SP_MIN_CTAS
#define QRS_SP_MIN_CTAS 2
#endif
constexpr int kSpThreads = QRS_SP_THREADS;
constexpr int kSpMinCtas = QRS_SP_MIN_CTAS;
static_assert(kSpThreads % 32 == 0 && kSpThreads >= 64 && kSpThreads <= 512, "meta setup uses 64 threads, w_cnt holds 16 warps");
constexpr int kSpWarpsCta = kSpThreads / 32;
constexpr int kSpAhead = 4;    // neighbours whose segment loads are in flight together (x2: double-buffered)
constexpr int kSpSpan = 2 * kSpThreads;   // CSR positions one CTA pass covers: two consecutive entries per thread

// columns each warp sweeps in the output passes; the per-warp lists live at [w * range, (w + 1) * range)
__host__ __device__ constexpr int sparse_warp_range(int tile) { return round_up(ceil_div(tile, kSpWarpsCta), 32); }

// shared memory of one CTA: acc[tile] | u16 jlist[warps*range] | metadata
__host__ __device__ constexpr size_t sparse_cta_bytes(int tile, bool milli) {
  return round_up<size_t>(size_t(tile) * (milli ? 8 : 16), 16) +
         round_up<size_t>(size_t(sparse_warp_range(tile)) * kSpWarpsCta * 2, 16) + 64 * 8 + 256 + 26 * 32 * 4 + 64 * 16;
}

// predicated 8-byte shared-memory load / store (cond != 0), by shared-window byte address
__device__ __forceinline__ void lds_u64_if(unsigned long long &x, unsigned addr, unsigned cond) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.u64 %0, [%1];\n\t}"
               : "=l"(x) : "r"(addr), "r"(cond) : "memory");
}
__device__ __forceinline__ void sts_u64_if(unsigned addr, unsigned long long x, unsigned cond) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u64 [%0], %1;\n\t}"
               :: "r"(addr), "l"(x), "r"(cond) : "memory");
}

constexpr int kTieBitWords = 26 * 32;   // single-rater half-way bits: 26 ratings (5, 15, .., 255) x 1024 milli

__device__ __forceinline__ int predict_from(double num, double den) {
  return den != 0.0 ? hybrid_rule(0.0, num / den) : 0;
}

// the reference's sequential fp64 evaluation of one cell, re-gathered from the neighbour segments
__device__ __noinline__ int resolve_tie(const int *__restrict__ colidx, const uint8_t *__restrict__ val,
                                        const int4 *meta4, const double *m_s, int c, int col) {
  double num = 0.0, den = 0.0;
  for (int t = 0; t < c; ++t) {
    const int a = meta4[t].y, n = meta4[t].z - a;
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(colidx + a + mid) < col) lo = mid + 1; else hi = mid;
    }
    if (lo < n && __ldg(colidx + a + lo) == col) {
      const int r = val[a + lo];
      if (r != 0) {
        num = __dadd_rn(num, __dmul_rn(double(r), m_s[t]));
        den = __dadd_rn(den, m_s[t]);
      }
    }
  }
  return predict_from(num, den);
}

// One CTA per (user, column tile).  The neighbours' segments are streamed one after the other (table
// order) by all threads — entries of one neighbour hit distinct columns, so the shared-memory
// read-modify-writes need no atomics, and a block barrier between neighbours keeps the reference's
// left-to-right order per column.  The next group of neighbours is being loaded into a second
// register set while the current group is accumulated.  The output sweep is per warp over its own
// column range: list the touched columns, reserve one slice for the CTA, predict into the slice.
template <bool MILLI>
__global__ void __launch_bounds__(kSpThreads, kSpMinCtas)
predict_sparse_kernel(const int *__restrict__ colidx, const uint8_t *__restrict__ val, int Q,
                      const int *__restrict__ u_idx, const double *__restrict__ u_val,
                      const int *__restrict__ u_milli, const int *__restrict__ u_cnt, int ku, int64_t user_base,
                      int64_t n_users, int tile, int nt, const int64_t *__restrict__ tilepos,
                      int64_t *__restrict__ row_start, int *__restrict__ row_cnt, int *__restrict__ cols,
                      uint8_t *__restrict__ preds, int64_t region_cap, int n_regions,
                      unsigned long long *__restrict__ cursors, unsigned long long *__restrict__ tie_count,
                      ulonglong2 *__restrict__ tie_list, int64_t tie_capacity, int cols_mode,
                      const unsigned *__restrict__ tie_bits_g) {
  using acc_t = typename AccOf<MILLI>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int range = sparse_warp_range(tile);
  acc_t *acc = reinterpret_cast<acc_t *>(smem_raw);
  unsigned char *ptr = smem_raw + round_up<size_t>(size_t(tile) * (MILLI ? 8 : 16), 16);
  uint16_t *jlist = reinterpret_cast<uint16_t *>(ptr);
  unsigned char *meta = ptr + round_up<size_t>(size_t(range) * kSpWarpsCta * 2, 16);
  double *m_s = reinterpret_cast<double *>(meta);                    // [64] similarity of each neighbour
  int *w_cnt = reinterpret_cast<int *>(meta + 64 * 8);              // [16] prediction bytes per warp
  int *w_off = w_cnt + 16;                                          // [16] exclusive prefix of w_cnt
  unsigned long long *s_base = reinterpret_cast<unsigned long long *>(w_cnt + 48);
  unsigned *tie_bits = reinterpret_cast<unsigned *>(w_cnt + 64);    // [26 * 32] MILLI: single-rater half-way bits
  // [64] per neighbour, one 16-byte broadcast load: {even-aligned first position, first position, end position,
  // MILLI increment}; positions are 32-bit (nnz < 2^31, checked by the host side)
  int4 *meta4 = reinterpret_cast<int4 *>(w_cnt + 64 + kTieBitWords);
  acc_t zero;
  if constexpr (MILLI) zero = 0ull; else zero = Acc{0.0, 0.0};
  for (int j = tid; j < tile; j += kSpThreads) acc[j] = zero;
  if constexpr (MILLI)
    for (int w = tid; w < kTieBitWords; w += kSpThreads) tie_bits[w] = tie_bits_g[w];
  if (tid < 64) meta4[tid] = make_int4(0, 0, 0, 0);
  __syncthreads();

  const unsigned lt = lanemask_lt();
  const unsigned acc_s = unsigned(__cvta_generic_to_shared(acc));   // shared-window byte address of acc[0]
  const int stride = nt + 1;
  // output slices are reserved from one of n_regions cursors (a single cursor would serialise
  // millions of same-address atomics in L2); region r owns [r * region_cap, (r + 1) * region_cap)
  const int region = int(blockIdx.x % n_regions);
  unsigned long long *cursor = cursors + region;
  const int64_t region_base = int64_t(region) * region_cap;

  // one rating r of a neighbour into column slot j (inc: MILLI increment milli | 1 << 24; g: neighbour slot)
  auto add = [&](int j, int r, unsigned inc, int g) {
    if constexpr (MILLI) {
      // u64 = r (8 bits, meaningful for a single rater) | N = sum r*milli (24) | raters (12) | D = sum milli (20)
      acc[j] += (static_cast<unsigned long long>(unsigned(r) * inc) << 32) | inc;
    } else {
      const double s = m_s[g];
      Acc x = acc[j];
      x.num = __dadd_rn(x.num, __dmul_rn(double(r), s));
      x.den = __dadd_rn(x.den, s);
      acc[j] = x;
    }
  };

  const int64_t items = n_users * nt;
  long long nz_total = 0;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t ul = item / nt;                 // local user
    const int t = int(item - ul * nt);
    const int64_t ug = user_base + ul;            // global user = CSR row
    const int q0 = t * tile;
    const int width = min(tile, Q - q0);
    const int c = min(u_cnt[ul], ku);
    const int c4 = round_up(c, kSpAhead);         // padded with empty neighbours: no bounds checks in the loop
    bool is_long = false;
    if (tid < c4) {
      int4 mt = make_int4(0, 0, 0, 0);
      if (tid < c) {
        const int v = u_idx[ul * ku + tid];
        const int64_t *tp = tilepos + int64_t(v) * stride + t;
        const int64_t a = __ldg(tp), e = __ldg(tp + 1);
        unsigned inc = 0;
        if constexpr (MILLI) {
          const int milli = u_milli[ul * ku + tid];
          m_s[tid] = double(milli) / 1000.0;        // == the table's double (u_val == u_milli / 1000.0)
          inc = unsigned(milli) | (1u << 24);
        } else {
          m_s[tid] = u_val[ul * ku + tid];
        }
        mt = make_int4(int(a) & ~1, int(a), int(e), int(inc));
        is_long = (int(a) & ~1) + kSpSpan < int(e);
      }
      meta4[tid] = mt;
    }
    const int64_t own_a = __ldg(tilepos + ug * stride + t), own_e = __ldg(tilepos + ug * stride + t + 1);
    const bool any_long = __syncthreads_or(is_long);

    // ---- neighbours in table order, kSpAhead at a time, two register sets in ping-pong: the NEXT group's
    //      entries are being loaded while the current group is accumulated.  A thread takes two consecutive
    //      CSR positions of a segment (one 8-byte column load + one 2-byte rating load from an even position).
    struct Group { int2 col[kSpAhead]; unsigned rat[kSpAhead]; unsigned inc[kSpAhead]; };
    auto load_group = [&](Group &G, int g0) {
#pragma unroll
      for (int g = 0; g < kSpAhead; ++g) {
        const int4 mt = meta4[g0 + g];
        const int p0 = mt.x + 2 * tid;
        G.inc[g] = unsigned(mt.w);
        G.rat[g] = 0u;
        G.col[g] = make_int2(0, 0);
        if (p0 < mt.z) {
          G.col[g] = *reinterpret_cast<const int2 *>(colidx + p0);
          unsigned rr = *reinterpret_cast<const unsigned short *>(val + p0);
          if (p0 < mt.y) rr &= 0xFF00u;                           // position before the segment
          if (p0 + 1 >= mt.z) rr &= 0x00FFu;                      // position after the segment
          G.rat[g] = rr;
        }
      }
    };
    auto process_group = [&](const Group &G, int g0) {
#pragma unroll
      for (int g = 0; g < kSpAhead; ++g) {
        if constexpr (MILLI) {
          // both read-modify-writes of the thread at once, predicated instead of branched (the two entries
          // of one neighbour are different columns).  inc = milli | 1 << 24 is the low word's increment (D and
          // the rater count) AND the high word's multiplier: r * inc = r << 24 | r * milli, as r * milli < 2^24
          const unsigned r0 = G.rat[g] & 0xFFu, r1 = G.rat[g] >> 8;
          const unsigned a0 = acc_s + unsigned(G.col[g].x - q0) * 8u, a1 = acc_s + unsigned(G.col[g].y - q0) * 8u;
          unsigned long long x0, x1;
          lds_u64_if(x0, a0, r0);
          lds_u64_if(x1, a1, r1);
          x0 += (static_cast<unsigned long long>(r0 * G.inc[g]) << 32) | G.inc[g];
          x1 += (static_cast<unsigned long long>(r1 * G.inc[g]) << 32) | G.inc[g];
          sts_u64_if(a0, x0, r0);
          sts_u64_if(a1, x1, r1);
        } else {
          if (G.rat[g] & 0xFFu) add(G.col[g].x - q0, int(G.rat[g] & 0xFFu), G.inc[g], g0 + g);
          if (G.rat[g] >> 8) add(G.col[g].y - q0, int(G.rat[g] >> 8), G.inc[g], g0 + g);
        }
        if (!MILLI && any_long) {                   // block-uniform; the rest of a long segment straight from memory
          const int4 mt = meta4[g0 + g];
#pragma unroll 1
          for (int x = mt.x + kSpSpan + tid; x < mt.z; x += kSpThreads) {
            const int r = val[x];
            if (r != 0) add(__ldg(colidx + x) - q0, r, unsigned(mt.w), g0 + g);
          }
        }
        __syncthreads();
      }
    };
    {
      Group A, B;
      if (c4 > 0) load_group(A, 0);
      for (int g0 = 0; g0 < c4; g0 += 2 * kSpAhead) {
        const bool more = g0 + kSpAhead < c4;
        if (more) load_group(B, g0 + kSpAhead);
        process_group(A, g0);
        if (more) {
          if (g0 + 2 * kSpAhead < c4) load_group(A, g0 + 2 * kSpAhead);
          process_group(B, g0 + kSpAhead);
        }
      }
    }

    // ---- MILLI: the sums are exact integers, so the order of the additions is free — the rest of the segments
    //      longer than one CTA pass is added here, neighbour by neighbour (a barrier between two neighbours: they
    //      may share columns), which keeps the test out of the loop above
    if (MILLI && any_long) {
      for (int g = 0; g < c; ++g) {
        const int4 mt = meta4[g];
#pragma unroll 1
        for (int x = mt.x + kSpSpan + tid; x < mt.z; x += kSpThreads) {
          const int r = val[x];
          if (r != 0) add(__ldg(colidx + x) - q0, r, unsigned(mt.w), g);
        }
        __syncthreads();
      }
    }

    // ---- the user's own rated cells are not predicted (recommender.py:302): drop what landed there
    for (int64_t x = own_a + tid; x < own_e; x += kSpThreads)
      if (val[x] != 0) acc[__ldg(colidx + x) - q0] = zero;
    __syncthreads();

    // ---- pass 1 (per warp, its own column range): ascending list of the columns that get a prediction byte.
    //      A lane looks at columns j0 + lane and j0 + 32 + lane, so the two ballots ARE the bitmap words.
    const bool bitmap = cols_mode == 2;
    const int j_lo = warp * range, j_hi = min(width, j_lo + range);
    uint16_t *wj = jlist + j_lo;
    unsigned *const bitmap_row = reinterpret_cast<unsigned *>(cols) + item * (tile >> 5);
    int touched = 0;
#pragma unroll 2
    for (int j0 = j_lo; j0 < j_hi; j0 += 64) {
      const int ja = j0 + lane, jb = ja + 32;
      bool hit0 = false, hit1 = false;
      if constexpr (MILLI) {
        // listed iff D != 0 (then the prediction is >= 18); an accumulator whose similarities are all 0 is
        // cleared right here and predicts nothing (weightSum == 0, recommender.py:44)
        if (ja < j_hi) {
          const unsigned long long x = acc[ja];
          hit0 = (unsigned(x) & 0xFFFFFFu) != 0u;
          if (!hit0 && x != 0ull) acc[ja] = 0ull;
        }
        if (jb < j_hi) {
          const unsigned long long x = acc[jb];
          hit1 = (unsigned(x) & 0xFFFFFFu) != 0u;
          if (!hit1 && x != 0ull) acc[jb] = 0ull;
        }
      } else {
        if (ja < j_hi) { const Acc x = acc[ja]; hit0 = x.den != 0.0 || x.num != 0.0; }
        if (jb < j_hi) { const Acc x = acc[jb]; hit1 = x.den != 0.0 || x.num != 0.0; }
      }
      const unsigned m0 = __ballot_sync(kFull, hit0), m1 = __ballot_sync(kFull, hit1);
      if (bitmap && lane < 2 && j0 + 32 * lane < j_hi) bitmap_row[(j0 >> 5) + lane] = lane == 0 ? m0 : m1;
      const int n0 = __popc(m0);
      if (hit0) wj[touched + __popc(m0 & lt)] = uint16_t(ja);
      if (hit1) wj[touched + n0 + __popc(m1 & lt)] = uint16_t(jb);
      touched += n0 + __popc(m1);
    }
    if (lane == 0) w_cnt[warp] = touched;
    __syncthreads();
    if (tid == 0) {
      int sum = 0;
#pragma unroll
      for (int w = 0; w < kSpWarpsCta; ++w) { w_off[w] = sum; sum += w_cnt[w]; }
      unsigned long long base = 0;
      if (sum > 0) base = atomicAdd(cursor, (unsigned long long)sum);
      row_start[item] = region_base + int64_t(base);
      row_cnt[item] = sum;
      *s_base = base;
    }
    __syncthreads();
    // ---- pass 2: one prediction byte per listed column, written straight into the reserved slice (ascending:
    //      warp ranges are ascending, lists are ascending); accumulators back to zero
    {
      // this warp's part of the slice: 64-bit bases once, 32-bit offsets per prediction
      const int64_t wbase = int64_t(*s_base) + w_off[warp];
      const int fit = int(max(int64_t(0), min(int64_t(touched), region_cap - wbase)));   // region overflow: dropped
      uint8_t *const w_preds = preds + (region_base + wbase);
      char *const w_cols = reinterpret_cast<char *>(cols) + (region_base + wbase) * (cols_mode == 1 ? 2 : 4);
      int nz = 0;
#pragma unroll 2
      for (int x0 = 0; x0 < touched; x0 += 32) {
        const int x = x0 + lane;
        int p = 0;
        if (x < touched) {
          const int j = wj[x];
          const acc_t a = acc[j];
          acc[j] = zero;
          bool tie = false;
          if constexpr (MILLI) {
            const unsigned lo = unsigned(a), hi = unsigned(a >> 32);
            const unsigned D = lo & 0xFFFFFFu, raters = lo >> 24;
            const unsigned N = hi & 0xFFFFFFu;
            const unsigned X = 7u * N + 180u * D, Y = 10u * D;    // x = X / Y = 0.7 * N / D + 18
            // q = X / Y via a float estimate (q < 256, estimate off by at most one) corrected exactly
            unsigned q = unsigned(__fdividef(float(X), float(Y)));
            int rem = int(X - q * Y);
            if (rem < 0) { --q; rem += int(Y); } else if (rem >= int(Y)) { ++q; rem -= int(Y); }
            const unsigned rem2 = 2u * unsigned(rem);
            p = int(q) + (rem2 > Y ? 1 : 0);
            if (rem2 == Y) {
              // exact half-way: the reference's fp64 rounding noise decides.  One rater (N = r * D, so
              // x = (7r + 180) / 10): the fp64 chain's bit from the table; several: the fix-up kernel
              if (raters == 1u) {
                const unsigned r1 = hi >> 24;
                p += int((tie_bits[(r1 / 10u) * 32u + (D >> 5)] >> (D & 31u)) & 1u);
              } else {
                tie = true;
              }
            }
          } else {
            p = predict_from(a.num, a.den);
          }
          if (x < fit) {
            if (MILLI && tie) {
              const unsigned long long slot = atomicAdd(tie_count, 1ull);
              if (int64_t(slot) < tie_capacity)
                tie_list[slot] = make_ulonglong2((unsigned long long)(w_preds + x - preds),   // position in preds
                                                 ((unsigned long long)(unsigned)ul << 32) | unsigned(q0 + j));
              else
                p = resolve_tie(colidx, val, meta4, m_s, c, q0 + j);   // list full: settle here
            }
            if (cols_mode == 1) reinterpret_cast<uint16_t *>(w_cols)[x] = uint16_t(q0 + j);
            else if (cols_mode == 0) reinterpret_cast<int *>(w_cols)[x] = q0 + j;
            w_preds[x] = uint8_t(p);
          }
        }
        nz += __popc(__ballot_sync(kFull, p != 0));
      }
      nz_total += nz;
    }
    __syncthreads();
  }
  // non-zero predictions of this launch (statistics): one atomic per warp
  if (lane == 0 && nz_total > 0) atomicAdd(cursors + n_regions, (unsigned long long)nz_total);
}

// ------------------------------------------------------------------ dense form
// smem: double num[Q] | double den[Q] | u8 own[Q]
__global__ void __launch_bounds__(256)
predict_dense_kernel(const int64_t *__restrict__ rowptr, const int *__restrict__ colidx,
                     const uint8_t *__restrict__ val, int Q, const int *__restrict__ q_idx,
                     const double *__restrict__ q_val, const int *__restrict__ q_cnt, int kq,
                     const int *__restrict__ u_idx, const double *__restrict__ u_val,
                     const int *__restrict__ u_cnt, int ku, int64_t user_base, int64_t n_users,
                     int *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *num = reinterpret_cast<double *>(smem_raw);
  double *den = num + Q;
  uint8_t *own = reinterpret_cast<uint8_t *>(den + Q);
  const int tid = threadIdx.x;
  for (int64_t ul = blockIdx.x; ul < n_users; ul += gridDim.x) {
    const int64_t ug = user_base + ul;
    for (int j = tid; j < Q; j += blockDim.x) { num[j] = 0.0; den[j] = 0.0; own[j] = 0; }
    __syncthreads();
    {
      const int64_t lo = rowptr[ug];
      const int len = int(rowptr[ug + 1] - lo);
      for (int x = tid; x < len; x += blockDim.x) own[colidx[lo + x]] = val[lo + x];
    }
    if (u_idx != nullptr) {
      const int c = min(u_cnt[ul], ku);
      for (int t = 0; t < c; ++t) {
        const int v = u_idx[ul * ku + t];
        const double s = u_val[ul * ku + t];
        const int64_t lo = rowptr[v];
        const int len = int(rowptr[v + 1] - lo);
        for (int x = tid; x < len; x += blockDim.x) {
          const int r = val[lo + x];
          if (r != 0) {
            const int j = colidx[lo + x];
            num[j] = __dadd_rn(num[j], __dmul_rn(double(r), s));
            den[j] = __dadd_rn(den[j], s);
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
    for (int j = tid; j < Q; j += blockDim.x) {
      int res = own[j];
      if (res == 0) {
        double q = 0.0;
        if (q_idx != nullptr) {
          const int c = min(q_cnt[j], kq);
          double qn = 0.0, qw = 0.0;
          for (int t = 0; t < c; ++t) {
            const int r = own[q_idx[int64_t(j) * kq + t]];
            const double s = q_val[int64_t(j) * kq + t];
            qn = __dadd_rn(qn, __dmul_rn(double(r), s));
            if (r != 0) qw = __dadd_rn(qw, s);
          }
          if (qw != 0.0) q = qn / qw;
        }
        const double u = den[j] != 0.0 ? num[j] / den[j] : 0.0;
        res = hybrid_rule(q, u);
      }
      out[ul * Q + j] = res;
    }
    __syncthreads();
  }
}

// One rater of rating r and similarity milli/1000: the exact value is x = (7r + 180) / 10, independent of milli,
// so round(x) is an integer expression unless x is half-way (r = 10t + 5); for those 26 ratings the
// reference's fp64 chain  u = (r*s)/s, round(u*0.7 + 18)  (recommender.py:36-47, 327) lands on the lower
// or the upper neighbour depending on milli — one BIT per (t, milli), 26 x 1024 bits, kept in shared memory.
__global__ void __launch_bounds__(256) single_rater_tie_bits_kernel(unsigned *__restrict__ bits) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;      // one 32-bit word per thread
  if (w >= kTieBitWords) return;
  const int t = w >> 5, r = 10 * t + 5;
  unsigned word = 0;
  for (int b = 0; b < 32; ++b) {
    const int milli = (w & 31) * 32 + b;
    if (milli >= 1 && milli <= 1000) {
      const double s = double(milli) / 1000.0;
      const int p = predict_from(__dmul_rn(double(r), s), s);
      if (p == (7 * r + 180) / 10 + 1) word |= 1u << b;
    }
  }
  bits[w] = word;
}

// Exact half-way cells with several raters (MILLI flavour): one warp per cell re-gathers the
// neighbours' ratings of that column (lane <-> neighbour, binary search in the neighbour's row) and
// replays the reference's sequential fp64 sums in table order.
__global__ void __launch_bounds__(256)
predict_tie_fixup_kernel(const int64_t *__restrict__ rowptr, const int *__restrict__ colidx,
                         const uint8_t *__restrict__ val, const int *__restrict__ u_idx,
                         const int *__restrict__ u_milli, const int *__restrict__ u_cnt, int ku,
                         const unsigned long long *__restrict__ tie_count, const ulonglong2 *__restrict__ tie_list,
                         int64_t tie_capacity, uint8_t *__restrict__ preds) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  int64_t n = int64_t(*tie_count);
  if (n > tie_capacity) n = tie_capacity;
  for (int64_t i = warp_global; i < n; i += nwarps) {
    const ulonglong2 rec = tie_list[i];
    const int64_t ul = int64_t(rec.y >> 32);
    const int col = int(unsigned(rec.y));
    const int c = min(u_cnt[ul], ku);
    double num = 0.0, den = 0.0;
    for (int t0 = 0; t0 < c; t0 += 32) {
      int r = 0;
      double s = 0.0;
      if (t0 + lane < c) {
        const int v = u_idx[ul * ku + t0 + lane];
        s = double(u_milli[ul * ku + t0 + lane]) / 1000.0;
        const int64_t a = rowptr[v];
        const int len = int(rowptr[v + 1] - a);
        int lo = 0, hi = len;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(colidx + a + mid) < col) lo = mid + 1; else hi = mid;
        }
        if (lo < len && __ldg(colidx + a + lo) == col) r = val[a + lo];
      }
      const int nn = min(32, c - t0);
      for (int t = 0; t < nn; ++t) {
        const int rt = __shfl_sync(kFull, r, t);
        const double st = __shfl_sync(kFull, s, t);
        if (rt != 0) {
          num = __dadd_rn(num, __dmul_rn(double(rt), st));
          den = __dadd_rn(den, st);
        }
      }
    }
    if (lane == 0) preds[rec.x] = uint8_t(predict_from(num, den));
  }
}

static size_t sparse_smem_bytes(int tile, bool milli) { return sparse_cta_bytes(tile, milli); }

template <bool MILLI>
static int launch_predict_sparse(const int *colidx, const uint8_t *val, int Q, const int *u_idx, const double *u_val,
                                 const int *u_milli, const int *u_cnt, int ku, int64_t user_base, int64_t n_users,
                                 int tile, int nt, const int64_t *tilepos, int64_t *row_start, int *row_cnt, int *cols,
                                 uint8_t *preds, int64_t region_cap, int n_regions, unsigned long long *cursors,
                                 unsigned long long *tie_count, ulonglong2 *tie_list, int64_t tie_capacity,
                                 int cols_mode, const unsigned *single_tab, cudaStream_t st) {
  auto kern = predict_sparse_kernel<MILLI>;
  const size_t smem = sparse_smem_bytes(tile, MILLI);
  QRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  QRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSpThreads, smem));
  if (per_sm < 1) return QRS_ERR_UNSUPPORTED;
  int64_t grid = int64_t(sm_count()) * per_sm;
  const int64_t need = n_users * nt;
  if (grid > need) grid = need;
  kern<<<unsigned(grid), kSpThreads, smem, st>>>(colidx, val, Q, u_idx, u_val, u_milli, u_cnt, ku, user_base,
                                                    n_users, tile, nt, tilepos, row_start, row_cnt, cols, preds,
                                                    region_cap, n_regions, cursors, tie_count, tie_list, tie_capacity,
                                                    cols_mode, single_tab);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // namespace qrs

using namespace qrs;

extern "C" {

/* largest column tile that still lets kSpMinCtas CTAs share an SM (two CTAs of <= ~112 KB of shared memory each as
 * shipped) */
int32_t qrs_predict_max_tile(int32_t integer_tables) {
  if (kSpThreads == 384 && kSpMinCtas == 2) return integer_tables ? 10016 : 5824;
  const size_t budget = size_t(228) * 1024 / kSpMinCtas - 1024;   // 1 KB per CTA is reserved by the system
  int tile = 32;
  while (sparse_cta_bytes(tile + 32, integer_tables != 0) <= budget) tile += 32;
  return tile;
}

/* CSR positions of one neighbour segment that one CTA pass covers (two per thread); longer segments take a slower
 * tail loop, so the host picks the column tile that keeps an average segment below ~0.9 of this */
int32_t qrs_predict_segment_span(void) { return kSpSpan; }

static int64_t tie_list_capacity(int64_t capacity) { return capacity / 32 + 65536; }
constexpr size_t kSingleTabBytes = round_up<size_t>(26 * 32 * 4, 256);

size_t qrs_predict_sparse_workspace_bytes(int64_t U, int32_t Q, int32_t tile, int64_t capacity) {
  if (U < 0 || Q <= 0 || tile <= 0 || capacity < 0) return 0;
  const int64_t nt = ceil_div<int64_t>(Q, tile);
  // tile index | tie counter | single-rater table | list of exact half-way cells awaiting their fp64 replay
  return round_up<size_t>(size_t(U > 0 ? U : 1) * size_t(nt + 1) * 8, 256) + 256 + kSingleTabBytes +
         round_up<size_t>(size_t(tie_list_capacity(capacity)) * 16, 256);
}

int qrs_predict_sparse(const int64_t *rowptr, const int32_t *colidx, const uint8_t *val, int64_t U, int32_t Q,
                       const int32_t *u_idx, const double *u_val, const int32_t *u_milli, const int32_t *u_cnt,
                       int32_t ku, int64_t user_base, int64_t n_users, int32_t tile, void *ws, size_t ws_bytes,
                       int64_t *row_start, int32_t *row_cnt, void *cols_any, int32_t cols_mode, uint8_t *preds,
                       int64_t capacity, int32_t n_regions, unsigned long long *cursors, void *stream) {
  int32_t *cols = static_cast<int32_t *>(cols_any);
  if (cols_mode < 0 || cols_mode > 2 || (cols_mode == 1 && Q > 65536) || (cols_mode == 2 && tile % 32 != 0))
    return QRS_ERR_INVALID_ARG;
  if (!rowptr || U < 0 || Q <= 0 || !u_idx || (!u_val && !u_milli) || !u_cnt || ku <= 0 || user_base < 0 ||
      n_users < 0 || user_base + n_users > U || !row_start || !row_cnt || !cursors || capacity < 0 || n_regions <= 0)
    return QRS_ERR_INVALID_ARG;
  const int64_t region_cap = capacity / n_regions;
  if (capacity > 0 && (!cols || !preds)) return QRS_ERR_INVALID_ARG;
  if (tile <= 0 || tile % 32 != 0) return QRS_ERR_INVALID_ARG;
  if (tile > qrs_predict_max_tile(u_milli != nullptr) || ku > QRS_MAX_TOPK) return QRS_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < qrs_predict_sparse_workspace_bytes(U, Q, tile, capacity)) return QRS_ERR_WORKSPACE;
  if (n_users == 0 || U == 0) return QRS_OK;
  if (!colidx || !val) return QRS_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  const int nt = int(ceil_div<int64_t>(Q, tile));
  int64_t *tilepos = static_cast<int64_t *>(ws);
  char *wsp = static_cast<char *>(ws) + round_up<size_t>(size_t(U) * size_t(nt + 1) * 8, 256);
  unsigned long long *tie_count = reinterpret_cast<unsigned long long *>(wsp);
  unsigned *single_tab = reinterpret_cast<unsigned *>(wsp + 256);
  ulonglong2 *tie_list = reinterpret_cast<ulonglong2 *>(wsp + 256 + kSingleTabBytes);
  const int64_t tie_cap = tie_list_capacity(capacity);
  QRS_CUDA(cudaMemsetAsync(tie_count, 0, 8, st));
  if (u_milli) {
    single_rater_tie_bits_kernel<<<ceil_div(26 * 32, 256), 256, 0, st>>>(single_tab);
    QRS_LAUNCH_CHECK();
  }
  {
    const int64_t total = U * (nt + 1);
    int64_t blocks = ceil_div<int64_t>(total, 256);
    const int64_t cap = int64_t(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    tile_index_kernel<<<unsigned(blocks), 256, 0, st>>>(rowptr, colidx, U, tile, nt, tilepos);
    QRS_LAUNCH_CHECK();
  }
  if (u_milli) {
    int rc = launch_predict_sparse<true>(colidx, val, Q, u_idx, u_val, u_milli, u_cnt, ku, user_base, n_users, tile,
                                         nt, tilepos, row_start, row_cnt, cols, preds, region_cap, n_regions, cursors,
                                         tie_count, tie_list, tie_cap, cols_mode, single_tab, st);
    if (rc != QRS_OK || capacity == 0) return rc;
    predict_tie_fixup_kernel<<<sm_count() * 8, 256, 0, st>>>(rowptr, colidx, val, u_idx, u_milli, u_cnt, ku, tie_count,
                                                             tie_list, tie_cap, preds);
    QRS_LAUNCH_CHECK();
    return QRS_OK;
  }
  return launch_predict_sparse<false>(colidx, val, Q, u_idx, u_val, u_milli, u_cnt, ku, user_base, n_users, tile, nt,
                                      tilepos, row_start, row_cnt, cols, preds, region_cap, n_regions, cursors,
                                      tie_count, tie_list, tie_cap, cols_mode, single_tab, st);
}

int qrs_predict_dense(const int64_t *rowptr, const int32_t *colidx, const uint8_t *val, int64_t U, int32_t Q,
                      const int32_t *q_idx, const double *q_val, const int32_t *q_cnt, int32_t kq,
                      const int32_t *u_idx, const double *u_val, const int32_t *u_cnt, int32_t ku,
                      int64_t user_base, int64_t n_users, int32_t *out, void *stream) {
  if (!rowptr || U < 0 || Q <= 0 || user_base < 0 || n_users < 0 || user_base + n_users > U || !out)
    return QRS_ERR_INVALID_ARG;
  if (q_idx && (!q_val || !q_cnt || kq <= 0)) return QRS_ERR_INVALID_ARG;
  if (u_idx && (!u_val || !u_cnt || ku <= 0)) return QRS_ERR_INVALID_ARG;
  const size_t smem = size_t(Q) * 17;
  if (smem > 200 * 1024) return QRS_ERR_UNSUPPORTED;
  if (n_users == 0) return QRS_OK;
  if (!colidx || !val) return QRS_ERR_INVALID_ARG;
  QRS_CUDA(cudaFuncSetAttribute(predict_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  QRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, predict_dense_kernel, 256, smem));
  if (per_sm < 1) return QRS_ERR_UNSUPPORTED;
  int64_t grid = int64_t(sm_count()) * per_sm;
  if (grid > n_users) grid = n_users;
  predict_dense_kernel<<<unsigned(grid), 256, smem, as_stream(stream)>>>(
      rowptr, colidx, val, Q, q_idx, q_val, q_cnt, kq, u_idx, u_val, u_cnt, ku, user_base, n_users, out);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // extern "C"
