// a10 + a11 — similarity-weighted rating prediction (recommender.py:36-47) and the hybrid
// rule of compute_scores (recommender.py:311-331).
//
// The reference visits every zero cell (i, j) and gathers column j at user i's neighbours.
// Here the loop is turned inside out: for user i the neighbours' CSR rows are streamed in
// table order and scattered into per-column fp64 accumulators (num += r*s, den += s), which
// reproduces, per cell, exactly the reference's left-to-right fp64 sums — cells no
// neighbour rated keep den == 0 and are "missed" (prediction 0) without being visited.
//
//   sparse form : one WARP per (user, column window); accumulators for the window live in the
//                 warp's slice of shared memory (exact integers when the similarities are
//                 multiples of 0.001, fp64 otherwise), neighbour rows are read from a packed,
//                 window-major copy of the ratings built once per CSR (qrs_predict_prepare),
//                 a few steps ahead of the accumulation; only the columns a neighbour touched
//                 are predicted and written.  No block barrier in the loop.
//   dense form  : one CTA per user at reference scale, adds the query branch.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

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

// ------------------------------------------------------------------ row preparation (once per ratings CSR)
// The predictor streams, for user i and column window t, the entries of every neighbour row that fall into the
// window.  qrs_predict_prepare lays the rows out for exactly that access: row v is cut at the window bounds
// (binary searches, rows are ascending) and every (row, window) sub-segment is written as 32-bit words
//   (column - t * tile) << 8 | rating
// padded to a multiple of 32 with words of rating 0 whose column is the spare accumulator slot (pw_spare_col), so
// that one warp step is ONE aligned 128-byte line, needs no bounds
// or alignment handling, and costs one load per entry instead of a column and a rating load.  seg_off[v * nt + t]
// is the sub-segment's first step (units of 32 words); sub-segments are consecutive in (v, t) order, so the next
// entry is its end.  The caller's colidx / val need no alignment or padding.

// Accumulator slots of one warp: the window's columns rounded up to the 64-column sweep step, then one spare slot
// (padded to 128 bytes) that the padding words of the prepared rows point at — a padding lane read-modify-writes
// the spare slot like any other, so the integer flavour's accumulation carries no predicate.
__host__ __device__ constexpr int pw_spare_col(int tile) { return round_up(tile, 64); }
__host__ __device__ constexpr int pw_acc_slots(int tile) { return round_up(tile, 64) + 16; }

// tilepos[v*(nt+1) + t] = rowptr[v] + number of entries of row v with col < t*tile (t = 0..nt)
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

__global__ void __launch_bounds__(256)
seg_steps_kernel(const int64_t *__restrict__ tilepos, int64_t U, int nt, uint32_t *__restrict__ steps) {
  const int64_t total = U * nt;
  for (int64_t x = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; x < total; x += int64_t(gridDim.x) * blockDim.x) {
    const int64_t v = x / nt;
    const int64_t *tp = tilepos + v * (nt + 1) + (x - v * nt);
    steps[x] = uint32_t((tp[1] - tp[0] + 31) >> 5);
  }
}

// One warp per ROW; one warp iteration = one 128-byte output step.  The row's window bounds (tilepos) and step
// offsets (seg_off) are held one per lane, 31 windows at a time, so the window of output step q is a ballot + popc
// and its bounds are three shuffles: no dependent global load per sub-segment, every iteration independent of the
// others (the loads of several steps are in flight together), loads and stores coalesced.
__global__ void __launch_bounds__(256, 5)
pack_rows_kernel(const int *__restrict__ colidx, const uint8_t *__restrict__ val, const int64_t *__restrict__ tilepos,
                 const uint64_t *__restrict__ seg_off, int64_t U, int nt, int tile, uint32_t *__restrict__ packed,
                 unsigned *__restrict__ zero_flag) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const uint32_t pad = uint32_t(pw_spare_col(tile)) << 8;      // rating 0, column = the spare accumulator slot
  bool saw_zero = false;
  for (int64_t v = warp_global; v < U; v += nwarps) {
    const int64_t *tp = tilepos + v * (nt + 1);
    const uint64_t *so = seg_off + v * nt;
    const int64_t row_lo = tp[0];                   // = rowptr[v]: positions inside the row fit 32 bits
    const int *const row_col = colidx + row_lo;
    const uint8_t *const row_val = val + row_lo;
    for (int tb = 0; tb < nt; tb += 31) {
      // lane l: bounds of window tb + l (lanes beyond the row's last bound repeat it: empty windows)
      const int tl = min(tb + lane, nt);
      const unsigned my_tp = unsigned(tp[tl] - row_lo);
      const uint64_t my_so = so[tl];              // so[nt] = the next row's first step (seg_off has U * nt + 1 entries)
      const int nwin = min(31, nt - tb);
      const uint64_t q_first = __shfl_sync(kFull, my_so, 0);
      const unsigned rel = unsigned(my_so - q_first);
      const unsigned q_count = __shfl_sync(kFull, rel, nwin);
      uint32_t *const out = packed + (q_first << 5) + lane;
#pragma unroll 8
      for (unsigned q = 0; q < q_count; ++q) {
        // window of step q: the last one (of this chunk) whose first step is <= q
        const int w = __popc(__ballot_sync(kFull, rel <= q && lane <= nwin)) - 1;
        const unsigned a = __shfl_sync(kFull, my_tp, w), e = __shfl_sync(kFull, my_tp, w + 1);
        const unsigned i = a + (q - __shfl_sync(kFull, rel, w)) * 32u + unsigned(lane);
        uint32_t word = pad;
        if (i < e) {
          const uint32_t r = row_val[i];
          word = (uint32_t(ld_stream_s32(row_col + i) - (tb + w) * tile) << 8) | r;
          saw_zero |= r == 0;              // an explicit rating 0 (= not rated, recommender.py:41): the predictor
        }                                  // takes its guarded accumulation path
        out[size_t(q) << 5] = word;
      }
    }
  }
  if (saw_zero) *zero_flag = 1u;
}

// ------------------------------------------------------------------ bank-class layout of a sub-segment
// The predictor's 8-byte read-modify-writes are served half a warp at a time, and two lanes of a half-warp whose
// columns agree mod 16 (same pair of banks, different address) cost an extra shared-memory wavefront each: with the
// entries of a sub-segment in column order a half-warp access takes 2.6 wavefronts instead of 1 (ncu, and the
// balls-in-bins model of 16 random columns), which is what bounds the kernel.  The order of the entries INSIDE one
// neighbour's sub-segment is free (distinct columns), so a sub-segment of n <= kDealMax entries is laid out by bank
// class: its steps * 32 words are H = 2 * steps half-steps of 16 lanes, and the i-th entry of
// class c = column mod 16 sits in half-step i at lane c (i = the order the class counter handed out, not the
// column order) — every half-step then touches 16 different bank pairs.
// Classes with more than H entries put the rest ("extras") into cells left free by shorter classes, last half-step
// first; an extra shares its half-step with an entry of its own class (one more wavefront there).  Padding words in
// lane l point at spare accumulator slot l mod 16, which no entry of that half-step can share a bank with.
// Mean wavefronts per half-step access: 2.61 -> 1.67 at C2 (50 entries), 1.71 -> 1.19 at C3 (10 entries).
// Longer sub-segments keep the column order.
constexpr int kDealMax = 128;
__device__ __forceinline__ unsigned deal_cell(unsigned half_step, unsigned lane16) {
  return (half_step >> 1) * 32u + (half_step & 1u) * 16u + lane16;
}

// One window by itself, whatever its class sizes (the careful path): the entries take their rank inside their class
// from a shared-memory counter, are scattered into a staging area at their cells and leave as whole 128-byte steps;
// extras park their words in a list and the free cells, last step first, pick them up.  scratch = 2 * kDealMax + 32
// words of the warp.  c0 / v0 point at this lane's first entry of the window.
__device__ __noinline__ void deal_window(const int *c0, const uint8_t *v0, unsigned n, int q0, uint32_t pad,
                                         uint32_t *out, uint32_t *scratch) {
  const int lane = threadIdx.x & 31;
  uint32_t *const stage = scratch, *const extra = scratch + kDealMax, *const count = scratch + 2 * kDealMax;
  const unsigned steps = (n + 31u) >> 5, H = 2u * steps;
  count[lane] = 0u;       // 16 class counters | extras parked | free cells handed out
  uint32_t word[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    word[j] = pad;
    if (unsigned(j) < steps) {                                   // warp-uniform
      if (unsigned(j * 32 + lane) < n) word[j] = (uint32_t(ld_stream_s32(c0 + j * 32) - q0) << 8) | v0[j * 32];
      stage[j * 32 + lane] = pad;
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (unsigned(j) < steps && word[j] != pad) {                 // (a real entry's column is inside the window)
      const unsigned cls = (word[j] >> 8) & 15u;
      const unsigned rank = atomicAdd(count + cls, 1u);
      if (rank < H) stage[deal_cell(rank, cls)] = word[j];
      else extra[atomicAdd(count + 16, 1u)] = word[j];
    }
  }
  __syncwarp();
  const unsigned n_extra = count[16];
  if (n_extra > 0u) {                                            // warp-uniform
    const unsigned cnt = count[lane & 15];
    for (int s2 = int(steps) - 1; s2 >= 0; --s2) {
      if (cnt <= 2u * unsigned(s2) + unsigned(lane >> 4)) {      // this lane's cell of step s2 is free
        const unsigned idx = atomicAdd(count + 17, 1u);
        if (idx < n_extra) stage[s2 * 32 + lane] = extra[idx];
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (unsigned(j) < steps) out[size_t(j) << 5] = stage[j * 32 + lane];
  __syncwarp();
}

// One warp per ROW, kDealGroup windows at a time.  Pass 1 runs over the group's entries, 32 per iteration and every
// iteration independent of the others: an entry finds its window, takes its rank from the (window, class) counter
// (shared-memory atomic: lanes of one class are ranked in whatever order the unit serves them — any order inside a
// class is as good) and is stored straight into its cell, or parked in its window's short list of extras.  Pass 2
// visits the cells the counters leave free: the window's parked extras first, padding words for the rest.  (Per
// window — rank, stage, store as whole steps — the same work is a chain of dependent latencies for ~50 entries:
// 0.33 ms for C2's rows against 0.17 ms for the plain column order; profiles/r2_predict_variants.md.)
constexpr int kDealGroup = 16;      // windows of a row handled together
constexpr int kDealExtras = 16;     // parked extras per window; a window with more goes the careful way
#ifndef QRS_PACK_CTAS
#define QRS_PACK_CTAS 5
#endif
__global__ void __launch_bounds__(256, QRS_PACK_CTAS)
pack_rows_dealt_kernel(const int *__restrict__ colidx, const uint8_t *__restrict__ val,
                       const int64_t *__restrict__ tilepos, const uint64_t *__restrict__ seg_off, int64_t U, int nt,
                       int tile, uint32_t *__restrict__ packed, unsigned *__restrict__ zero_flag) {
  __shared__ uint32_t s_cnt[8][kDealGroup * 16];                 // entries per (window, class)
  __shared__ uint32_t s_tab[8][4 * kDealGroup];                  // per window: H | first step | extras parked | free cells served
  __shared__ uint32_t s_park[8][kDealGroup * kDealExtras];
  __shared__ uint32_t s_scratch[8][2 * kDealMax + 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const uint32_t pad = uint32_t(pw_spare_col(tile) + (lane & 15)) << 8;      // rating 0, column = this lane's spare slot
  uint32_t *const cnt = s_cnt[warp];
  uint32_t *const tab_h = s_tab[warp], *const tab_so = tab_h + kDealGroup, *const tab_x = tab_so + kDealGroup,
                  *const tab_f = tab_x + kDealGroup;
  uint32_t *const park = s_park[warp];
  const float inv_tile = 1.0f / float(tile);
  uint32_t least = 255u;          // smallest rating seen: 0 = an explicit "not rated" entry (recommender.py:41)
  for (int64_t v = warp_global; v < U; v += nwarps) {
    const int64_t *tp = tilepos + v * (nt + 1);
    const uint64_t *so = seg_off + v * nt;
    const int64_t row_lo = tp[0];
    const int *const row_col = colidx + row_lo + lane;
    const uint8_t *const row_val = val + row_lo + lane;
    for (int tb = 0; tb < nt; tb += kDealGroup) {
      const int nwin = min(kDealGroup, nt - tb);
      const int tl = min(tb + lane, nt);
      const unsigned my_tp = unsigned(tp[tl] - row_lo);
      const uint64_t my_so = so[tl];              // so[nt] = the next row's first step (seg_off has U * nt + 1 entries)
      const unsigned g_lo = __shfl_sync(kFull, my_tp, 0), g_hi = __shfl_sync(kFull, my_tp, nwin);
      const unsigned my_n = __shfl_down_sync(kFull, my_tp, 1) - my_tp;
      if (g_lo == g_hi) continue;                                // warp-uniform: nothing in these windows
      const uint64_t g_so = __shfl_sync(kFull, my_so, 0);
      uint32_t *const g_out = packed + (g_so << 5);
      if (lane < nwin) {
        tab_h[lane] = (my_n == 0u || my_n > unsigned(kDealMax)) ? 0u : 2u * ((my_n + 31u) >> 5);
        tab_so[lane] = unsigned(my_so - g_so) << 5;
        tab_x[lane] = 0u;
        tab_f[lane] = 0u;
      }
      for (int j = lane; j < nwin * 16; j += 32) cnt[j] = 0u;
      __syncwarp();
      // ---- pass 1 (two iterations' loads in flight together)
      auto place = [&](uint32_t r, int c) {
        least = min(least, r);
        unsigned wq = unsigned(float(c) * inv_tile);             // c < 2^24: off by one at most
        if (wq * unsigned(tile) > unsigned(c)) --wq;
        else if ((wq + 1u) * unsigned(tile) <= unsigned(c)) ++wq;
        const unsigned w = wq - unsigned(tb);
        const unsigned H = tab_h[w];
        if (H != 0u) {
          const unsigned rel = unsigned(c) - wq * unsigned(tile), cls = rel & 15u;
          const uint32_t word = (rel << 8) | r;
          const unsigned rank = atomicAdd(cnt + w * 16u + cls, 1u);
          if (rank < H) {
            g_out[tab_so[w] + deal_cell(rank, cls)] = word;
          } else {
            const unsigned k = atomicAdd(tab_x + w, 1u);
            if (k < unsigned(kDealExtras)) park[w * kDealExtras + k] = word;
          }
        }
      };
      for (unsigned i = g_lo; i < g_hi; i += 64u) {
        const bool in0 = i + unsigned(lane) < g_hi, in1 = i + 32u + unsigned(lane) < g_hi;
        uint32_t r0 = 0, r1 = 0;
        int c0 = 0, c1 = 0;
        if (in0) { r0 = row_val[i]; c0 = ld_stream_s32(row_col + i); }
        if (in1) { r1 = row_val[i + 32u]; c1 = ld_stream_s32(row_col + i + 32u); }
        if (in0) place(r0, c0);
        if (in1) place(r1, c1);
      }
      __syncwarp();
      // ---- pass 2
      for (int w = 0; w < nwin; ++w) {
        const unsigned a = __shfl_sync(kFull, my_tp, w);
        const unsigned n = __shfl_sync(kFull, my_n, w);
        if (n == 0u) continue;                                   // warp-uniform
        uint32_t *const out = g_out + tab_so[w] + lane;
        const unsigned steps = (n + 31u) >> 5;
        const int q0 = (tb + w) * tile;
        if (n > unsigned(kDealMax)) {                            // long sub-segment: column order
          for (unsigned s2 = 0; s2 < steps; ++s2) {
            uint32_t x = pad;
            if (s2 * 32u + unsigned(lane) < n) x = (uint32_t(ld_stream_s32(row_col + a + s2 * 32u) - q0) << 8) | row_val[a + s2 * 32u];
            out[size_t(s2) << 5] = x;
          }
          continue;
        }
        const unsigned parked = tab_x[w];
        if (parked > unsigned(kDealExtras)) {                    // (one class holds most of the window)
          deal_window(row_col + a, row_val + a, n, q0, pad, out, s_scratch[warp]);
          continue;
        }
        const unsigned in_class = cnt[w * 16 + (lane & 15)];
        for (int s2 = int(steps) - 1; s2 >= 0; --s2) {
          if (in_class <= 2u * unsigned(s2) + unsigned(lane >> 4)) {      // this lane's cell of step s2 is free
            uint32_t x = pad;
            if (parked != 0u) {
              const unsigned idx = atomicAdd(tab_f + w, 1u);
              if (idx < parked) x = park[w * kDealExtras + idx];
            }
            out[size_t(s2) << 5] = x;
          }
        }
      }
      __syncwarp();
    }
  }
  if (least == 0u) *zero_flag = 1u;
  if (blockIdx.x == 0 && threadIdx.x == 0) zero_flag[1] = 1u;      // layout of this workspace: bank-class cells
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

// A WARP owns one (user, column window) at a time: accumulators for the window's `tile` columns in its own slice
// of shared memory, the neighbours' sub-segments streamed in table order (entries of one neighbour hit distinct
// columns, so the read-modify-writes need no atomics; neighbours are serialised by the warp's own program order
// — no block barrier anywhere in the loop), then the window is swept by the same warp.  Warps never talk to each
// other; a CTA is just as many warps as fit one SM's shared memory.
//
// Data movement: a lane takes word `lane` of every 128-byte step straight into a register (one coalesced load
// per step), kPwBatch list entries (= up to 2 * kPwBatch steps) at a time in two ping-pong register sets.
// Two other schemes were built and measured at C2 (profiles/r2_predict_variants.md): staging the steps through
// a shared-memory ring filled by cp.async.bulk + mbarrier (7.3 ms against 4.6 ms: 128..256-byte copies are far
// too small for the TMA unit — the bulk copies are used where the pieces are kilobytes, in the MinHash tile scan),
// and reading tables / offsets / first batch of the NEXT item ahead of the sweep (no gain at equal occupancy, and
// its 120 registers cost 4 of 20 warps).  The kernel is bound by instruction issue and shared-memory wavefronts,
// not by global latency, so what counts is instructions per entry and warps per SM.
constexpr int kPwListCap = 128;             // step-list entries built per round
constexpr unsigned kPwSingle = 1u << 30;    // entry covers ONE 32-word step (else two consecutive steps of one neighbour)
constexpr int kPwBatch = 4;                 // list entries whose loads are in flight together (x2: double-buffered)
constexpr int kPwMaxWarps = 20;             // 640 threads: up to 96 registers per thread, no spills
constexpr int kTieBitWords = 26 * 32;       // single-rater half-way bits: 26 ratings (5, 15, .., 255) x 1024 milli
constexpr size_t kPwCtaBytes = round_up<size_t>(kTieBitWords * 4, 256);

// shared memory of one warp: acc[slots] | u16 jlist[tile] | step list | (F64) similarity of each neighbour
__host__ __device__ constexpr size_t pw_warp_bytes(int tile, bool milli) {
  return size_t(pw_acc_slots(tile)) * (milli ? 8 : 16) + round_up<size_t>(size_t(tile) * 2, 128) +
         size_t(kPwListCap) * 8 + (milli ? 0 : 64 * 8);
}

// 8-byte shared-memory load / store as two 32-bit halves, by shared-window byte address
__device__ __forceinline__ void lds_v2(unsigned &lo, unsigned &hi, unsigned addr) {
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts_v2(unsigned addr, unsigned lo, unsigned hi) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ uint4 ld_stream_v4(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned ld_stream_u32(const uint32_t *p) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ int predict_from(double num, double den) {
  return den != 0.0 ? hybrid_rule(0.0, num / den) : 0;
}

// Where the prepared rows live.  Rows are sharded over `n_peers` ranks of `rows_per_peer` rows each; every rank's
// preparation workspace has the same layout (same U, nt), so one set of byte offsets addresses all of them and
// base[r] is rank r's workspace as mapped into THIS process (CUDA IPC; base[my] = the local one).  Neighbour ids
// are global row ids; a neighbour row on another rank is read in place over NVLink.
//
// Staged peer rows (optional, qrs_peer_rows_stage): reading a neighbour's steps over NVLink one 128-byte line at a
// time leaves the warp waiting a link round trip per batch (C2 shards on 8 GPUs: 5 % of the gathered bytes are remote
// and the predictor takes 80 % longer).  So the rows this rank's users name on other ranks are first copied, whole
// and by many warps at once, into a local buffer; reference x = local user * ku + table slot owns the offset-table
// row stage_seg[x * nt .. x * nt + nt] (consecutive sub-segments like everywhere else), and base[n_peers] is set so
// that peer_packed(pv, n_peers) is the buffer's step area.  stage_hdr[1] != 0 says the staging of this step fitted;
// otherwise (and in the rare tie paths, which go by global row id) rows are read in place.
constexpr int kMaxPeers = 16;
struct PeerView {
  unsigned long long base[kMaxPeers];
  int n_peers;
  int my;
  long long rows_per_peer;
  unsigned long long off_tilepos, off_seg, off_packed, off_flag;   // byte offsets inside a workspace
  const unsigned long long *stage_hdr;      // null: no staged rows.  [0] steps this step needs, [1] staged and complete
  const uint64_t *stage_seg;
};
__device__ __forceinline__ const int64_t *peer_tilepos(const PeerView &pv, int r) {
  return reinterpret_cast<const int64_t *>(pv.base[r] + pv.off_tilepos);
}
__device__ __forceinline__ const uint64_t *peer_seg(const PeerView &pv, int r) {
  return reinterpret_cast<const uint64_t *>(pv.base[r] + pv.off_seg);
}
__device__ __forceinline__ const uint32_t *peer_packed(const PeerView &pv, int r) {
  return reinterpret_cast<const uint32_t *>(pv.base[r] + pv.off_packed);
}

// rating of global row v at column `col` (0 if not rated), from the prepared rows.  Column order: the words of a
// sub-segment are ascending in column, the valid ones are the first tilepos[t + 1] - tilepos[t].  Bank-class layout
// (`dealt`, sub-segments of at most kDealMax entries): the column's class cells in half-step order, and every word
// of the sub-segment only if the class filled all its cells (then it may be an extra).
__device__ __forceinline__ int prepared_rating(const PeerView &pv, int v, int col, int tile, int nt) {
  const int r = pv.n_peers == 1 ? 0 : int(unsigned(v) / unsigned(pv.rows_per_peer));      // (no 64-bit division)
  const bool dealt = reinterpret_cast<const unsigned *>(pv.base[r] + pv.off_flag)[1] != 0u;   // the owner's layout
  const long long row = v - r * pv.rows_per_peer;
  const int t = col / tile;
  const unsigned key = unsigned(col - t * tile) << 8;
  const int64_t *tp = peer_tilepos(pv, r) + row * (nt + 1) + t;
  const int n = int(tp[1] - tp[0]);
  if (n == 0) return 0;
  const uint32_t *w = peer_packed(pv, r) + (peer_seg(pv, r)[row * nt + t] << 5);
  if (dealt && n <= kDealMax) {
    // the class's cells, all requested at once: a hit, or a cell that is not the class's (a padding word — its
    // column lies beyond the window — or another class's extra) ends the search
    const unsigned steps = unsigned(n + 31) >> 5, cls = (key >> 8) & 15u;
    unsigned x[8];
#pragma unroll
    for (unsigned h = 0; h < 8; ++h) x[h] = h < 2u * steps ? w[deal_cell(h, cls)] : 0xFFFFFFFFu;
    bool full = true;
#pragma unroll
    for (unsigned h = 0; h < 8; ++h) {
      if (h < 2u * steps) {
        if ((x[h] >> 8) == (key >> 8)) return int(x[h] & 0xFFu);
        full = full && (x[h] >> 8) < unsigned(tile) && ((x[h] >> 8) & 15u) == cls;
      }
    }
    if (!full) return 0;
    // every cell holds another entry of the class: the column may be one of the class's extras, anywhere
    int found = 0;
    for (unsigned i = 0; i < steps * 32u; i += 32) {             // one step per round, its eight loads in flight together
      uint4 q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) q[j] = *reinterpret_cast<const uint4 *>(w + i + 4 * j);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if ((q[j].x >> 8) == (key >> 8)) found = int(q[j].x & 0xFFu) | 0x100;
        if ((q[j].y >> 8) == (key >> 8)) found = int(q[j].y & 0xFFu) | 0x100;
        if ((q[j].z >> 8) == (key >> 8)) found = int(q[j].z & 0xFFu) | 0x100;
        if ((q[j].w >> 8) == (key >> 8)) found = int(q[j].w & 0xFFu) | 0x100;
      }
      if (found) break;
    }
    return found & 0xFF;
  }
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (w[mid] < key) lo = mid + 1; else hi = mid;
  }
  if (lo < n) {
    const unsigned x = w[lo];
    if ((x >> 8) == (key >> 8)) return int(x & 0xFFu);
  }
  return 0;
}

// the reference's sequential fp64 evaluation of one cell, re-gathered from the neighbours' rows (rare path:
// only when the list of half-way cells is full)
__device__ __noinline__ int resolve_tie(const PeerView &pv, const int *__restrict__ u_idx,
                                        const int *__restrict__ u_milli, int64_t ul, int ku, int c, int col, int tile,
                                        int nt) {
  double num = 0.0, den = 0.0;
  for (int t = 0; t < c; ++t) {
    const int r = prepared_rating(pv, u_idx[ul * ku + t], col, tile, nt);
    if (r != 0) {
      const double s = double(u_milli[ul * ku + t]) / 1000.0;
      num = __dadd_rn(num, __dmul_rn(double(r), s));
      den = __dadd_rn(den, s);
    }
  }
  return predict_from(num, den);
}

// ONE_PEER (compile time): every neighbour row is local — no rank arithmetic, no base-pointer selection per step.
// WIDE_K (compile time): the tables hold more than 32 neighbours, i.e. a lane carries two table slots; with at most 32
// the second slot's offsets, scans and list loop do not exist
template <bool MILLI, bool ONE_PEER, bool WIDE_K>
__global__ void __launch_bounds__(kPwMaxWarps * 32, 1)
predict_sparse_kernel(const __grid_constant__ PeerView pv, int Q,
                      const int *__restrict__ u_idx, const double *__restrict__ u_val,
                      const int *__restrict__ u_milli, const int *__restrict__ u_cnt, int ku, int64_t user_base,
                      int64_t n_users, int tile, int nt,
                      int64_t *__restrict__ row_start, int *__restrict__ row_cnt, int *__restrict__ cols,
                      uint8_t *__restrict__ preds, int64_t region_cap, int n_regions,
                      unsigned long long *__restrict__ cursors, unsigned long long *__restrict__ tie_count,
                      ulonglong2 *__restrict__ tie_list, int64_t tie_capacity, int cols_mode,
                      const unsigned *__restrict__ tie_bits_g) {
  using acc_t = typename AccOf<MILLI>::type;
  unsigned long long *const user_counter = tie_count + 1;      // zeroed with the tie counter before every launch
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warps_cta = blockDim.x >> 5;
  unsigned *tie_bits = reinterpret_cast<unsigned *>(smem_raw);           // [26 * 32] MILLI: single-rater half-way bits
  unsigned char *mine = smem_raw + kPwCtaBytes + size_t(warp) * pw_warp_bytes(tile, MILLI);
  acc_t *acc = reinterpret_cast<acc_t *>(mine);
  uint16_t *jlist = reinterpret_cast<uint16_t *>(mine + size_t(pw_acc_slots(tile)) * (MILLI ? 8 : 16));
  uint2 *steps = reinterpret_cast<uint2 *>(reinterpret_cast<unsigned char *>(jlist) + round_up<size_t>(size_t(tile) * 2, 128));
  double *m_s = reinterpret_cast<double *>(steps + kPwListCap);          // [64] F64 only: similarity of each neighbour

  acc_t zero;
  if constexpr (MILLI) zero = 0ull; else zero = Acc{0.0, 0.0};
  for (int j = lane; j < pw_acc_slots(tile); j += 32) acc[j] = zero;
  if constexpr (MILLI)
    for (int w = tid; w < kTieBitWords; w += blockDim.x) tie_bits[w] = tie_bits_g[w];
  __syncthreads();          // the only block barrier: the table every warp reads

  const unsigned lt = lanemask_lt();
  const unsigned acc_s = unsigned(__cvta_generic_to_shared(acc));   // shared-window byte address of acc[0]
  // output slices are reserved from one of n_regions cursors (a single cursor would serialise millions of
  // same-address atomics in L2); region r owns [r * region_cap, (r + 1) * region_cap) and takes the users
  // ul = r (mod n_regions), so a region's demand does not depend on which warp happened to take which user

  struct Batch { unsigned p0[kPwBatch], p1[kPwBatch], tag[kPwBatch]; };
  const unsigned pad = unsigned(pw_spare_col(tile) + (lane & 15)) << 8;      // a padding word: this lane's spare slot
  // explicit zero ratings anywhere in the rows this launch can read?  (warp-uniform, rare: guarded accumulation)
  bool has_zeros = false;
  for (int r = 0; r < pv.n_peers; ++r)
    has_zeros |= *reinterpret_cast<const unsigned *>(pv.base[r] + pv.off_flag) != 0u;
  constexpr bool one_peer = ONE_PEER;
  const bool staged = !ONE_PEER && pv.stage_hdr != nullptr && pv.stage_hdr[1] != 0ull;      // warp-uniform

  // A warp takes one USER at a time (dynamic: a device counter) and walks its windows in order, so what depends
  // on the user only — neighbour ids, ranks, similarities, the rows' offset tables — is set up once, and since the
  // sub-segments of a row are consecutive in (row, window) order the end of one window is the start of the next:
  // one 4-byte load per neighbour and window.
  // non-zero predictions of this warp = the columns it listed (a listed column has D != 0, i.e. a prediction >= 18;
  // far below 2^32 per warp).  Keeping the count in shared memory instead was measured 0.5 % slower.
  unsigned nz_reg = 0;
  const uint32_t *const my_packed = peer_packed(pv, pv.my);
  for (;;) {
    unsigned ul = 0;           // a launch has fewer than 2^32 users (the tie list packs the user in 32 bits as well)
    if (lane == 0) ul = unsigned(min(atomicAdd(user_counter, 1ull), 0xFFFFFFFFull));
    ul = __shfl_sync(kFull, ul, 0);
    if (int64_t(ul) >= n_users) break;
    const int c = min(u_cnt[ul], ku);
    const int64_t tbase = int64_t(ul) * ku;

    // ---- lane <-> table slots lane and lane + 32; a neighbour is a GLOBAL row id: rank = id / rows_per_peer
    //      (bits 26..29 of the tag), row on that rank = the remainder.  so0 / so1 = the low words of the row's
    //      offset table (step counts stay below 2^32), null when the slot is empty or its similarity is 0.000
    //      (such a neighbour adds nothing to N or D, recommender.py:36-47)
    const unsigned *so0 = nullptr, *so1 = nullptr;
    unsigned tag0 = 0, tag1 = 0;
    __syncwarp();            // F64: the previous user's m_s has been consumed
    if (lane < c) {
      const int v = u_idx[tbase + lane];
      int pr = one_peer ? 0 : int(unsigned(v) / unsigned(pv.rows_per_peer));
      const long long row = one_peer ? v : v - pr * pv.rows_per_peer;
      so0 = reinterpret_cast<const unsigned *>(peer_seg(pv, pr) + row * nt);
      if (!one_peer && staged && pr != pv.my) {         // a remote row, copied here before the launch
        so0 = reinterpret_cast<const unsigned *>(pv.stage_seg + ((user_base + int64_t(ul)) * ku + lane) * nt);
        pr = pv.n_peers;
      }
      if constexpr (MILLI) {
        tag0 = unsigned(u_milli[tbase + lane]) | (1u << 24);
        if (tag0 == (1u << 24)) so0 = nullptr;
      } else {
        tag0 = unsigned(lane);
        m_s[lane] = u_val[tbase + lane];
      }
      tag0 |= unsigned(pr) << 26;
    }
    if (WIDE_K && lane + 32 < c) {
      const int v = u_idx[tbase + lane + 32];
      int pr = one_peer ? 0 : int(unsigned(v) / unsigned(pv.rows_per_peer));
      const long long row = one_peer ? v : v - pr * pv.rows_per_peer;
      so1 = reinterpret_cast<const unsigned *>(peer_seg(pv, pr) + row * nt);
      if (!one_peer && staged && pr != pv.my) {
        so1 = reinterpret_cast<const unsigned *>(pv.stage_seg + ((user_base + int64_t(ul)) * ku + lane + 32) * nt);
        pr = pv.n_peers;
      }
      if constexpr (MILLI) {
        tag1 = unsigned(u_milli[tbase + lane + 32]) | (1u << 24);
        if (tag1 == (1u << 24)) so1 = nullptr;
      } else {
        tag1 = unsigned(lane + 32);
        m_s[lane + 32] = u_val[tbase + lane + 32];
      }
      tag1 |= unsigned(pr) << 26;
    }
    // the user's own offset table: an index (a launch has fewer than 2^32 sub-segments) off a uniform base
    const uint64_t *const seg_my = peer_seg(pv, pv.my);
    const unsigned own_x = unsigned((user_base + int64_t(ul)) * nt);
    unsigned s0 = so0 ? so0[0] : 0u, s1 = (WIDE_K && so1) ? so1[0] : 0u, sown = unsigned(__ldg(seg_my + own_x));     // starts of window 0
   for (int t = 0; t < nt; ++t) {
    const int64_t item = int64_t(ul) * nt + t;
    const int q0 = t * tile;
    const int width = min(tile, Q - q0);
    // this window's sub-segments: [start, end) in 128-byte steps; the end is the next window's start
    const unsigned o0 = s0, o1 = s1, own_o = sown;
    s0 = so0 ? so0[2 * (t + 1)] : 0u;
    s1 = (WIDE_K && so1) ? so1[2 * (t + 1)] : 0u;
    sown = unsigned(__ldg(seg_my + own_x + unsigned(t + 1)));
    const unsigned n0 = s0 - o0, n1 = s1 - o1, own_n = sown - own_o;
    // the user's own row in this window (its rated cells are cleared after the accumulation): the first two steps
    // are requested below and sit in two registers until then
    unsigned own_p0 = 0, own_p1 = 0;
    // list entries (one or two steps each) in table order: slots 0..31, then 32..63
    const unsigned d0 = (n0 + 1) >> 1, d1 = (n1 + 1) >> 1;
    unsigned ex0 = d0, ex1 = d1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y0 = __shfl_up_sync(kFull, ex0, o);
      if (lane >= o) ex0 += y0;
      if constexpr (WIDE_K) {
        const unsigned y1 = __shfl_up_sync(kFull, ex1, o);
        if (lane >= o) ex1 += y1;
      }
    }
    const unsigned tot0 = __shfl_sync(kFull, ex0, 31);
    unsigned total = tot0;
    if constexpr (WIDE_K) total += __shfl_sync(kFull, ex1, 31);
    ex0 -= d0;
    ex1 = ex1 - d1 + tot0;

    if (total == 0) {       // nothing gathered: no prediction in this window (the bitmap rows are zeroed by the caller)
      if (lane == 0) { row_start[item] = int64_t(ul % unsigned(n_regions)) * region_cap; row_cnt[item] = 0; }
      continue;
    }
    if (own_n > 0) own_p0 = ld_stream_u32(my_packed + (size_t(own_o) << 5) + lane);
    if (own_n > 1) own_p1 = ld_stream_u32(my_packed + (size_t(own_o) << 5) + 32 + lane);

    for (unsigned R0 = 0; R0 < total; R0 += kPwListCap) {
      __syncwarp();          // the previous round's list has been consumed (and m_s is written)
      {
        const unsigned hi = R0 + kPwListCap;
        for (unsigned k = max(ex0, R0); k < min(ex0 + d0, hi); ++k) {
          const unsigned j = k - ex0;
          steps[k - R0] = make_uint2(o0 + 2 * j, tag0 | ((2 * j + 1 >= n0) ? kPwSingle : 0u));
        }
        if constexpr (WIDE_K) {
          for (unsigned k = max(ex1, R0); k < min(ex1 + d1, hi); ++k) {
            const unsigned j = k - ex1;
            steps[k - R0] = make_uint2(o1 + 2 * j, tag1 | ((2 * j + 1 >= n1) ? kPwSingle : 0u));
          }
        }
      }
      __syncwarp();
      const int cnt = int(min(unsigned(kPwListCap), total - R0));

      // ---- list entries in order, kPwBatch at a time, two register sets in ping-pong: the NEXT batch's words are
      //      being loaded while the current batch is accumulated.  A lane takes word `lane` of each step.
      auto load_batch = [&](Batch &B, int f0) {
#pragma unroll
        for (int i = 0; i < kPwBatch; ++i) {
          unsigned a = pad, b2 = pad, tg = 0;
          if (f0 + i < cnt) {                       // warp-uniform
            const uint2 e = steps[f0 + i];
            tg = e.y;
            const uint32_t *src = (one_peer ? my_packed : peer_packed(pv, int((tg >> 26) & 15u))) + (size_t(e.x) << 5) + lane;
            a = ld_stream_u32(src);
            if (!(tg & kPwSingle)) b2 = ld_stream_u32(src + 32);
          }
          B.p0[i] = a; B.p1[i] = b2; B.tag[i] = tg;
        }
      };
      // ZEROS (compile-time tag): the rows may hold explicit ratings of 0, which must not count as raters
      auto process_batch = [&](const Batch &B, auto zeros_tag) {
        constexpr bool ZEROS = decltype(zeros_tag)::value;
#pragma unroll
        for (int i = 0; i < kPwBatch; ++i) {
          const unsigned r0 = B.p0[i] & 0xFFu, r1 = B.p1[i] & 0xFFu;
          if constexpr (MILLI) {
            // both read-modify-writes of the lane at once, no predicate: a padding word (rating 0) points at the
            // spare slot, and the two entries belong to one neighbour, i.e. different columns.  inc = milli | 1 << 24
            // is the low word's increment (D and the rater count, no carry out of 32 bits: at most 64 raters of
            // 1000) AND the high word's multiplier: r * inc = r << 24 | r * milli, as r * milli < 2^24
            const unsigned inc = B.tag[i] & 0x1FFFFFFu;       // bits 26.. are the rank and the single-step flag
            const unsigned a0 = acc_s + ((B.p0[i] >> 5) & 0x7FFFFF8u), a1 = acc_s + ((B.p1[i] >> 5) & 0x7FFFFF8u);
            unsigned lo0, hi0, lo1, hi1;
            lds_v2(lo0, hi0, a0);
            lds_v2(lo1, hi1, a1);
            lo0 += ZEROS ? (r0 ? inc : 0u) : inc;
            lo1 += ZEROS ? (r1 ? inc : 0u) : inc;
            hi0 += r0 * inc;
            hi1 += r1 * inc;
            sts_v2(a0, lo0, hi0);
            sts_v2(a1, lo1, hi1);
          } else {
            const double s = m_s[B.tag[i] & 63u];
            if (r0) {
              Acc x = acc[B.p0[i] >> 8];
              x.num = __dadd_rn(x.num, __dmul_rn(double(r0), s));
              x.den = __dadd_rn(x.den, s);
              acc[B.p0[i] >> 8] = x;
            }
            if (r1) {
              Acc x = acc[B.p1[i] >> 8];
              x.num = __dadd_rn(x.num, __dmul_rn(double(r1), s));
              x.den = __dadd_rn(x.den, s);
              acc[B.p1[i] >> 8] = x;
            }
          }
          __syncwarp();      // the next entry may be another neighbour: same columns from other lanes
        }
      };
      // (a flavour that skips the second word when every entry of a batch is a single step — C3's short
      //  sub-segments — was measured: no gain at C3, +7 % at C2 from the larger loop body; not kept)
      auto process = [&](const Batch &B) {
        if (has_zeros) process_batch(B, std::true_type{});
        else process_batch(B, std::false_type{});
      };
      {
        Batch A, B;
        load_batch(A, 0);
        for (int f0 = 0; f0 < cnt; f0 += 2 * kPwBatch) {
          const bool more = f0 + kPwBatch < cnt;
          if (more) load_batch(B, f0 + kPwBatch);
          process(A);
          if (more) {
            if (f0 + 2 * kPwBatch < cnt) load_batch(A, f0 + 2 * kPwBatch);
            process(B);
          }
        }
      }
    }

    // ---- the user's own rated cells are not predicted (recommender.py:302): drop what landed there
    if (own_p0 & 0xFFu) acc[own_p0 >> 8] = zero;
    if (own_p1 & 0xFFu) acc[own_p1 >> 8] = zero;
    for (unsigned s = 2; s < own_n; ++s) {          // rows with more than 64 entries in one window
      const unsigned p = ld_stream_u32(my_packed + ((size_t(own_o) + s) << 5) + lane);
      if (p & 0xFFu) acc[p >> 8] = zero;
    }
    __syncwarp();

    // ---- pass 1: ascending list of the window's columns that get a prediction byte, 64 columns per round.
    const bool bitmap = cols_mode == 2;
    unsigned *const bitmap_row = reinterpret_cast<unsigned *>(cols) + item * (tile >> 5);
    int touched = 0;
    if constexpr (MILLI) {
      // A lane looks at columns j0 + lane and j0 + 32 + lane: the two ballots ARE the bitmap words of the 64 columns
      // Listed iff D != 0 (then the prediction is >= 18).  Neighbours of similarity 0 never enter the step list, so
      // an accumulator is either untouched (all zero) or has D != 0; columns beyond the window's width stay zero.
      // Only the low words are read (32-bit loads, explicit so that the compiler does not widen the test).
      unsigned my_word = 0u;
      const unsigned jl_s = unsigned(__cvta_generic_to_shared(jlist));
      unsigned a_s = acc_s + 8u * unsigned(lane);
      const unsigned ltm = (1u << lane) - 1u;      // = %lanemask_lt, as ALU work (the special-register read is slow)
      // one 64-column round; BITMAP (compile-time tag): the ballots are also kept as the window's bitmap words
      auto list_round = [&](int j0, auto bitmap_tag) {
        constexpr bool BITMAP = decltype(bitmap_tag)::value;
        unsigned x0, x1;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x0) : "r"(a_s));            // acc holds round_up(tile, 64) words
        asm volatile("ld.shared.u32 %0, [%1+256];" : "=r"(x1) : "r"(a_s));
        a_s += 512u;
        const bool hit0 = (x0 & 0xFFFFFFu) != 0u, hit1 = (x1 & 0xFFFFFFu) != 0u;
        const unsigned m0 = __ballot_sync(kFull, hit0), m1 = __ballot_sync(kFull, hit1);
        const int c0 = __popc(m0);
        if (hit0) asm volatile("st.shared.u16 [%0], %1;" ::"r"(jl_s + 2u * unsigned(touched + __popc(m0 & ltm))), "h"((unsigned short)(j0 + lane)) : "memory");
        if (hit1) asm volatile("st.shared.u16 [%0], %1;" ::"r"(jl_s + 2u * unsigned(touched + c0 + __popc(m1 & ltm))), "h"((unsigned short)(j0 + 32 + lane)) : "memory");
        touched += c0 + __popc(m1);
        if constexpr (BITMAP) {
          // lane w keeps word w of the current 1024-column group; one coalesced store per group.  (Parking the words in
          // shared memory through lane 0 instead: fewer instructions, measured 2 % slower.)
          const int w0 = (j0 >> 5) & 31;
          my_word = lane == w0 ? m0 : (lane == w0 + 1 ? m1 : my_word);
          if (w0 == 30 || j0 + 64 >= width) {        // warp-uniform: the group of 32 words is complete (or the window ends)
            if ((j0 & ~1023) + 32 * lane < width) bitmap_row[((j0 >> 10) << 5) + lane] = my_word;
            my_word = 0u;
          }
        }
      };
      if (bitmap) {
#pragma unroll 1
        for (int j0 = 0; j0 < width; j0 += 64) list_round(j0, std::true_type{});
      } else {        // explicit column ids (sparse windows, C3 / C4): the listing pass is a quarter of the kernel there
#pragma unroll 1
        for (int j0 = 0; j0 < width; j0 += 64) list_round(j0, std::false_type{});
      }
    } else {
#pragma unroll 2
      for (int j0 = 0; j0 < width; j0 += 64) {
        const int ja = j0 + lane, jb = ja + 32;
        bool hit0 = false, hit1 = false;
        if (ja < width) { const Acc x = acc[ja]; hit0 = x.den != 0.0 || x.num != 0.0; }
        if (jb < width) { const Acc x = acc[jb]; hit1 = x.den != 0.0 || x.num != 0.0; }
        const unsigned m0 = __ballot_sync(kFull, hit0), m1 = __ballot_sync(kFull, hit1);
        if (bitmap && lane < 2 && j0 + 32 * lane < width) bitmap_row[(j0 >> 5) + lane] = lane == 0 ? m0 : m1;
        const int c0 = __popc(m0);
        if (hit0) jlist[touched + __popc(m0 & lt)] = uint16_t(ja);
        if (hit1) jlist[touched + c0 + __popc(m1 & lt)] = uint16_t(jb);
        touched += c0 + __popc(m1);
      }
    }
    unsigned long long base = 0;
    const int region = int(ul % unsigned(n_regions));
    const int64_t region_base = int64_t(region) * region_cap;
    if (lane == 0) {
      if (touched > 0) base = atomicAdd(cursors + region, (unsigned long long)touched);
      row_start[item] = region_base + int64_t(base);
      row_cnt[item] = touched;
    }
    base = __shfl_sync(kFull, base, 0);
    __syncwarp();            // jlist is read by other lanes than those that wrote it

    // ---- pass 2: one prediction byte per listed column, written straight into the reserved slice (ascending);
    //      accumulators back to zero
    {
      const int64_t wbase = int64_t(base);
      const int fit = int(max(int64_t(0), min(int64_t(touched), region_cap - wbase)));   // region overflow: dropped
      uint8_t *const w_preds = preds + (region_base + wbase);
      char *const w_cols = reinterpret_cast<char *>(cols) + (region_base + wbase) * (cols_mode == 1 ? 2 : 4);
      nz_reg += unsigned(touched);
#pragma unroll 2
      for (int x0 = 0; x0 < touched; x0 += 32) {
        const int x = x0 + lane;
        int p = 0;
        if (x < touched) {
          const int j = jlist[x];
          const acc_t a = acc[j];
          acc[j] = zero;
          bool tie = false;
          if constexpr (MILLI) {
            const unsigned lo = unsigned(a), hi = unsigned(a >> 32);
            const unsigned D = lo & 0xFFFFFFu, raters = lo >> 24;
            const unsigned N = hi & 0xFFFFFFu;
            const unsigned X = 7u * N + 180u * D, Y = 10u * D;    // x = X / Y = 0.7 * N / D + 18
            // q = X / Y via a float estimate (q < 256, estimate off by at most one) corrected exactly
            float ry;                                               // Y >= 10: no range handling needed
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(float(Y)));
            unsigned q = unsigned(float(X) * ry);
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
                                                 ((unsigned long long)ul << 32) | unsigned(q0 + j));
              else
                p = resolve_tie(pv, u_idx, u_milli, int64_t(ul), ku, c, q0 + j, tile, nt);   // list full: settle here
            }
            if (cols_mode == 1) reinterpret_cast<uint16_t *>(w_cols)[x] = uint16_t(q0 + j);
            else if (cols_mode == 0) reinterpret_cast<int *>(w_cols)[x] = q0 + j;
            w_preds[x] = uint8_t(p);
          }
        }
      }
    }
    __syncwarp();
   }
  }
  // non-zero predictions of this launch (statistics): one atomic per warp
  if (lane == 0 && nz_reg > 0) atomicAdd(cursors + n_regions, (unsigned long long)nz_reg);
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
  extern __shared__ __align__(128) unsigned char smem_raw[];
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
// neighbours' ratings of that column (lane <-> neighbour, binary search in the neighbour's prepared
// sub-segment, wherever it lives) and replays the reference's sequential fp64 sums in table order.
__global__ void __launch_bounds__(256)
predict_tie_fixup_kernel(const __grid_constant__ PeerView pv, int tile, int nt, const int *__restrict__ u_idx,
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
        s = double(u_milli[ul * ku + t0 + lane]) / 1000.0;
        r = prepared_rating(pv, u_idx[ul * ku + t0 + lane], col, tile, nt);
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

// ------------------------------------------------------------------ staged peer rows
// steps of the prepared row behind reference x = (local user, table slot), 0 unless the row lives on another rank
__global__ void __launch_bounds__(256)
peer_ref_steps_kernel(const __grid_constant__ PeerView pv, int nt, const int *__restrict__ u_idx,
                      const int *__restrict__ u_cnt, int ku, int64_t n_users, uint32_t *__restrict__ ref_steps,
                      uint64_t *__restrict__ ref_src) {
  const int64_t total = n_users * ku;
  for (int64_t x = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; x < total; x += int64_t(gridDim.x) * blockDim.x) {
    const int64_t ul = x / ku;
    const int t = int(x - ul * ku);
    uint32_t steps = 0;
    if (t < min(u_cnt[ul], ku)) {
      const int v = u_idx[x];
      const int pr = int(unsigned(v) / unsigned(pv.rows_per_peer));
      if (pr != pv.my) {
        const uint64_t *seg = peer_seg(pv, pr) + (v - pr * pv.rows_per_peer) * nt;
        const uint64_t first = seg[0];
        steps = uint32_t(seg[nt] - first);
        ref_src[x] = first;        // (so that the copy needs no second trip over the link before it can start)
      }
    }
    ref_steps[x] = steps;
  }
}

// One warp per remote reference: the row's steps and its offset-table row, rebased to the staging buffer.  Nothing
// is copied (and hdr[1] stays 0: rows are read in place this step) when the step needs more than `cap_steps`.
__global__ void __launch_bounds__(256)
peer_rows_copy_kernel(const __grid_constant__ PeerView pv, int nt, const int *__restrict__ u_idx,
                      const int *__restrict__ u_cnt, int ku, int64_t n_refs,
                      const uint32_t *__restrict__ ref_steps, const uint64_t *__restrict__ ref_src,
                      const uint64_t *__restrict__ ref_first,
                      int64_t cap_steps, unsigned long long *__restrict__ hdr, uint64_t *__restrict__ stage_seg,
                      uint32_t *__restrict__ stage_packed) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const uint64_t need = ref_first[n_refs];
  const bool fits = need <= uint64_t(cap_steps);
  if (blockIdx.x == 0 && threadIdx.x == 0) { hdr[0] = need; hdr[1] = fits ? 1ull : 0ull; }
  if (!fits) return;
  // 32 references looked at per warp iteration (most are local rows or empty slots), the remote ones copied one
  // after the other by the whole warp
  for (int64_t x0 = warp_global * 32; x0 < n_refs; x0 += nwarps * 32) {
    const int64_t xl = x0 + lane;
    bool remote = false;
    int v = 0, pr = 0;
    uint32_t my_steps = 0;
    uint64_t my_src = 0, my_dst = 0;
    if (xl < n_refs) {
      const int64_t ul = xl / ku;
      if (int(xl - ul * ku) < min(u_cnt[ul], ku)) {
        v = u_idx[xl];
        pr = int(unsigned(v) / unsigned(pv.rows_per_peer));
        remote = pr != pv.my;                                    // (an empty remote row still needs its offsets)
        if (remote) { my_steps = ref_steps[xl]; my_src = ref_src[xl]; my_dst = ref_first[xl]; }
      }
    }
    for (unsigned todo = __ballot_sync(kFull, remote); todo != 0u; todo &= todo - 1u) {
      const int from = __ffs(todo) - 1;
      const int64_t x = x0 + from;
      const int vx = __shfl_sync(kFull, v, from), prx = __shfl_sync(kFull, pr, from);
      const uint32_t steps = __shfl_sync(kFull, my_steps, from);
      const uint64_t src0 = __shfl_sync(kFull, my_src, from), dst0 = __shfl_sync(kFull, my_dst, from);
      const uint64_t *seg = peer_seg(pv, prx) + (vx - prx * pv.rows_per_peer) * nt;
      // the row's steps as one run of 16-byte pieces (eight per step): the offset-table row and up to 32 steps
      // leave in ONE trip over the link, 512 bytes per warp load
      const uint4 *src = reinterpret_cast<const uint4 *>(peer_packed(pv, prx) + (src0 << 5)) + lane;
      uint4 *dst = reinterpret_cast<uint4 *>(stage_packed + (dst0 << 5)) + lane;
      const uint32_t pieces = steps * 8u;
      uint64_t off[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) off[j] = lane + 32 * j <= nt ? seg[lane + 32 * j] : 0ull;
      for (uint32_t q = 0; q < pieces; q += 256u) {
        uint4 w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (q + 32u * j + unsigned(lane) < pieces) w[j] = ld_stream_v4(src + q + 32 * j);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (q + 32u * j + unsigned(lane) < pieces) dst[q + 32 * j] = w[j];
      }
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (lane + 32 * j <= nt) stage_seg[x * nt + lane + 32 * j] = dst0 + (off[j] - src0);
      for (int t = lane + 64; t <= nt; t += 32) stage_seg[x * nt + t] = dst0 + (seg[t] - src0);
    }
  }
}

// warps (each with its own window) that fit one CTA = one SM: 227 KB minus the CTA's table and 1 KB the system keeps
static int predict_warps_per_cta(int tile, bool milli) {
  const size_t budget = size_t(227) * 1024 - 1024 - kPwCtaBytes;
  const size_t per = pw_warp_bytes(tile, milli);
  const size_t w = budget / per;
  return int(w > size_t(kPwMaxWarps) ? size_t(kPwMaxWarps) : w);
}

template <bool MILLI>
static int launch_predict_sparse(const PeerView &pv, int Q, const int *u_idx, const double *u_val,
                                 const int *u_milli, const int *u_cnt, int ku, int64_t user_base, int64_t n_users,
                                 int tile, int nt, int64_t *row_start, int *row_cnt, int *cols,
                                 uint8_t *preds, int64_t region_cap, int n_regions, unsigned long long *cursors,
                                 unsigned long long *tie_count, ulonglong2 *tie_list, int64_t tie_capacity,
                                 int cols_mode, const unsigned *single_tab, cudaStream_t st) {
  auto kern = pv.n_peers == 1 ? (ku > 32 ? predict_sparse_kernel<MILLI, true, true> : predict_sparse_kernel<MILLI, true, false>)
                              : (ku > 32 ? predict_sparse_kernel<MILLI, false, true> : predict_sparse_kernel<MILLI, false, false>);
  const int warps = predict_warps_per_cta(tile, MILLI);
  if (warps < 1) return QRS_ERR_UNSUPPORTED;
  const size_t smem = kPwCtaBytes + size_t(warps) * pw_warp_bytes(tile, MILLI);
  QRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int64_t grid = sm_count();                     // one CTA per SM: as many independent warps as its shared memory holds
  const int64_t need = ceil_div<int64_t>(n_users * nt, warps);
  if (grid > need) grid = need;
  kern<<<unsigned(grid), warps * 32, smem, st>>>(pv, Q, u_idx, u_val, u_milli, u_cnt, ku, user_base, n_users, tile, nt,
                                                  row_start, row_cnt, cols, preds, region_cap, n_regions, cursors,
                                                  tie_count, tie_list, tie_capacity, cols_mode, single_tab);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // namespace qrs

using namespace qrs;

namespace {

int64_t tie_list_capacity(int64_t capacity) { return capacity / 32 + 65536; }

// Layout of the prepared rows.  The bank-class cells pay where the predictor is bound by shared-memory wavefronts,
// i.e. where a sub-segment fills its steps (C2: 50 entries per window, predict stage 3.86 -> 3.60 ms for 0.18 ms more
// preparation); with ~10 entries per window (C3 / C4) the predictor is bound by instruction issue, gains nothing
// (24.8 -> 24.4 ms per 100k rows) and the dearer preparation (+1.0 ms) is a loss: column order there.  The rule
// looks at this workspace's rows only; the layout is recorded in the workspace (flag word 1) for whoever reads it.
// QRS_PREDICT_PACK=class|column overrides (A/B measurements, tests).
bool pack_dealt(int64_t U, int64_t nt, int64_t nnz, int tile) {
  if (nt * int64_t(tile) > (int64_t(1) << 24)) return false;      // (the kernel finds an entry's window in fp32)
  const char *e = getenv("QRS_PREDICT_PACK");
  if (e && strcmp(e, "column") == 0) return false;
  if (e && strcmp(e, "class") == 0) return true;
  return nnz >= 24 * U * nt;
}
constexpr size_t kSingleTabBytes = round_up<size_t>(26 * 32 * 4, 256);

// workspace: tile index | steps per sub-segment | seg_off | scan scratch | tie counter | single-rater table |
// packed rows | tie list
struct PredictWs {
  int64_t *tilepos;
  uint32_t *steps;
  uint64_t *seg_off;
  void *scan_ws;
  size_t scan_bytes;
  unsigned long long *tie_count;
  unsigned *single_tab;
  ulonglong2 *tie_list;
  int64_t tie_cap;
  uint32_t *packed;
  size_t total;
};

PredictWs carve_predict_ws(void *ws, int64_t U, int64_t nt, int64_t nnz, int64_t capacity) {
  PredictWs w;
  char *p = static_cast<char *>(ws);
  const size_t segs = size_t(U > 0 ? U : 1) * size_t(nt);
  w.tilepos = reinterpret_cast<int64_t *>(p);
  p += round_up<size_t>(size_t(U > 0 ? U : 1) * size_t(nt + 1) * 8, 256);
  w.steps = reinterpret_cast<uint32_t *>(p);
  p += round_up<size_t>(segs * 4, 256);
  w.seg_off = reinterpret_cast<uint64_t *>(p);
  p += round_up<size_t>((segs + 1) * 8, 256);
  w.scan_ws = p;
  w.scan_bytes = scan_workspace_bytes(int64_t(segs));
  p += round_up<size_t>(w.scan_bytes, 256);
  w.tie_count = reinterpret_cast<unsigned long long *>(p);
  p += 256;
  w.single_tab = reinterpret_cast<unsigned *>(p);
  p += kSingleTabBytes;
  w.packed = reinterpret_cast<uint32_t *>(p);
  // every sub-segment is padded to whole 32-word steps: at most nnz + 31 words per sub-segment
  p += round_up<size_t>((size_t(nnz > 0 ? nnz : 0) + 32 * segs) * 4, 256);
  // last, so that everything the preparation wrote sits at offsets that do not depend on `capacity`
  w.tie_cap = tie_list_capacity(capacity);
  w.tie_list = reinterpret_cast<ulonglong2 *>(p);
  p += round_up<size_t>(size_t(w.tie_cap) * 16, 256);
  w.total = size_t(p - static_cast<char *>(ws));
  return w;
}

}  // namespace

extern "C" {

/* widest column window one warp can own (its accumulators + column list must fit one SM's shared memory) */
int32_t qrs_predict_max_tile(int32_t integer_tables) {
  int tile = 32;
  while (tile + 32 <= 65536 && predict_warps_per_cta(tile + 32, integer_tables != 0) >= 1) tile += 32;
  return tile;
}

/* independent warps (each owning one (user, window) at a time) that share an SM for this window width */
int32_t qrs_predict_warps_per_sm(int32_t tile, int32_t integer_tables) {
  if (tile <= 0 || tile % 32 != 0) return 0;
  return predict_warps_per_cta(tile, integer_tables != 0);
}

size_t qrs_predict_sparse_workspace_bytes(int64_t U, int64_t nnz, int32_t Q, int32_t tile, int64_t capacity) {
  if (U < 0 || nnz < 0 || Q <= 0 || tile <= 0 || capacity < 0) return 0;
  return carve_predict_ws(nullptr, U, ceil_div<int64_t>(Q, tile), nnz, capacity).total;
}

int qrs_predict_prepare(const int64_t *rowptr, const int32_t *colidx, const uint8_t *val, int64_t U, int64_t nnz,
                        int32_t Q, int32_t tile, int64_t capacity, void *ws, size_t ws_bytes, void *stream) {
  if (!rowptr || U < 0 || nnz < 0 || Q <= 0 || capacity < 0) return QRS_ERR_INVALID_ARG;
  if (tile <= 0 || tile % 32 != 0) return QRS_ERR_INVALID_ARG;
  if (nnz >= (int64_t(1) << 36)) return QRS_ERR_UNSUPPORTED;
  const int64_t nt64 = ceil_div<int64_t>(Q, tile);
  if (nt64 > (1 << 20) || U * nt64 >= (int64_t(1) << 32)) return QRS_ERR_UNSUPPORTED;
  const int nt = int(nt64);
  if (!ws || ws_bytes < qrs_predict_sparse_workspace_bytes(U, nnz, Q, tile, capacity)) return QRS_ERR_WORKSPACE;
  if (U == 0) return QRS_OK;
  if (nnz > 0 && (!colidx || !val)) return QRS_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  const PredictWs w = carve_predict_ws(ws, U, nt, nnz, capacity);
  QRS_CUDA(cudaMemsetAsync(w.tie_count, 0, 256, st));      // tie counter | next-user counter | "explicit zero ratings" flag, layout flag
  const int64_t cap = int64_t(sm_count()) * 8;
  {
    int64_t blocks = ceil_div<int64_t>(U * (nt + 1), 256);
    if (blocks > cap) blocks = cap;
    tile_index_kernel<<<unsigned(blocks), 256, 0, st>>>(rowptr, colidx, U, tile, nt, w.tilepos);
    QRS_LAUNCH_CHECK();
  }
  {
    int64_t blocks = ceil_div<int64_t>(U * nt, 256);
    if (blocks > cap) blocks = cap;
    seg_steps_kernel<<<unsigned(blocks), 256, 0, st>>>(w.tilepos, U, nt, w.steps);
    QRS_LAUNCH_CHECK();
  }
  int rc = exclusive_scan_u32_to_u64(w.steps, w.seg_off, U * nt, w.scan_ws, w.scan_bytes, st);
  if (rc != QRS_OK) return rc;
  {
    int64_t blocks = ceil_div<int64_t>(U, 8);
    const int64_t cap2 = int64_t(sm_count()) * 16;
    if (blocks > cap2) blocks = cap2;
    if (pack_dealt(U, nt, nnz, tile))
      pack_rows_dealt_kernel<<<unsigned(blocks), 256, 0, st>>>(colidx, val, w.tilepos, w.seg_off, U, nt, tile, w.packed,
                                                               reinterpret_cast<unsigned *>(w.tie_count + 2));
    else
      pack_rows_kernel<<<unsigned(blocks), 256, 0, st>>>(colidx, val, w.tilepos, w.seg_off, U, nt, tile, w.packed,
                                                         reinterpret_cast<unsigned *>(w.tie_count + 2));
    QRS_LAUNCH_CHECK();
  }
  single_rater_tie_bits_kernel<<<ceil_div(26 * 32, 256), 256, 0, st>>>(w.single_tab);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

namespace {
struct StageBuf {
  unsigned long long *hdr;
  uint64_t *seg;
  uint32_t *ref_steps;
  uint64_t *ref_src;
  uint64_t *ref_first;
  void *scan_ws;
  size_t scan_bytes;
  uint32_t *packed;
  size_t total;
};
StageBuf carve_stage(void *buf, int64_t n_users, int64_t ku, int64_t nt, int64_t cap_steps) {
  StageBuf b;
  char *p = static_cast<char *>(buf);
  const size_t refs = size_t(n_users > 0 ? n_users : 1) * size_t(ku);
  b.hdr = reinterpret_cast<unsigned long long *>(p);
  p += 256;
  b.seg = reinterpret_cast<uint64_t *>(p);
  p += round_up<size_t>((refs * size_t(nt) + 1) * 8, 256);
  b.ref_steps = reinterpret_cast<uint32_t *>(p);
  p += round_up<size_t>(refs * 4, 256);
  b.ref_src = reinterpret_cast<uint64_t *>(p);
  p += round_up<size_t>(refs * 8, 256);
  b.ref_first = reinterpret_cast<uint64_t *>(p);
  p += round_up<size_t>((refs + 1) * 8, 256);
  b.scan_ws = p;
  b.scan_bytes = scan_workspace_bytes(int64_t(refs));
  p += round_up<size_t>(b.scan_bytes, 256);
  b.packed = reinterpret_cast<uint32_t *>(p);
  p += round_up<size_t>(size_t(cap_steps > 0 ? cap_steps : 0) * 128, 256);
  b.total = size_t(p - static_cast<char *>(buf));
  return b;
}

int fill_peer_view(PeerView &pv, void *ws, const PredictWs &w, const uint64_t *peer_ws, int n_peers, int my_peer,
                   int64_t U) {
  memset(&pv, 0, sizeof(pv));
  pv.n_peers = peer_ws ? n_peers : 1;
  pv.my = peer_ws ? my_peer : 0;
  pv.rows_per_peer = U > 0 ? U : 1;
  for (int r = 0; r < pv.n_peers; ++r) pv.base[r] = peer_ws ? peer_ws[r] : 0ull;
  if (ws) pv.base[pv.my] = static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(ws));
  pv.off_tilepos = static_cast<unsigned long long>(reinterpret_cast<char *>(w.tilepos) - static_cast<char *>(ws));
  pv.off_seg = static_cast<unsigned long long>(reinterpret_cast<char *>(w.seg_off) - static_cast<char *>(ws));
  pv.off_packed = static_cast<unsigned long long>(reinterpret_cast<char *>(w.packed) - static_cast<char *>(ws));
  pv.off_flag = static_cast<unsigned long long>(reinterpret_cast<char *>(w.tie_count + 2) - static_cast<char *>(ws));
  for (int r = 0; r < pv.n_peers; ++r)
    if (pv.base[r] == 0ull) return QRS_ERR_INVALID_ARG;
  return QRS_OK;
}
}  // namespace

size_t qrs_peer_rows_stage_bytes(int64_t n_users, int32_t ku, int32_t Q, int32_t tile, int64_t cap_steps) {
  if (n_users < 0 || ku <= 0 || Q <= 0 || tile <= 0 || cap_steps < 0) return 0;
  return carve_stage(nullptr, n_users, ku, ceil_div<int64_t>(Q, tile), cap_steps).total;
}

int qrs_peer_rows_stage(const uint64_t *peer_ws, int32_t n_peers, int32_t my_peer, int64_t U, int32_t Q, int32_t tile,
                        const int32_t *u_idx, const int32_t *u_cnt, int32_t ku, int64_t n_users, int64_t cap_steps,
                        void *stage, size_t stage_bytes, void *stream) {
  if (!peer_ws || n_peers < 2 || n_peers >= kMaxPeers || my_peer < 0 || my_peer >= n_peers || U <= 0 || Q <= 0 ||
      tile <= 0 || tile % 32 != 0 || !u_idx || !u_cnt || ku <= 0 || n_users < 0 || n_users > U || cap_steps < 0)
    return QRS_ERR_INVALID_ARG;
  if (!stage || (reinterpret_cast<uintptr_t>(stage) & 255) != 0 ||
      stage_bytes < qrs_peer_rows_stage_bytes(n_users, ku, Q, tile, cap_steps))
    return QRS_ERR_WORKSPACE;
  const int nt = int(ceil_div<int64_t>(Q, tile));
  const StageBuf b = carve_stage(stage, n_users, ku, nt, cap_steps);
  cudaStream_t st = as_stream(stream);
  QRS_CUDA(cudaMemsetAsync(b.hdr, 0, 256, st));
  if (n_users == 0) return QRS_OK;
  // every rank's workspace has the same layout up to the packed rows (it does not depend on nnz or capacity)
  PeerView pv;
  void *mine = reinterpret_cast<void *>(uintptr_t(peer_ws[my_peer]));
  int rc = fill_peer_view(pv, mine, carve_predict_ws(mine, U, nt, 0, 0), peer_ws, n_peers, my_peer, U);
  if (rc != QRS_OK) return rc;
  const int64_t refs = n_users * ku;
  {
    int64_t blocks = ceil_div<int64_t>(refs, 256);
    const int64_t cap = int64_t(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    peer_ref_steps_kernel<<<unsigned(blocks), 256, 0, st>>>(pv, nt, u_idx, u_cnt, ku, n_users, b.ref_steps, b.ref_src);
    QRS_LAUNCH_CHECK();
  }
  rc = exclusive_scan_u32_to_u64(b.ref_steps, b.ref_first, refs, b.scan_ws, b.scan_bytes, st);
  if (rc != QRS_OK) return rc;
  peer_rows_copy_kernel<<<sm_count() * 8, 256, 0, st>>>(pv, nt, u_idx, u_cnt, ku, refs, b.ref_steps, b.ref_src, b.ref_first, cap_steps,
                                                        b.hdr, b.seg, b.packed);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

int qrs_predict_sparse(const int64_t *rowptr, const int32_t *colidx, const uint8_t *val, int64_t U, int64_t nnz,
                       int32_t Q, const int32_t *u_idx, const double *u_val, const int32_t *u_milli,
                       const int32_t *u_cnt, int32_t ku, int64_t user_base, int64_t n_users, int32_t tile, void *ws,
                       size_t ws_bytes, int32_t prepared, const uint64_t *peer_ws, int32_t n_peers, int32_t my_peer,
                       const void *peer_stage, int64_t stage_users, int64_t *row_start, int32_t *row_cnt, void *cols_any,
                       int32_t cols_mode, uint8_t *preds, int64_t capacity, int32_t n_regions,
                       unsigned long long *cursors, void *stream) {
  int32_t *cols = static_cast<int32_t *>(cols_any);
  if (cols_mode < 0 || cols_mode > 2 || (cols_mode == 1 && Q > 65536)) return QRS_ERR_INVALID_ARG;
  if (U < 0 || nnz < 0 || Q <= 0 || !u_idx || (!u_val && !u_milli) || !u_cnt || ku <= 0 || user_base < 0 ||
      n_users < 0 || user_base + n_users > U || !row_start || !row_cnt || !cursors || capacity < 0 || n_regions <= 0)
    return QRS_ERR_INVALID_ARG;
  if (!prepared && !rowptr) return QRS_ERR_INVALID_ARG;
  if (peer_ws ? (n_peers < 1 || n_peers > kMaxPeers || my_peer < 0 || my_peer >= n_peers || !prepared)
              : (n_peers > 1))
    return QRS_ERR_INVALID_ARG;      // remote rows can only be read from workspaces every rank has prepared already
  const int64_t region_cap = capacity / n_regions;
  if (capacity > 0 && (!cols || !preds)) return QRS_ERR_INVALID_ARG;
  if (tile <= 0 || tile % 32 != 0) return QRS_ERR_INVALID_ARG;
  if (tile > qrs_predict_max_tile(u_milli != nullptr) || ku > QRS_MAX_TOPK) return QRS_ERR_UNSUPPORTED;
  if (n_users >= (int64_t(1) << 32)) return QRS_ERR_UNSUPPORTED;      // the kernel counts users in 32 bits
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255) != 0 ||
      ws_bytes < qrs_predict_sparse_workspace_bytes(U, nnz, Q, tile, capacity))
    return QRS_ERR_WORKSPACE;
  if (n_users == 0 || U == 0) return QRS_OK;
  cudaStream_t st = as_stream(stream);
  if (!prepared) {
    if (nnz > 0 && (!colidx || !val)) return QRS_ERR_INVALID_ARG;
    int rc = qrs_predict_prepare(rowptr, colidx, val, U, nnz, Q, tile, capacity, ws, ws_bytes, stream);
    if (rc != QRS_OK) return rc;
  }
  const int nt = int(ceil_div<int64_t>(Q, tile));
  const PredictWs w = carve_predict_ws(ws, U, nt, nnz, capacity);
  PeerView pv;
  {
    int rc = fill_peer_view(pv, ws, w, peer_ws, n_peers, my_peer, U);
    if (rc != QRS_OK) return rc;
  }
  if (peer_stage) {      // rows of other ranks staged by qrs_peer_rows_stage for the table of `stage_users` local users
    if (!peer_ws || n_peers >= kMaxPeers || stage_users < user_base + n_users) return QRS_ERR_INVALID_ARG;
    const StageBuf b = carve_stage(const_cast<void *>(peer_stage), stage_users, ku, nt, 0);
    pv.stage_hdr = b.hdr;
    pv.stage_seg = b.seg;
    pv.base[pv.n_peers] = static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(b.packed)) - pv.off_packed;
  }
  QRS_CUDA(cudaMemsetAsync(w.tie_count, 0, 16, st));      // tie counter | next-user counter
  if (u_milli) {
    int rc = launch_predict_sparse<true>(pv, Q, u_idx, u_val, u_milli, u_cnt, ku, user_base, n_users, tile, nt,
                                         row_start, row_cnt, cols, preds, region_cap, n_regions, cursors, w.tie_count,
                                         w.tie_list, w.tie_cap, cols_mode, w.single_tab, st);
    if (rc != QRS_OK || capacity == 0) return rc;
    predict_tie_fixup_kernel<<<sm_count() * 8, 256, 0, st>>>(pv, tile, nt, u_idx, u_milli, u_cnt, ku, w.tie_count,
                                                             w.tie_list, w.tie_cap, preds);
    QRS_LAUNCH_CHECK();
    return QRS_OK;
  }
  return launch_predict_sparse<false>(pv, Q, u_idx, u_val, u_milli, u_cnt, ku, user_base, n_users, tile, nt, row_start,
                                      row_cnt, cols, preds, region_cap, n_regions, cursors, w.tie_count, w.tie_list,
                                      w.tie_cap, cols_mode, w.single_tab, st);
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
