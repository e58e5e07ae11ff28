// a10 + a11 — similarity-weighted rating prediction (recommender.py:36-47) and the hybrid
// rule of compute_scores (recommender.py:311-331).
//
// The reference visits every zero cell (i, j) and gathers column j at user i's neighbours.
// Here the loop is turned inside out: for user i the neighbours' CSR rows are streamed in
// table order and scattered into per-column fp64 accumulators (num += r*s, den += s), which
// reproduces, per cell, exactly the reference's left-to-right fp64 sums — cells no
// neighbour rated keep den == 0 and are "missed" (prediction 0) without being visited.
//
//   sparse form : one warp per (user, column tile); accumulators for the tile live in shared
//                 memory, neighbour row segments are located through a per-row tile index
//                 and loaded 8 neighbours ahead of the accumulation.
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
constexpr int kSpWarps = 4;
constexpr int kSpAhead = 8;   // neighbour segments whose loads are in flight together

struct __align__(16) Acc { double num, den; };

// Two accumulator flavours.
//  F64  : {num, den} in fp64, updated in table order — the reference's arithmetic verbatim.
//  MILLI: similarities are multiples of 0.001 (np.around(., 3), recommender.py:204/270), so
//         N = sum r*milli, D = sum milli and the rater count are exact integers packed in one
//         u64 (N << 32 | cnt << 20 | D).  round(0.7*N/D + 18) is then decided in exact integer
//         arithmetic; it can differ from the reference's fp64 evaluation only when
//         (7N + 180D) / (10D) is EXACTLY half-way between two integers (the fp64 rounding noise
//         of ~1e-12 is far below the next possible distance 1/(20D) >= 7e-7), and exactly those
//         cells are re-evaluated with the reference's sequential fp64 arithmetic.
template <bool MILLI>
struct AccOf { using type = Acc; };
template <>
struct AccOf<true> { using type = unsigned long long; };

// bytes of shared memory one warp needs for a tile: acc[tile] | u16 jlist[tile] | u8 plist[tile] | metadata
__host__ __device__ constexpr size_t sparse_warp_bytes(int tile, bool milli) {
  return round_up<size_t>(size_t(tile) * (milli ? 11 : 19), 16) + 64 * (8 + 8 + 4 + 4);
}

__device__ __forceinline__ int predict_from(double num, double den) {
  return den != 0.0 ? hybrid_rule(0.0, num / den) : 0;
}

// the reference's sequential fp64 evaluation of one cell, re-gathered from the neighbour segments
__device__ __noinline__ int resolve_tie(const int *__restrict__ colidx, const uint8_t *__restrict__ val,
                                        const int64_t *m_a, const int *m_n, const double *m_s, int c, int col) {
  double num = 0.0, den = 0.0;
  for (int t = 0; t < c; ++t) {
    const int64_t a = m_a[t];
    const int n = m_n[t];
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

// One warp per user (grid-stride), walking the column tiles in order so that the per-user state
// (neighbour ids, similarities, segment cursors) is loaded once and every tile needs exactly one
// new position per neighbour — fetched one tile ahead of its use.
template <bool MILLI>
__global__ void __launch_bounds__(kSpWarps * 32, MILLI ? 8 : 6)
predict_sparse_kernel(const int *__restrict__ colidx, const uint8_t *__restrict__ val, int Q,
                      const int *__restrict__ u_idx, const double *__restrict__ u_val,
                      const int *__restrict__ u_milli, const int *__restrict__ u_cnt, int ku, int64_t user_base,
                      int64_t n_users, int tile, int nt, const int64_t *__restrict__ tilepos,
                      int64_t *__restrict__ row_start, int *__restrict__ row_cnt, int *__restrict__ cols,
                      uint8_t *__restrict__ preds, int64_t region_cap, int n_regions,
                      unsigned long long *__restrict__ cursors) {
  using acc_t = typename AccOf<MILLI>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *mine = smem_raw + size_t(warp) * sparse_warp_bytes(tile, MILLI);
  acc_t *acc = reinterpret_cast<acc_t *>(mine);
  uint16_t *jlist = reinterpret_cast<uint16_t *>(acc + tile);
  uint8_t *plist = reinterpret_cast<uint8_t *>(jlist + tile);
  unsigned char *meta = mine + round_up<size_t>(size_t(tile) * (MILLI ? 11 : 19), 16);
  int64_t *m_a = reinterpret_cast<int64_t *>(meta);           // [64] first entry of the neighbour's segment
  double *m_s = reinterpret_cast<double *>(meta + 64 * 8);    // [64] its similarity
  int *m_n = reinterpret_cast<int *>(meta + 64 * 16);         // [64] entries in the segment
  unsigned *m_inc = reinterpret_cast<unsigned *>(meta + 64 * 20);   // [64] MILLI: milli | 1 << 20
  acc_t zero;
  if constexpr (MILLI) zero = 0ull; else zero = Acc{0.0, 0.0};
  for (int j = lane; j < tile; j += 32) acc[j] = zero;
  __syncwarp();

  const int64_t warp_global = int64_t(blockIdx.x) * kSpWarps + warp;
  const int64_t nwarps = int64_t(gridDim.x) * kSpWarps;
  const unsigned lt = lanemask_lt();
  const int stride = nt + 1;
  // output slices are reserved from one of n_regions cursors (a single cursor would serialise
  // millions of same-address atomics in L2); region r owns [r * region_cap, (r + 1) * region_cap)
  const int region = int(warp_global % n_regions);
  unsigned long long *cursor = cursors + region;
  const int64_t region_base = int64_t(region) * region_cap;

  // one contribution (rating r of neighbour slot g) to column slot j
  auto add = [&](int j, int r, int g) {
    if constexpr (MILLI) {
      const unsigned inc = m_inc[g];
      acc[j] += (static_cast<unsigned long long>(unsigned(r) * (inc & 0xFFFFFu)) << 32) | inc;
    } else {
      const double s = m_s[g];
      Acc x = acc[j];
      x.num = __dadd_rn(x.num, __dmul_rn(double(r), s));
      x.den = __dadd_rn(x.den, s);
      acc[j] = x;
    }
  };

  for (int64_t ul = warp_global; ul < n_users; ul += nwarps) {
    const int64_t ug = user_base + ul;            // global user = CSR row
    const int c = min(u_cnt[ul], ku);
    // lane <-> neighbour (slot 0: neighbours 0..31, slot 1: 32..63); cursors into each neighbour's row
    const int64_t *tp[2];
    int64_t cur[2], nxt[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int me = h * 32 + lane;
      tp[h] = tilepos;
      cur[h] = nxt[h] = 0;
      if (me < c) {
        const int v = u_idx[ul * ku + me];
        // MILLI: the table value IS milli/1000.0 (checked by the host), rebuilt here the way numpy does
        if constexpr (MILLI) {
          const int milli = u_milli[ul * ku + me];
          m_s[me] = double(milli) / 1000.0;
          m_inc[me] = unsigned(milli) | (1u << 20);
        } else {
          m_s[me] = u_val[ul * ku + me];
        }
        tp[h] = tilepos + int64_t(v) * stride;
        cur[h] = __ldg(tp[h]);
        nxt[h] = __ldg(tp[h] + 1);
      }
    }
    const int64_t *own_tp = tilepos + ug * stride;
    int64_t own_cur = __ldg(own_tp), own_nxt = __ldg(own_tp + 1);
    __syncwarp();

    for (int t = 0; t < nt; ++t) {
      const int q0 = t * tile;
      const int width = min(tile, Q - q0);
      // this tile's segments = [cur, nxt); fetch the end of the NEXT tile's segments now
      int64_t ahead[2] = {0, 0};
      int64_t own_ahead = 0;
      const bool more = t + 1 < nt;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int me = h * 32 + lane;
        if (me < c) {
          m_a[me] = cur[h];
          m_n[me] = int(nxt[h] - cur[h]);
          if (more) ahead[h] = __ldg(tp[h] + t + 2);
        }
      }
      if (more) own_ahead = __ldg(own_tp + t + 2);
      __syncwarp();

      for (int nb0 = 0; nb0 < c; nb0 += 32) {
        const int nn = min(32, c - nb0);
        // segments longer than one warp-load are the exception (the tile is sized for ~24 entries per
        // row): their first 32 entries take the unrolled path like everyone else, the rest a short loop
        const int my_n = lane < nn ? m_n[nb0 + lane] : 0;
        const unsigned long_mask = __ballot_sync(kFull, my_n > 32);
        for (int g0 = 0; g0 < nn; g0 += kSpAhead) {
          int colv[kSpAhead];
          int ratv[kSpAhead];
#pragma unroll
          for (int g = 0; g < kSpAhead; ++g) {
            ratv[g] = 0;
            colv[g] = 0;
            if (g0 + g < nn) {
              const int64_t a = m_a[nb0 + g0 + g];
              if (lane < m_n[nb0 + g0 + g]) {
                colv[g] = ld_stream_s32(colidx + a + lane);
                ratv[g] = val[a + lane];
              }
            }
          }
#pragma unroll
          for (int g = 0; g < kSpAhead; ++g) {
            if (g0 + g < nn) {     // warp-uniform
              if (ratv[g] != 0) add(colv[g] - q0, ratv[g], nb0 + g0 + g);
              if ((long_mask >> (g0 + g)) & 1u) {
                const int64_t a = m_a[nb0 + g0 + g];
                const int cn = m_n[nb0 + g0 + g];
#pragma unroll 1
                for (int off = 32; off < cn; off += 32) {
                  if (off + lane < cn) {
                    const int r = val[a + off + lane];
                    if (r != 0) add(__ldg(colidx + a + off + lane) - q0, r, nb0 + g0 + g);
                  }
                }
              }
              __syncwarp();
            }
          }
        }
      }

      // ---- the user's own rated cells are not predicted (recommender.py:302): drop what landed there
      for (int64_t x = own_cur + lane; x < own_nxt; x += 32)
        if (val[x] != 0) acc[__ldg(colidx + x) - q0] = zero;
      __syncwarp();

      // ---- pass 1: ascending list of the touched columns
      int touched = 0;
      for (int j0 = 0; j0 < width; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        if (j < width) {
          if constexpr (MILLI) {
            hit = acc[j] != 0ull;
          } else {
            const Acc x = acc[j];
            hit = x.den != 0.0 || x.num != 0.0;
          }
        }
        const unsigned m = __ballot_sync(kFull, hit);
        if (hit) jlist[touched + __popc(m & lt)] = uint16_t(j);
        touched += __popc(m);
      }
      __syncwarp();
      // ---- pass 2: prediction of every touched column (dense lanes), accumulators back to zero
      int total = 0;
      for (int x0 = 0; x0 < touched; x0 += 32) {
        const int x = x0 + lane;
        int p = 0;
        if (x < touched) {
          const int j = jlist[x];
          const acc_t a = acc[j];
          acc[j] = zero;
          if constexpr (MILLI) {
            const unsigned lo = unsigned(a), N = unsigned(a >> 32);
            const unsigned D = lo & 0xFFFFFu, raters = lo >> 20;
            if (D != 0u) {
              const unsigned X = 7u * N + 180u * D, Y = 10u * D;      // x = X / Y = 0.7 * N / D + 18
              const unsigned q = X / Y, rem2 = 2u * (X - q * Y);
              if (rem2 < Y) p = int(q);
              else if (rem2 > Y) p = int(q) + 1;
              else if (raters == 1u) {
                const double s = double(D) / 1000.0;                   // the one rater's similarity
                p = predict_from(__dmul_rn(double(N / D), s), s);
              } else {
                p = resolve_tie(colidx, val, m_a, m_n, m_s, c, q0 + j);
              }
            }
          } else {
            p = predict_from(a.num, a.den);
          }
          plist[x] = uint8_t(p);
        }
        total += __popc(__ballot_sync(kFull, p != 0));
      }
      __syncwarp();
      unsigned long long base = 0;
      if (lane == 0) {
        if (total > 0) base = atomicAdd(cursor, (unsigned long long)total);
        row_start[ul * nt + t] = region_base + int64_t(base);
        row_cnt[ul * nt + t] = total;
      }
      base = __shfl_sync(kFull, base, 0);
      // ---- pass 3: the slice, ascending in column
      if (total > 0 && int64_t(base) < region_cap) {
        int written = 0;
        for (int x0 = 0; x0 < touched; x0 += 32) {
          const int x = x0 + lane;
          const int p = x < touched ? int(plist[x]) : 0;
          const unsigned m = __ballot_sync(kFull, p != 0);
          if (p != 0) {
            const int64_t pos = int64_t(base) + written + __popc(m & lt);
            if (pos < region_cap) { cols[region_base + pos] = q0 + int(jlist[x]); preds[region_base + pos] = uint8_t(p); }
          }
          written += __popc(m);
        }
      }
      __syncwarp();
      // advance the cursors to the next tile
#pragma unroll
      for (int h = 0; h < 2; ++h) { cur[h] = nxt[h]; nxt[h] = ahead[h]; }
      own_cur = own_nxt;
      own_nxt = own_ahead;
    }
  }
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

static size_t sparse_smem_bytes(int tile, bool milli) { return size_t(kSpWarps) * sparse_warp_bytes(tile, milli); }

template <bool MILLI>
static int launch_predict_sparse(const int *colidx, const uint8_t *val, int Q, const int *u_idx, const double *u_val,
                                 const int *u_milli, const int *u_cnt, int ku, int64_t user_base, int64_t n_users,
                                 int tile, int nt, const int64_t *tilepos, int64_t *row_start, int *row_cnt, int *cols,
                                 uint8_t *preds, int64_t region_cap, int n_regions, unsigned long long *cursors,
                                 cudaStream_t st) {
  auto kern = predict_sparse_kernel<MILLI>;
  const size_t smem = sparse_smem_bytes(tile, MILLI);
  QRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  QRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSpWarps * 32, smem));
  if (per_sm < 1) return QRS_ERR_UNSUPPORTED;
  int64_t grid = int64_t(sm_count()) * per_sm;
  const int64_t need = ceil_div<int64_t>(n_users, kSpWarps);
  if (grid > need) grid = need;
  kern<<<unsigned(grid), kSpWarps * 32, smem, st>>>(colidx, val, Q, u_idx, u_val, u_milli, u_cnt, ku, user_base,
                                                    n_users, tile, nt, tilepos, row_start, row_cnt, cols, preds,
                                                    region_cap, n_regions, cursors);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // namespace qrs

using namespace qrs;

extern "C" {

int32_t qrs_predict_max_tile(void) { return 2560; }   /* 4 warps x 19 B x 2560 = 195 KB */

size_t qrs_predict_sparse_workspace_bytes(int64_t U, int32_t Q, int32_t tile) {
  if (U < 0 || Q <= 0 || tile <= 0) return 0;
  const int64_t nt = ceil_div<int64_t>(Q, tile);
  return round_up<size_t>(size_t(U > 0 ? U : 1) * size_t(nt + 1) * 8, 256);
}

int qrs_predict_sparse(const int64_t *rowptr, const int32_t *colidx, const uint8_t *val, int64_t U, int32_t Q,
                       const int32_t *u_idx, const double *u_val, const int32_t *u_milli, const int32_t *u_cnt,
                       int32_t ku, int64_t user_base, int64_t n_users, int32_t tile, void *ws, size_t ws_bytes,
                       int64_t *row_start, int32_t *row_cnt, int32_t *cols, uint8_t *preds, int64_t capacity,
                       int32_t n_regions, unsigned long long *cursors, void *stream) {
  if (!rowptr || U < 0 || Q <= 0 || !u_idx || (!u_val && !u_milli) || !u_cnt || ku <= 0 || user_base < 0 ||
      n_users < 0 || user_base + n_users > U || !row_start || !row_cnt || !cursors || capacity < 0 || n_regions <= 0)
    return QRS_ERR_INVALID_ARG;
  const int64_t region_cap = capacity / n_regions;
  if (capacity > 0 && (!cols || !preds)) return QRS_ERR_INVALID_ARG;
  if (tile <= 0 || tile % 32 != 0) return QRS_ERR_INVALID_ARG;
  if (tile > qrs_predict_max_tile() || ku > QRS_MAX_TOPK) return QRS_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < qrs_predict_sparse_workspace_bytes(U, Q, tile)) return QRS_ERR_WORKSPACE;
  if (n_users == 0 || U == 0) return QRS_OK;
  if (!colidx || !val) return QRS_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  const int nt = int(ceil_div<int64_t>(Q, tile));
  int64_t *tilepos = static_cast<int64_t *>(ws);
  {
    const int64_t total = U * (nt + 1);
    int64_t blocks = ceil_div<int64_t>(total, 256);
    const int64_t cap = int64_t(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    tile_index_kernel<<<unsigned(blocks), 256, 0, st>>>(rowptr, colidx, U, tile, nt, tilepos);
    QRS_LAUNCH_CHECK();
  }
  if (u_milli)
    return launch_predict_sparse<true>(colidx, val, Q, u_idx, u_val, u_milli, u_cnt, ku, user_base, n_users, tile, nt,
                                       tilepos, row_start, row_cnt, cols, preds, region_cap, n_regions, cursors, st);
  return launch_predict_sparse<false>(colidx, val, Q, u_idx, u_val, u_milli, u_cnt, ku, user_base, n_users, tile, nt,
                                      tilepos, row_start, row_cnt, cols, preds, region_cap, n_regions, cursors, st);
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
