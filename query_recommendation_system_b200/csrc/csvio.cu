// f-4: the wire format on the input side of the path.  utility_matrix.csv as
// written by the reference's generator (generator.py:59-77, 329-332: csv.writer, "\r\n" rows,
// header = query ids WITHOUT a user field, then `user,r,r,,r,...` with an empty field = not rated,
// integers 1..100) is parsed on the GPU straight into the ratings CSR the path consumes — what
// main.py:75 does with dt.fread into a dense frame.
//
//   qrs_csv_line_ends      positions of every '\n' (chunk counts -> scan -> ordered write)
//   qrs_csv_ratings_count  per row: rated cells, end of the user-id field, format checks
//   qrs_csv_ratings_fill   column ids + rating bytes at the scanned offsets
//
#include "common.cuh"

namespace qrs {

constexpr int kCsvWarpBytes = 4096;        // bytes one warp scans for line ends (8 sweeps of 512)

__device__ __forceinline__ uint4 load16_guarded(const uint8_t *text, int64_t p, int64_t n) {
  if (p + 16 <= n) return *reinterpret_cast<const uint4 *>(text + p);
  unsigned w[4] = {0u, 0u, 0u, 0u};
  for (int k = 0; k < 16; ++k)
    if (p + k < n) w[k >> 2] |= unsigned(text[p + k]) << (8 * (k & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// bit k of the result = byte k of the 16 loaded bytes is '\n'
__device__ __forceinline__ unsigned newline_bits(uint4 v) {
  const unsigned w[4] = {v.x, v.y, v.z, v.w};
  unsigned bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned eq = __vcmpeq4(w[i], 0x0A0A0A0Au);           // 0xFF per matching byte
    bits |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * i);
  }
  return bits;
}

// WRITE = false: counts[chunk] = newlines in the chunk.  WRITE = true: line_end[offsets[chunk] + k] = position
// of the chunk's k-th newline.  `text` is 16-byte aligned (checked by the entry point).
template <bool WRITE>
__global__ void __launch_bounds__(256)
csv_newlines_kernel(const uint8_t *__restrict__ text, int64_t n, int64_t n_chunks, uint32_t *__restrict__ counts,
                    const uint64_t *__restrict__ offsets, int64_t *__restrict__ line_end, int64_t capacity) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t ch = warp_global; ch < n_chunks; ch += nwarps) {
    const int64_t base = ch * kCsvWarpBytes;
    int64_t out = WRITE ? int64_t(offsets[ch]) : 0;
    unsigned total = 0;
#pragma unroll 2
    for (int s = 0; s < kCsvWarpBytes / 512; ++s) {
      const int64_t p = base + s * 512 + lane * 16;
      unsigned bits = p < n ? newline_bits(load16_guarded(text, p, n)) : 0u;
      const unsigned c = __popc(bits);
      if (WRITE) {
        unsigned incl = c;                                       // inclusive prefix of c over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned up = __shfl_up_sync(kFull, incl, d);
          if (lane >= d) incl += up;
        }
        const unsigned before = incl - c;
        int64_t o = out + before;
        while (bits) {
          const int k = __ffs(bits) - 1;
          bits &= bits - 1;
          if (o < capacity) line_end[o] = p + k;
          ++o;
        }
      }
      unsigned sum = c;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
      out += sum;
      total += sum;
    }
    if (!WRITE && lane == 0) counts[ch] = total;
  }
}

// exact per-byte tests on a 32-bit word: bit 7 of every byte that matches
__device__ __forceinline__ unsigned bytes_equal(unsigned w, unsigned pattern) {
  const unsigned x = w ^ pattern;
  return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
__device__ __forceinline__ unsigned bytes_digit(unsigned w) {
  const unsigned m7 = w & 0x7F7F7F7Fu;
  const unsigned ge30 = m7 + 0x50505050u, ge3a = m7 + 0x46464646u;   // bit 7 <=> low 7 bits >= '0' / > '9'
  return ge30 & ~ge3a & ~w & 0x80808080u;
}

// exclusive prefix over the lanes (and the warp total) of a per-lane count in 0..7: three bit-plane ballots
__device__ __forceinline__ int prefix_small(int c, unsigned lt, int &total) {
  const unsigned b0 = __ballot_sync(kFull, c & 1), b1 = __ballot_sync(kFull, c & 2), b2 = __ballot_sync(kFull, c & 4);
  total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
  return __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
}

// the aligned word at position a of a row [begin, end]: bytes outside the row are blanked, the byte at `end`
// (the '\r' / '\n' / end of text that closes the last field) reads as one more ','; vm flags the bytes inside
__device__ __forceinline__ unsigned row_word(const uint8_t *__restrict__ text, int64_t n_bytes, int64_t a, int64_t begin,
                                             int64_t end, unsigned &vm) {
  vm = 0u;
  if (a > end || a + 3 < begin) return 0u;
  unsigned w = 0u;
  if (a + 4 <= n_bytes) {
    w = *reinterpret_cast<const unsigned *>(text + a);
  } else {
    for (int k = 0; k < 4; ++k)
      if (a + k < n_bytes) w |= unsigned(text[a + k]) << (8 * k);
  }
  vm = 0x80808080u;
  if (a < begin || a + 3 >= end) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t pos = a + k;
      const unsigned m = 0xFFu << (8 * k);
      if (pos < begin || pos > end) { w &= ~m; vm &= ~m; }
      else if (pos == end) w = (w & ~m) | (unsigned(',') << (8 * k));
    }
  }
  return w;
}

// A warp owns a row and walks it 128 bytes per step (one aligned word per lane).  Comma and digit flags
// come from word-wide byte tests; a comma right after a digit closes a rated cell, and only those (a few
// per step) are parsed: the integer is read from the 4 bytes in front (own word + the previous lane's),
// its column is the number of commas before it (per-lane counts -> bit-plane ballot prefix).  Emission
// order is ascending by construction, so rows come out sorted.  The counting pass needs neither prefix:
// it sums per lane and reduces once per row (only the steps up to the first comma, where the user id may
// hold digits, take the indexed path).
template <bool FILL>
__global__ void __launch_bounds__(256)
csv_ratings_kernel(const uint8_t *__restrict__ text, int64_t n_bytes, const int64_t *__restrict__ line_end,
                   int64_t first_line, int64_t n_rows, int Q, uint32_t *__restrict__ row_cnt,
                   int64_t *__restrict__ id_end, int *__restrict__ err, const int64_t *__restrict__ rowptr,
                   int *__restrict__ colidx, uint8_t *__restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const unsigned lt = lanemask_lt();
  for (int64_t r = warp_global; r < n_rows; r += nwarps) {
    const int64_t L = first_line + r;
    const int64_t begin = L == 0 ? 0 : line_end[L - 1] + 1;
    int64_t end = line_end[L];                                 // the '\n' (or one past the last byte of the text)
    if (end > begin && text[end - 1] == '\r') --end;           // csv.writer's "\r\n"
    const int64_t out_base = FILL ? rowptr[r] : 0;
    int fields = 0, emitted = 0;                               // FILL (and every row's first steps): warp-uniform totals
    int my_fields = 0, my_emitted = 0;                         // counting pass: per-lane sums, reduced at the end
    unsigned prev_w = 0u, prev_dg = 0u;                        // lane 31's word / digit flags of the previous step
    int64_t first_term = -1;
    bool bad = false;
    const int64_t a0 = begin & ~int64_t(3);
    unsigned vm_next;
    unsigned w_next = row_word(text, n_bytes, a0 + 4 * lane, begin, end, vm_next);
    for (int64_t s = a0; s <= end; s += 128) {
      const unsigned w = w_next, vm = vm_next;
      w_next = row_word(text, n_bytes, s + 128 + 4 * lane, begin, end, vm_next);   // in flight during this step
      const unsigned cm = bytes_equal(w, 0x2C2C2C2Cu) & vm;
      const unsigned dg = bytes_digit(w);
      const unsigned other = vm & ~(cm | dg);
      unsigned pw = __shfl_up_sync(kFull, w, 1), pdg = __shfl_up_sync(kFull, dg, 1);
      if (lane == 0) { pw = prev_w; pdg = prev_dg; }
      prev_w = __shfl_sync(kFull, w, 31);
      prev_dg = __shfl_sync(kFull, dg, 31);
      // commas whose preceding byte is a digit: the cells that hold a number
      unsigned cand = cm & unsigned((((unsigned long long)dg << 32) | pdg) >> 24);
      const bool indexed = FILL || first_term < 0;             // warp-uniform
      int f0 = 1, ctotal = 0;
      if (indexed) {
        f0 = fields + prefix_small(__popc(cm), lt, ctotal);    // field of this lane's first byte
        if (first_term < 0 && ctotal > 0) {                    // the comma closing the user id
          const int src = __ffs(__ballot_sync(kFull, cm != 0u)) - 1;
          const unsigned cm_src = __shfl_sync(kFull, cm, src);
          first_term = s + 4 * src + ((__ffs(cm_src) - 1) >> 3);
        }
        fields += ctotal;
      } else {
        my_fields += __popc(cm);
      }
      if (other != 0u) {                                       // user-id text, or something that does not belong
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((other >> (8 * k + 7)) & 1u) {
            const int fk = indexed ? f0 + __popc(cm & ((1u << (8 * k)) - 1u)) : 1;
            if (fk >= 1 || ((w >> (8 * k)) & 0xFFu) == unsigned('"')) bad = true;
          }
      }
      const unsigned long long win = (unsigned long long)pw | ((unsigned long long)w << 32);
      int v4[4], c4[4];
      unsigned emit4 = 0u;
      while (cand != 0u) {                                     // a few lanes per step, one or two turns each
        const int bit = __ffs(cand) - 1, k = bit >> 3;
        cand &= cand - 1u;
        const int fk = indexed ? f0 + __popc(cm & ((1u << (8 * k)) - 1u)) : 1;
        if (fk < 1) continue;                                  // digits at the end of the user id
        const unsigned before = unsigned(win >> (8 * k));      // the 4 bytes in front of byte k, nearest in bits 24-31
        const unsigned d1 = (before >> 24) - '0', d2 = ((before >> 16) & 0xFFu) - '0',
                       d3 = ((before >> 8) & 0xFFu) - '0', d4 = (before & 0xFFu) - '0';
        int v = int(d1);
        if (d2 < 10u) {
          v += 10 * int(d2);
          if (d3 < 10u) {
            v += 100 * int(d3);
            if (d4 < 10u) bad = true;                          // four digits or more
          }
        }
        if (v > 255) bad = true;
        if (v != 0) {                                          // 0 = not rated, like an empty field
          emit4 |= 1u << k;
          if (FILL) {
            if (k == 0) { v4[0] = v; c4[0] = fk - 1; }
            else if (k == 1) { v4[1] = v; c4[1] = fk - 1; }
            else if (k == 2) { v4[2] = v; c4[2] = fk - 1; }
            else { v4[3] = v; c4[3] = fk - 1; }
          }
        }
      }
      if (FILL) {
        int etotal;
        const int ebefore = prefix_small(__popc(emit4), lt, etotal);
        if (emit4 != 0u) {
          int64_t o = out_base + emitted + ebefore;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((emit4 >> k) & 1u) {
              colidx[o] = c4[k];
              val[o] = uint8_t(v4[k]);
              ++o;
            }
        }
        emitted += etotal;
      } else {
        my_emitted += __popc(emit4);
      }
    }
    if (!FILL) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        my_fields += __shfl_xor_sync(kFull, my_fields, o);
        my_emitted += __shfl_xor_sync(kFull, my_emitted, o);
      }
      fields += my_fields;
      emitted = my_emitted;
    }
    bad = __any_sync(kFull, bad) || fields != Q + 1;
    if (!FILL && lane == 0) {
      row_cnt[r] = bad ? 0u : unsigned(emitted);
      id_end[r] = first_term < 0 ? end : first_term;
      if (bad) {
        atomicAdd(err, 1);
        atomicMin(err + 1, int(r < 0x7FFFFFFF ? r : 0x7FFFFFFE));
      }
    }
  }
}

}  // namespace qrs

using namespace qrs;

namespace {
// compact CSR wire format: 16-bit column ids (m <= 65536) widened on the device, 16 bytes in / 32 bytes out per thread
__global__ void __launch_bounds__(256) widen_u16_kernel(const uint16_t *__restrict__ in, int64_t n, int *__restrict__ out) {
  const int64_t n8 = n >> 3;
  const bool vec = ((reinterpret_cast<uintptr_t>(in) & 15) | (reinterpret_cast<uintptr_t>(out) & 15)) == 0;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = int64_t(gridDim.x) * blockDim.x;
  if (vec) {
    for (int64_t i = tid; i < n8; i += nth) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in) + i);
      int4 *o = reinterpret_cast<int4 *>(out) + 2 * i;
      o[0] = make_int4(int(v.x & 0xFFFFu), int(v.x >> 16), int(v.y & 0xFFFFu), int(v.y >> 16));
      o[1] = make_int4(int(v.z & 0xFFFFu), int(v.z >> 16), int(v.w & 0xFFFFu), int(v.w >> 16));
    }
    for (int64_t i = (n8 << 3) + tid; i < n; i += nth) out[i] = in[i];
  } else {
    for (int64_t i = tid; i < n; i += nth) out[i] = in[i];
  }
}
}  // namespace

extern "C" {

int qrs_cols_widen_u16(const uint16_t *cols16, int64_t n, int32_t *colidx, void *stream) {
  if (n < 0 || (n > 0 && (!cols16 || !colidx))) return QRS_ERR_INVALID_ARG;
  if (n == 0) return QRS_OK;
  int64_t blocks = ceil_div<int64_t>(ceil_div<int64_t>(n, 8), 256);
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  widen_u16_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(cols16, n, colidx);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}


size_t qrs_csv_workspace_bytes(int64_t n_bytes, int64_t n_rows) {
  if (n_bytes < 0 || n_rows < 0) return 0;
  const int64_t chunks = ceil_div<int64_t>(n_bytes > 0 ? n_bytes : 1, kCsvWarpBytes);
  const int64_t entries = chunks > n_rows ? chunks : n_rows;
  return round_up<size_t>(size_t(entries) * 4, 256) + round_up<size_t>(size_t(chunks + 1) * 8, 256) +
         scan_workspace_bytes(entries);
}

int qrs_csv_line_ends(const uint8_t *text, int64_t n_bytes, int64_t *line_end, int64_t capacity, uint64_t *n_lines,
                      void *ws, size_t ws_bytes, void *stream) {
  if (n_bytes < 0 || capacity < 0 || !n_lines || (n_bytes > 0 && !text) || (capacity > 0 && !line_end))
    return QRS_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(text) & 15u) != 0) return QRS_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  if (n_bytes == 0) {
    QRS_CUDA(cudaMemsetAsync(n_lines, 0, sizeof(uint64_t), st));
    return QRS_OK;
  }
  if (!ws || ws_bytes < qrs_csv_workspace_bytes(n_bytes, 0)) return QRS_ERR_WORKSPACE;
  const int64_t chunks = ceil_div<int64_t>(n_bytes, kCsvWarpBytes);
  char *w = static_cast<char *>(ws);
  uint32_t *counts = reinterpret_cast<uint32_t *>(w);
  w += round_up<size_t>(size_t(chunks) * 4, 256);
  uint64_t *offsets = reinterpret_cast<uint64_t *>(w);
  w += round_up<size_t>(size_t(chunks + 1) * 8, 256);
  int64_t blocks = ceil_div<int64_t>(chunks, 8);
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  csv_newlines_kernel<false><<<unsigned(blocks), 256, 0, st>>>(text, n_bytes, chunks, counts, nullptr, nullptr, 0);
  QRS_LAUNCH_CHECK();
  int rc = exclusive_scan_u32_to_u64(counts, offsets, chunks, w, scan_workspace_bytes(chunks), st);
  if (rc != QRS_OK) return rc;
  QRS_CUDA(cudaMemcpyAsync(n_lines, offsets + chunks, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  if (capacity > 0) {
    csv_newlines_kernel<true><<<unsigned(blocks), 256, 0, st>>>(text, n_bytes, chunks, nullptr, offsets, line_end,
                                                               capacity);
    QRS_LAUNCH_CHECK();
  }
  return QRS_OK;
}

int qrs_csv_ratings_count(const uint8_t *text, int64_t n_bytes, const int64_t *line_end, int64_t first_line,
                          int64_t n_rows, int32_t Q, int64_t *rowptr, int64_t *id_end, int32_t *err, void *ws,
                          size_t ws_bytes, void *stream) {
  if (n_rows < 0 || n_bytes < 0 || first_line < 0 || Q <= 0 || !rowptr || !err) return QRS_ERR_INVALID_ARG;
  if (n_rows > 0 && (!text || !line_end || !id_end)) return QRS_ERR_INVALID_ARG;
  if (!ws || ws_bytes < qrs_csv_workspace_bytes(0, n_rows)) return QRS_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  uint32_t *counts = static_cast<uint32_t *>(ws);
  const int64_t chunks1 = 1;      // layout of qrs_csv_workspace_bytes(0, n_rows): counts | 1-chunk offsets | scan
  char *w = static_cast<char *>(ws) + round_up<size_t>(size_t(n_rows > 1 ? n_rows : 1) * 4, 256) +
            round_up<size_t>(size_t(chunks1 + 1) * 8, 256);
  const int init[2] = {0, 0x7FFFFFFF};
  QRS_CUDA(cudaMemcpyAsync(err, init, sizeof(init), cudaMemcpyHostToDevice, st));
  if (n_rows > 0) {
    int64_t blocks = ceil_div<int64_t>(n_rows, 8);
    const int64_t cap = int64_t(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    csv_ratings_kernel<false><<<unsigned(blocks), 256, 0, st>>>(text, n_bytes, line_end, first_line, n_rows, Q, counts, id_end,
                                                               err, nullptr, nullptr, nullptr);
    QRS_LAUNCH_CHECK();
  }
  return exclusive_scan_u32_to_u64(counts, reinterpret_cast<uint64_t *>(rowptr), n_rows, w,
                                   scan_workspace_bytes(n_rows), st);
}

int qrs_csv_ratings_fill(const uint8_t *text, int64_t n_bytes, const int64_t *line_end, int64_t first_line,
                         int64_t n_rows, int32_t Q, const int64_t *rowptr, int32_t *colidx, uint8_t *val, void *stream) {
  if (n_rows < 0 || n_bytes < 0 || first_line < 0 || Q <= 0 || !rowptr) return QRS_ERR_INVALID_ARG;
  if (n_rows == 0) return QRS_OK;
  if (!text || !line_end || !colidx || !val) return QRS_ERR_INVALID_ARG;
  int64_t blocks = ceil_div<int64_t>(n_rows, 8);
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  csv_ratings_kernel<true><<<unsigned(blocks), 256, 0, as_stream(stream)>>>(text, n_bytes, line_end, first_line, n_rows, Q,
                                                                           nullptr, nullptr, nullptr, rowptr, colidx,
                                                                           val);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// The output side: final_predictions.csv (main.py:97-111 -> generator.write_csv, generator.py:59-77:
// csv.writer, header = query ids, rows `user,v,v,...` + "\r\n") and, with zero_as_empty, the
// utility-matrix form itself (generator.py:329-332).  A row is the union of two sparse layers: the
// ratings CSR and (optionally) the sparse predictions of qrs_predict_sparse — they never share a cell
// (recommender.py:302 predicts only unrated cells).
namespace qrs {

constexpr int kFmtThreads = 256;
constexpr int kFmtTile = 4096;                     // columns formatted per sweep: at most 16 KB of text
constexpr int kFmtPerThread = kFmtTile / kFmtThreads;
constexpr int kFmtMaxQ = 192 * 1024;

__device__ __forceinline__ int digits_of(int v) { return v < 10 ? 1 : (v < 100 ? 2 : 3); }

// warp per row: bytes of the row's text
__global__ void __launch_bounds__(256)
csv_format_count_kernel(const int64_t *__restrict__ ids_off, const int64_t *__restrict__ rowptr,
                        const uint8_t *__restrict__ val, const int64_t *__restrict__ p_start,
                        const int *__restrict__ p_cnt, const uint8_t *__restrict__ preds, int nt, int64_t n_rows, int Q,
                        int zero_as_empty, uint32_t *__restrict__ row_bytes) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp_global; r < n_rows; r += nwarps) {
    int dig = 0, cells = 0;
    for (int64_t x = rowptr[r] + lane; x < rowptr[r + 1]; x += 32) {
      const int v = val[x];
      if (v != 0) { dig += digits_of(v); ++cells; }
    }
    if (p_start != nullptr)
      for (int t = 0; t < nt; ++t) {
        const int64_t a = p_start[r * nt + t];
        const int c = p_cnt[r * nt + t];
        for (int x = lane; x < c; x += 32) {
          const int v = preds[a + x];
          if (v != 0) { dig += digits_of(v); ++cells; }
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dig += __shfl_xor_sync(kFull, dig, o);
      cells += __shfl_xor_sync(kFull, cells, o);
    }
    if (lane == 0) {
      const int64_t id_len = ids_off[r + 1] - ids_off[r];
      row_bytes[r] = unsigned(id_len + Q + 2 + dig + (zero_as_empty ? 0 : Q - cells));
    }
  }
}

// CTA per row: the row's values by column in shared memory, then the text tile by tile.  A thread formats 16
// consecutive columns (one 16-byte shared load); the tile's text is staged in shared memory at the same
// 16-byte phase as its place in the output, so it leaves in aligned 16-byte stores.
__global__ void __launch_bounds__(kFmtThreads)
csv_format_fill_kernel(const uint8_t *__restrict__ ids, const int64_t *__restrict__ ids_off,
                       const int64_t *__restrict__ rowptr, const int *__restrict__ colidx,
                       const uint8_t *__restrict__ val, const int64_t *__restrict__ p_start,
                       const int *__restrict__ p_cnt, const void *__restrict__ p_cols, int cols_u16,
                       const uint8_t *__restrict__ preds, int nt, int64_t n_rows, int Q, int q_pad, int zero_as_empty,
                       const int64_t *__restrict__ text_off, uint8_t *__restrict__ text) {
  extern __shared__ __align__(16) unsigned char fmt_smem[];
  uint8_t *dense = fmt_smem;                                  // [q_pad], q_pad % 16 == 0 and >= Q rounded up to a tile
  uint8_t *tile_text = fmt_smem + q_pad;                      // [4 * kFmtTile + 16]
  __shared__ int warp_sum[kFmtThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
    for (int x = tid; x < q_pad / 16; x += kFmtThreads) reinterpret_cast<uint4 *>(dense)[x] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    for (int64_t x = rowptr[r] + tid; x < rowptr[r + 1]; x += kFmtThreads) dense[colidx[x]] = val[x];
    if (p_start != nullptr)
      for (int t = 0; t < nt; ++t) {
        const int64_t a = p_start[r * nt + t];
        const int c = p_cnt[r * nt + t];
        for (int x = tid; x < c; x += kFmtThreads) {
          const int col = cols_u16 ? int(static_cast<const uint16_t *>(p_cols)[a + x])
                                   : static_cast<const int *>(p_cols)[a + x];
          dense[col] = preds[a + x];
        }
      }
    int64_t out = text_off[r];
    const int64_t id_a = ids_off[r], id_len = ids_off[r + 1] - id_a;
    for (int64_t x = tid; x < id_len; x += kFmtThreads) text[out + x] = ids[id_a + x];
    out += id_len;
    __syncthreads();
    for (int q0 = 0; q0 < Q; q0 += kFmtTile) {
      // this thread's 16 consecutive columns: ',' + digits (or nothing for an empty cell)
      const int c0 = q0 + tid * kFmtPerThread;
      uint4 dv = make_uint4(0u, 0u, 0u, 0u);
      if (c0 < Q) dv = *reinterpret_cast<const uint4 *>(dense + c0);
      const unsigned dw[4] = {dv.x, dv.y, dv.z, dv.w};
      const int ncol = min(kFmtPerThread, Q - c0);            // <= 0 beyond the row
      int len = 0;
#pragma unroll
      for (int u = 0; u < kFmtPerThread; ++u) {
        if (u < ncol) {
          const int v = int((dw[u >> 2] >> (8 * (u & 3))) & 0xFFu);
          len += 1 + (v == 0 ? (zero_as_empty ? 0 : 1) : digits_of(v));
        }
      }
      int incl = len;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += up;
      }
      if (lane == 31) warp_sum[warp] = incl;
      __syncthreads();
      int before = incl - len, total = 0;
#pragma unroll
      for (int w = 0; w < kFmtThreads / 32; ++w) {
        if (w < warp) before += warp_sum[w];
        total += warp_sum[w];
      }
      const int phase = int(out & 15);
      uint8_t *o = tile_text + phase + before;
#pragma unroll
      for (int u = 0; u < kFmtPerThread; ++u) {
        if (u < ncol) {
          const int v = int((dw[u >> 2] >> (8 * (u & 3))) & 0xFFu);
          *o++ = ',';
          if (v >= 100) *o++ = uint8_t('0' + v / 100);
          if (v >= 10) *o++ = uint8_t('0' + (v / 10) % 10);
          if (v != 0 || !zero_as_empty) *o++ = uint8_t('0' + v % 10);
        }
      }
      __syncthreads();
      // buffer byte i <-> text[out - phase + i]; whole 16-byte vectors in the middle, bytes at the two ends
      uint8_t *dst = text + (out - phase);
      const int stop = phase + total;
      const int v_lo = phase == 0 ? 0 : 1, v_hi = stop >> 4;  // full vectors [v_lo, v_hi)
      for (int j = v_lo + tid; j < v_hi; j += kFmtThreads)
        reinterpret_cast<uint4 *>(dst)[j] = reinterpret_cast<const uint4 *>(tile_text)[j];
      if (tid < 16) {
        if (phase != 0 && tid >= phase && tid < stop) dst[tid] = tile_text[tid];          // head of the first vector
        const int x = (v_hi << 4) + tid;                                                  // tail after the last full one
        if (x < stop && (v_hi >= v_lo) && x >= 16 * v_lo) dst[x] = tile_text[x];
      }
      out += total;
      __syncthreads();
    }
    if (tid == 0) { text[out] = '\r'; text[out + 1] = '\n'; }
  }
}

}  // namespace qrs

extern "C" {

size_t qrs_csv_format_workspace_bytes(int64_t n_rows) {
  if (n_rows < 0) return 0;
  return round_up<size_t>(size_t(n_rows > 0 ? n_rows : 1) * 4, 256) + scan_workspace_bytes(n_rows);
}

int qrs_csv_format_count(const int64_t *ids_off, const int64_t *rowptr, const uint8_t *val, const int64_t *pred_start,
                         const int32_t *pred_cnt, const uint8_t *preds, int32_t nt, int64_t n_rows, int32_t Q,
                         int32_t zero_as_empty, int64_t *text_off, void *ws, size_t ws_bytes, void *stream) {
  if (n_rows < 0 || Q <= 0 || !text_off || !ids_off || !rowptr) return QRS_ERR_INVALID_ARG;
  if (pred_start && (!pred_cnt || !preds || nt <= 0)) return QRS_ERR_INVALID_ARG;
  if (Q > kFmtMaxQ) return QRS_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < qrs_csv_format_workspace_bytes(n_rows)) return QRS_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  uint32_t *counts = static_cast<uint32_t *>(ws);
  void *scan_ws = static_cast<char *>(ws) + round_up<size_t>(size_t(n_rows > 0 ? n_rows : 1) * 4, 256);
  if (n_rows > 0) {
    int64_t blocks = ceil_div<int64_t>(n_rows, 8);
    const int64_t cap = int64_t(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    csv_format_count_kernel<<<unsigned(blocks), 256, 0, st>>>(ids_off, rowptr, val, pred_start, pred_cnt, preds, nt,
                                                             n_rows, Q, zero_as_empty, counts);
    QRS_LAUNCH_CHECK();
  }
  return exclusive_scan_u32_to_u64(counts, reinterpret_cast<uint64_t *>(text_off), n_rows, scan_ws,
                                   scan_workspace_bytes(n_rows), st);
}

int qrs_csv_format_fill(const uint8_t *ids, const int64_t *ids_off, const int64_t *rowptr, const int32_t *colidx,
                        const uint8_t *val, const int64_t *pred_start, const int32_t *pred_cnt, const void *pred_cols,
                        int32_t cols_mode, const uint8_t *preds, int32_t nt, int64_t n_rows, int32_t Q,
                        int32_t zero_as_empty, const int64_t *text_off, uint8_t *text, void *stream) {
  if (n_rows < 0 || Q <= 0 || !text_off || !ids_off || !rowptr) return QRS_ERR_INVALID_ARG;
  if (n_rows == 0) return QRS_OK;
  if (!text || !ids || (reinterpret_cast<uintptr_t>(text) & 15u) != 0) return QRS_ERR_INVALID_ARG;
  if (pred_start && (!pred_cnt || !preds || !pred_cols || nt <= 0)) return QRS_ERR_INVALID_ARG;
  if (cols_mode != QRS_COLS_I32 && cols_mode != QRS_COLS_U16) return QRS_ERR_UNSUPPORTED;
  if (Q > kFmtMaxQ) return QRS_ERR_UNSUPPORTED;
  const int q_pad = round_up(Q, 16);
  const size_t smem = size_t(q_pad) + 4 * kFmtTile + 16;
  QRS_CUDA(cudaFuncSetAttribute(csv_format_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 1;
  QRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csv_format_fill_kernel, kFmtThreads, smem));
  if (per_sm < 1) return QRS_ERR_UNSUPPORTED;
  int64_t blocks = n_rows;
  const int64_t cap = int64_t(sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  csv_format_fill_kernel<<<unsigned(blocks), kFmtThreads, smem, as_stream(stream)>>>(
      ids, ids_off, rowptr, colidx, val, pred_start, pred_cnt, pred_cols, cols_mode == QRS_COLS_U16, preds, nt, n_rows, Q,
      q_pad, zero_as_empty, text_off, text);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // extern "C"
