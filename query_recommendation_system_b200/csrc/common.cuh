// Shared helpers for the libqrs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "qrs_b200.h"

namespace qrs {

void set_cuda_error(cudaError_t e, const char *file, int line);

#define QRS_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) {                                  \
      ::qrs::set_cuda_error(_e, __FILE__, __LINE__);          \
      return QRS_ERR_CUDA;                                    \
    }                                                         \
  } while (0)

// every kernel launch goes through here: error check + the diagnostic launch counter
void note_launch();
#define QRS_LAUNCH_CHECK()            \
  do {                                \
    ::qrs::note_launch();             \
    QRS_CUDA(cudaGetLastError());     \
  } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T>
__host__ __device__ constexpr T round_up(T a, T b) { return ceil_div(a, b) * b; }

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xFFFFFFFFu;

// SM count of the current device, cached per device (grids are sized in
// multiples of it).
int sm_count();

// ---------------------------------------------------------------- bucket records
// hash24 of one band: the r int16 lanes (low 16 bits of the signature values,
// lsh.py:28) taken two at a time as 32-bit words (lane 2k in the low half; a trailing
// odd lane alone) and folded with multiply / xor-shift rounds.  Grouping is by
// (hash24, band); qrs_pairs_score re-verifies every pair on the exact lanes, so
// collisions cost work, never correctness (measured on geometric lanes: the number of
// distinct hashes equals the birthday bound of an ideal 24-bit hash).
__host__ __device__ __forceinline__ uint32_t band_hash_init(uint32_t band) {
  return 0x9E3779B9u * (band + 1u);
}
__host__ __device__ __forceinline__ uint32_t band_hash_word(uint32_t h, uint32_t word) {
  h = (h ^ word) * 0x9E3779B1u;
  return h ^ (h >> 15);
}
__host__ __device__ __forceinline__ uint32_t band_hash_finish(uint32_t h, bool all_neg) {
  h *= 0x85EBCA6Bu;
  h ^= h >> 16;
  h >>= (32 - QRS_HASH_BITS);
  if (h == QRS_SENTINEL_HASH) h = QRS_SENTINEL_HASH - 1u;
  return all_neg ? QRS_SENTINEL_HASH : h;
}
__host__ __device__ __forceinline__ uint64_t make_record(uint32_t hash24, uint32_t band, uint32_t id) {
  return (uint64_t(hash24) << 40) | (uint64_t(band & 0xFFu) << 32) | uint64_t(id);
}
__host__ __device__ __forceinline__ uint32_t rec_key(uint64_t rec) { return uint32_t(rec >> 32); }
__host__ __device__ __forceinline__ uint32_t rec_hash(uint64_t rec) { return uint32_t(rec >> 40); }
__host__ __device__ __forceinline__ uint32_t rec_band(uint64_t rec) { return uint32_t(rec >> 32) & 0xFFu; }
__host__ __device__ __forceinline__ uint32_t rec_id(uint64_t rec) { return uint32_t(rec); }

// record of (set, band) from r consecutive signature values (int rows, or the 16-bit rows of the tile scan)
template <typename T>
__device__ __forceinline__ uint64_t band_record(const T *lanes, int r, uint32_t band, uint32_t id) {
  uint32_t h = band_hash_init(band);
  bool all_neg = true;
  for (int c = 0; c < r; c += 2) {
    uint32_t w = uint32_t(lanes[c]) & 0xFFFFu;
    all_neg = all_neg && (w == 0xFFFFu);
    if (c + 1 < r) {
      const uint32_t hi = uint32_t(lanes[c + 1]) & 0xFFFFu;
      all_neg = all_neg && (hi == 0xFFFFu);
      w |= hi << 16;
    }
    h = band_hash_word(h, w);
  }
  return make_record(band_hash_finish(h, all_neg), band, id);
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// sum over the warp of values below 2^48: three REDUX.SUM on 16-bit pieces instead of ten 32-bit shuffles
__device__ __forceinline__ unsigned long long warp_sum_u48(unsigned long long v) {
  const unsigned a = __reduce_add_sync(kFull, unsigned(v) & 0xFFFFu);
  const unsigned b = __reduce_add_sync(kFull, unsigned(v >> 16) & 0xFFFFu);
  const unsigned c = __reduce_add_sync(kFull, unsigned(v >> 32) & 0xFFFFu);
  return (unsigned long long)a + ((unsigned long long)b << 16) + ((unsigned long long)c << 32);
}
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// streaming (read-once) global loads: keep them out of L1
__device__ __forceinline__ int ld_stream_s32(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// ---------------------------------------------------------------- bulk asynchronous copies (TMA unit, 1-D form)
// global -> shared copies issued by ONE thread and completed on an mbarrier (cp.async.bulk, SASS UBLKCP): the
// data never passes through registers, so a warp can keep many 128-byte lines in flight without register cost.
// Addresses are shared-window byte addresses; sizes are multiples of 16 bytes, both ends 16-byte aligned.
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {      // make the initialised barriers visible to the async proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {   // order generic-proxy accesses before async-proxy ones
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}

// ---------------------------------------------------------------- device-wide scan
// Exclusive prefix sum of u32 counts into u64 offsets (out[n] = total).
// ws: scan_workspace_bytes(n).  Three launches: tile sums, one-CTA scan of the
// tile sums, tile scan + carry-in.
size_t scan_workspace_bytes(int64_t n);
int exclusive_scan_u32_to_u64(const uint32_t *in, uint64_t *out, int64_t n, void *ws,
                              size_t ws_bytes, cudaStream_t stream);

}  // namespace qrs
