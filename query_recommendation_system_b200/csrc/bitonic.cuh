// Register-resident 64-key bitonic network for one warp (two u64 keys per lane): the "top-k heap" of the
// neighbour-table and recommendation kernels.  Key e lives in lane e % 32, slot e / 32.
#pragma once
#include "common.cuh"

namespace qrs {

// compare-exchange step of the 64-key network (key e = lane + 32*slot)
__device__ __forceinline__ void cx_shfl(uint64_t &x, int j, bool ascending, int lane) {
  const uint64_t other = __shfl_xor_sync(kFull, x, j);
  const bool lower = (lane & j) == 0;
  const uint64_t mn = x < other ? x : other, mx = x < other ? other : x;
  x = (lower == ascending) ? mn : mx;
}
__device__ __forceinline__ void cx_local(uint64_t &a, uint64_t &b, bool ascending) {
  const uint64_t mn = a < b ? a : b, mx = a < b ? b : a;
  a = ascending ? mn : mx;
  b = ascending ? mx : mn;
}
__device__ __forceinline__ void bitonic_sort64(uint64_t &x0, uint64_t &x1, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      // slot 0 holds e = lane, slot 1 holds e = lane + 32; (e & k) is the same for both when k <= 16,
      // and for k == 32 slot 1 runs descending
      const bool asc0 = (lane & k) == 0;
      const bool asc1 = k == 32 ? false : asc0;
      cx_shfl(x0, j, k == 32 ? true : asc0, lane);
      cx_shfl(x1, j, asc1, lane);
    }
  }
  // k = 64: everything ascending; j = 32 is the in-thread exchange
  cx_local(x0, x1, true);
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    cx_shfl(x0, j, true, lane);
    cx_shfl(x1, j, true, lane);
  }
}
// sorts a bitonic 64-sequence ascending
__device__ __forceinline__ void bitonic_merge64(uint64_t &x0, uint64_t &x1, int lane) {
  cx_local(x0, x1, true);
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    cx_shfl(x0, j, true, lane);
    cx_shfl(x1, j, true, lane);
  }
}


// keep the 64 smallest of (held 64, 64 new keys); both lists come in / go out ascending
__device__ __forceinline__ void bitonic_keep_lowest64(uint64_t &h0, uint64_t &h1, uint64_t n0, uint64_t n1, int lane) {
  bitonic_sort64(n0, n1, lane);
  // lowest 64 of two ascending lists: min(h[e], n[63 - e]) is bitonic
  const uint64_t r1 = __shfl_sync(kFull, n1, 31 - lane);   // n[63 - lane]
  const uint64_t r0 = __shfl_sync(kFull, n0, 31 - lane);   // n[31 - lane] = n[63 - (lane + 32)]
  h0 = h0 < r1 ? h0 : r1;
  h1 = h1 < r0 ? h1 : r0;
  bitonic_merge64(h0, h1, lane);
}

// insert one key into the ascending 64-key list (the largest key falls off the end)
__device__ __forceinline__ void sorted_insert64(uint64_t &h0, uint64_t &h1, uint64_t key, int lane) {
  const int pos = __popc(__ballot_sync(kFull, h0 < key)) + __popc(__ballot_sync(kFull, h1 < key));   // keys below it
  const uint64_t up0 = __shfl_up_sync(kFull, h0, 1);
  uint64_t up1 = __shfl_up_sync(kFull, h1, 1);
  const uint64_t carry = __shfl_sync(kFull, h0, 31);       // element 31 moves to element 32
  if (lane == 0) up1 = carry;
  const int e0 = lane, e1 = lane + 32;
  h0 = e0 < pos ? h0 : (e0 == pos ? key : up0);
  h1 = e1 < pos ? h1 : (e1 == pos ? key : up1);
}

}  // namespace qrs
