// libqrs_b200: status/error plumbing and the device-wide scan shared by the stages.
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace qrs {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_cuda_error(cudaError_t e, const char *file, int line) {
  const char *base = strrchr(file, '/');
  snprintf(g_err, sizeof(g_err), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e),
           base ? base + 1 : file, line);
  cudaGetLastError();  // clear the sticky-less error so later calls report their own
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

// ------------------------------------------------------------------ scan
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t *total, uint64_t *smem /*>=8*/) {
  // inclusive warp scan
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint64_t t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint64_t w = lane < (kScanThreads / 32) ? smem[lane] : 0;
    uint64_t winc = w;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      uint64_t t = __shfl_up_sync(kFull, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < (kScanThreads / 32)) smem[lane] = winc - w;
    if (lane == (kScanThreads / 32) - 1) smem[kScanThreads / 32] = winc;
  }
  __syncthreads();
  uint64_t res = smem[warp] + inc - v;
  *total = smem[kScanThreads / 32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const uint32_t *__restrict__ in, int64_t n,
                                                               uint64_t *__restrict__ tile_sum) {
  __shared__ uint64_t sm[16];
  const int64_t base = int64_t(blockIdx.x) * kScanTile;
  uint64_t s = 0;
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) {
    int64_t p = base + it * kScanThreads + threadIdx.x;
    if (p < n) s += in[p];
  }
  uint64_t tot;
  block_exclusive_scan(s, &tot, sm);
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads) scan_of_tile_sums(uint64_t *__restrict__ tile_sum, int64_t tiles,
                                                                  uint64_t *__restrict__ grand_total) {
  __shared__ uint64_t sm[16];
  uint64_t carry = 0;
  for (int64_t base = 0; base < tiles; base += kScanThreads) {
    int64_t p = base + threadIdx.x;
    uint64_t v = p < tiles ? tile_sum[p] : 0;
    uint64_t tot;
    uint64_t ex = block_exclusive_scan(v, &tot, sm);
    if (p < tiles) tile_sum[p] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_tiles(const uint32_t *__restrict__ in, int64_t n,
                                                           const uint64_t *__restrict__ tile_off,
                                                           uint64_t *__restrict__ out) {
  __shared__ uint64_t sm[16];
  const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
  uint32_t v[kScanItems];
  uint64_t s = 0;
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) {
    int64_t p = base + it;
    v[it] = p < n ? in[p] : 0u;
    s += v[it];
  }
  uint64_t tot;
  uint64_t ex = block_exclusive_scan(s, &tot, sm) + tile_off[blockIdx.x];
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) {
    int64_t p = base + it;
    if (p < n) out[p] = ex;
    ex += v[it];
  }
}

size_t scan_workspace_bytes(int64_t n) {
  int64_t tiles = ceil_div<int64_t>(n > 0 ? n : 1, kScanTile);
  return round_up<size_t>(size_t(tiles + 1) * sizeof(uint64_t), 256);
}

int exclusive_scan_u32_to_u64(const uint32_t *in, uint64_t *out, int64_t n, void *ws, size_t ws_bytes,
                              cudaStream_t stream) {
  if (n < 0 || !out || (n > 0 && !in) || !ws) return QRS_ERR_INVALID_ARG;
  if (ws_bytes < scan_workspace_bytes(n)) return QRS_ERR_WORKSPACE;
  if (n == 0) {
    QRS_CUDA(cudaMemsetAsync(out, 0, sizeof(uint64_t), stream));
    return QRS_OK;
  }
  int64_t tiles = ceil_div<int64_t>(n, kScanTile);
  uint64_t *tile_sum = static_cast<uint64_t *>(ws);
  scan_tile_sums<<<unsigned(tiles), kScanThreads, 0, stream>>>(in, n, tile_sum);
  QRS_LAUNCH_CHECK();
  scan_of_tile_sums<<<1, kScanThreads, 0, stream>>>(tile_sum, tiles, out + n);
  QRS_LAUNCH_CHECK();
  scan_tiles<<<unsigned(tiles), kScanThreads, 0, stream>>>(in, n, tile_sum, out);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}

}  // namespace qrs

extern "C" {

int qrs_version(void) { return 100; }

unsigned long long qrs_launch_count(void) { return qrs::g_launches.load(std::memory_order_relaxed); }

const char *qrs_status_string(int status) {
  switch (status) {
    case QRS_OK: return "ok";
    case QRS_ERR_INVALID_ARG: return "invalid argument";
    case QRS_ERR_UNSUPPORTED: return "unsupported shape";
    case QRS_ERR_WORKSPACE: return "workspace too small";
    case QRS_ERR_CUDA: return "CUDA runtime error";
    default: return "unknown status";
  }
}

const char *qrs_last_cuda_error(void) { return qrs::g_err; }

int qrs_device_sm_count(int device) {
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return v;
}

}  // extern "C"
