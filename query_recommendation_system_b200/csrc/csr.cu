// CSR plumbing kernels for the multi-GPU prediction: pack the ratings rows other ranks asked for
// into contiguous send buffers (a10/a11 gather neighbour rows of R, recommender.py:320; with rows
// sharded across GPUs the remote ones have to be shipped first).
#include "common.cuh"

namespace qrs {

// warp per requested row: out[out_off[t] .. out_off[t+1]) = row rows[t]
__global__ void __launch_bounds__(256)
gather_rows_kernel(const int64_t *__restrict__ rowptr, const int *__restrict__ colidx, const uint8_t *__restrict__ val,
                   const int64_t *__restrict__ rows, int64_t n_rows, const int64_t *__restrict__ out_off,
                   int *__restrict__ out_col, uint8_t *__restrict__ out_val) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t t = warp_global; t < n_rows; t += nwarps) {
    const int64_t r = rows[t];
    const int64_t src = rowptr[r], len = rowptr[r + 1] - src, dst = out_off[t];
    for (int64_t x = lane; x < len; x += 32) {
      out_col[dst + x] = ld_stream_s32(colidx + src + x);
      out_val[dst + x] = val[src + x];
    }
  }
}

}  // namespace qrs

using namespace qrs;

extern "C" int qrs_gather_rows(const int64_t *rowptr, const int32_t *colidx, const uint8_t *val, const int64_t *rows,
                               int64_t n_rows, const int64_t *out_off, int32_t *out_col, uint8_t *out_val,
                               void *stream) {
  if (n_rows < 0 || !rowptr || (n_rows > 0 && (!rows || !out_off))) return QRS_ERR_INVALID_ARG;
  if (n_rows == 0) return QRS_OK;
  if (!colidx || !val || !out_col || !out_val) return QRS_ERR_INVALID_ARG;
  int64_t blocks = ceil_div<int64_t>(n_rows, 8);
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  gather_rows_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(rowptr, colidx, val, rows, n_rows, out_off,
                                                                     out_col, out_val);
  QRS_LAUNCH_CHECK();
  return QRS_OK;
}
