// Evaluation tail of a candidate sub-network (engine.py:195,222-233 of the reference): hard-label cross entropy and
// top-1 / top-5 hits of one batch of logits, accumulated on the device so a whole validation pass needs ONE device->host read (the
// reference synchronises three times per batch with .item()).
//   kernel 1: one warp per row -> (logsumexp - z[label], number of logits strictly larger than z[label])
//   kernel 2: one CTA, fixed-order reduction (deterministic) -> totals[5] (fp64): sum of per-batch mean losses, top-1 hits, top-5 hits,
//             samples, batches -- exactly the quantities utils.MetricLogger's meters average (loss: n=1 per batch; accuracy: n=batch).
#include "common.cuh"

namespace vsx {
namespace {

constexpr int EM_WARPS = 8;

__global__ void __launch_bounds__(EM_WARPS * 32) eval_rows_kernel(const float* __restrict__ logits, long ld, const long* __restrict__ labels, int rows,
                                                                   int cols, float* __restrict__ row_loss, int* __restrict__ row_rank) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long r = (long)blockIdx.x * EM_WARPS + warp; r < rows; r += (long)gridDim.x * EM_WARPS) {
    const float* x = logits + r * ld;
    const long lab = labels[r];
    const float zt = (lab >= 0 && lab < cols) ? x[lab] : -INFINITY;
    float mx = -INFINITY;
    int above = 0;
    for (int c = lane; c < cols; c += 32) {
      const float v = x[c];
      mx = fmaxf(mx, v);
      above += v > zt;
    }
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < cols; c += 32) se += expf(x[c] - mx);
    se = warp_sum(se);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
    if (lane == 0) {
      row_loss[r] = mx + logf(se) - zt;
      row_rank[r] = above;
    }
  }
}

__global__ void __launch_bounds__(256) eval_reduce_kernel(const float* __restrict__ row_loss, const int* __restrict__ row_rank, int rows,
                                                          double* __restrict__ totals) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double s_loss[256];
  __shared__ int s_top1[256], s_top5[256];
  double l = 0.0;
  int t1 = 0, t5 = 0;
  for (int r = threadIdx.x; r < rows; r += 256) {
    l += (double)row_loss[r];
    t1 += row_rank[r] < 1;
    t5 += row_rank[r] < 5;
  }
  s_loss[threadIdx.x] = l;
  s_top1[threadIdx.x] = t1;
  s_top5[threadIdx.x] = t5;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s_loss[threadIdx.x] += s_loss[threadIdx.x + s];
      s_top1[threadIdx.x] += s_top1[threadIdx.x + s];
      s_top5[threadIdx.x] += s_top5[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    totals[0] += s_loss[0] / (double)rows;
    totals[1] += (double)s_top1[0];
    totals[2] += (double)s_top5[0];
    totals[3] += (double)rows;
    totals[4] += 1.0;
  }
}

}  // namespace
}  // namespace vsx

using namespace vsx;

extern "C" int vsx_eval_metrics(const float* logits, long ld, const long* labels, int rows, int cols, float* row_loss, int* row_rank, double* totals,
                                void* stream) {
  VSX_REQUIRE(cols > 0 && ld >= cols, "vsx_eval_metrics: bad logits shape (cols=%d ld=%ld)", cols, ld);
  if (rows <= 0) return VSX_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  launch_pdl(eval_rows_kernel, dim3(std::min(ceil_div(rows, EM_WARPS), num_sms() * 4)), dim3(EM_WARPS * 32), 0, st, logits, ld, labels, rows, cols, row_loss, row_rank);
  int rc = check_launch("vsx_eval_metrics");
  if (rc) return rc;
  launch_pdl(eval_reduce_kernel, dim3(1), dim3(256), 0, st, row_loss, row_rank, rows, totals);
  return check_launch("vsx_eval_metrics");
}
