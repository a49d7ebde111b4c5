// The two ends of the train step around the model (engine.train_one_epoch, engine.py:152-157 and :175-177):
//   soft-target cross-entropy, forward + gradient in one pass over the logits (timm SoftTargetCrossEntropy, main.py:392-394)
//   multi-tensor AdamW over every parameter in one launch, also refreshing the bf16 operand copies of the weights
//   (timm create_optimizer -> torch.optim.AdamW per-tensor loop, main.py:385).
#include "common.cuh"

namespace vsx {
namespace {

constexpr int CE_WARPS = 8;
constexpr int CE_MAX_PER_LANE = 40;   // up to 1280 classes

// loss_sum += loss_scale * sum_rows( lse*sum(t) - sum(t*x) );  dlogits = grad_scale * (softmax(x)*sum(t) - t)
__global__ void __launch_bounds__(CE_WARPS * 32) soft_ce_kernel(const float* __restrict__ logits, long ld, const float* __restrict__ target,
                                                                 long ldt, int rows, int cols, float loss_scale, float grad_scale,
                                                                 float* __restrict__ loss_sum, float* __restrict__ dlogits, long ldd) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float local = 0.f;
  for (long r = (long)blockIdx.x * CE_WARPS + warp; r < rows; r += (long)gridDim.x * CE_WARPS) {
    const float* x = logits + r * ld;
    const float* t = target + r * ldt;
    float xv[CE_MAX_PER_LANE], tv[CE_MAX_PER_LANE];
    float mx = -INFINITY, st = 0.f, stx = 0.f;
#pragma unroll
    for (int i = 0; i < CE_MAX_PER_LANE; ++i) {
      const int c = i * 32 + lane;
      xv[i] = c < cols ? x[c] : -INFINITY;
      tv[i] = c < cols ? t[c] : 0.f;
      mx = fmaxf(mx, xv[i]);
      st += tv[i];
      stx += c < cols ? tv[i] * xv[i] : 0.f;
    }
    mx = warp_max(mx);
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < CE_MAX_PER_LANE; ++i) {
      xv[i] = (i * 32 + lane) < cols ? expf(xv[i] - mx) : 0.f;
      se += xv[i];
    }
    se = warp_sum(se);
    st = warp_sum(st);
    stx = warp_sum(stx);
    const float lse = mx + logf(se);
    if (lane == 0) local += lse * st - stx;
    if (dlogits != nullptr) {
      const float inv = st / se;
      float* d = dlogits + r * ldd;
#pragma unroll
      for (int i = 0; i < CE_MAX_PER_LANE; ++i) {
        const int c = i * 32 + lane;
        if (c < cols) d[c] = grad_scale * (xv[i] * inv - tv[i]);
      }
    }
  }
  __shared__ float red[CE_WARPS];
  if (lane == 0) red[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < CE_WARPS; ++w) s += red[w];
    atomicAdd(loss_sum, s * loss_scale);
  }
}

__global__ void scale_by_scalar_kernel(float* __restrict__ x, long n, const float* __restrict__ s) {
  pdl_launch_dependents();
  pdl_wait();
  const float v = *s;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) x[i] *= v;
}

constexpr int ADAM_CHUNK = 16384;   // elements per CTA

__global__ void __launch_bounds__(256) adamw_kernel(const vsx_adamw_tensor* __restrict__ tensors, const int* __restrict__ chunk_tensor,
                                                    const int* __restrict__ chunk_index, float lr, float beta1, float beta2, float omb1,
                                                    float omb2, float eps, float bc1, float bc2, const float* __restrict__ grad_scale_dev,
                                                    const float* __restrict__ guard_dev, int* __restrict__ nonfinite_dev) {
  pdl_launch_dependents();
  pdl_wait();
  if (guard_dev != nullptr) {
    // finite-loss guard (engine.py:168-173): a step whose loss is inf / nan leaves parameters, moments, shadows and averages untouched
    // and raises the sticky device flag the host reads at its logging interval -- no host sync on the step itself
    const float l = *guard_dev;
    if (!(fabsf(l) <= 3.402823466e38f)) {
      if (blockIdx.x == 0 && threadIdx.x == 0 && nonfinite_dev != nullptr) atomicAdd(nonfinite_dev, 1);
      return;
    }
  }
  const vsx_adamw_tensor t = tensors[chunk_tensor[blockIdx.x]];
  const long begin = (long)chunk_index[blockIdx.x] * ADAM_CHUNK;
  const long end = begin + ADAM_CHUNK < t.numel ? begin + ADAM_CHUNK : t.numel;
  const float gs = grad_scale_dev != nullptr ? *grad_scale_dev : 1.0f;
  const float decay = 1.0f - lr * t.weight_decay;
  const float step = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  bf16* hi = reinterpret_cast<bf16*>(t.shadow_hi);
  bf16* lo = reinterpret_cast<bf16*>(t.shadow_lo);
  float* ema = t.ema;
  const float ed = t.ema_decay, ed1 = 1.0f - t.ema_decay;
  auto update = [&](float& p, float g, float& m, float& v) {
    g *= gs;
    p *= decay;
    m = beta1 * m + omb1 * g;
    v = beta2 * v + omb2 * g * g;
    p -= step * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
  };
  // 16-byte path: 4 elements per thread per access (28 B/parameter of traffic is all this kernel does)
  const bool vec = (((uintptr_t)t.param | (uintptr_t)t.grad | (uintptr_t)t.exp_avg | (uintptr_t)t.exp_avg_sq | (uintptr_t)ema) & 15) == 0 &&
                   ((uintptr_t)hi & 7) == 0 && ((uintptr_t)lo & 7) == 0;
  const long vend = vec ? begin + ((end - begin) & ~3L) : begin;
  for (long i = begin + 4L * threadIdx.x; i < vend; i += 4 * 256) {
    float4 p = ld4(t.param + i), m = ld4(t.exp_avg + i), v = ld4(t.exp_avg_sq + i);
    const float4 g = ld4(t.grad + i);
    update(p.x, g.x, m.x, v.x);
    update(p.y, g.y, m.y, v.y);
    update(p.z, g.z, m.z, v.z);
    update(p.w, g.w, m.w, v.w);
    st4(t.param + i, p);
    st4(t.exp_avg + i, m);
    st4(t.exp_avg_sq + i, v);
    if (ema != nullptr) {
      float4 e = ld4(ema + i);
      e.x = ed * e.x + ed1 * p.x, e.y = ed * e.y + ed1 * p.y, e.z = ed * e.z + ed1 * p.z, e.w = ed * e.w + ed1 * p.w;
      st4(ema + i, e);
    }
    if (hi != nullptr) {
      st4(hi + i, p);
      if (lo != nullptr) {
        const float4 h4 = ld4(hi + i);
        st4(lo + i, make_float4(p.x - h4.x, p.y - h4.y, p.z - h4.z, p.w - h4.w));
      }
    }
  }
  for (long i = vend + threadIdx.x; i < end; i += 256) {
    float p = t.param[i], m = t.exp_avg[i], v = t.exp_avg_sq[i];
    update(p, t.grad[i], m, v);
    t.param[i] = p;
    t.exp_avg[i] = m;
    t.exp_avg_sq[i] = v;
    if (ema != nullptr) ema[i] = ed * ema[i] + ed1 * p;
    if (hi != nullptr) {
      const bf16 h = __float2bfloat16_rn(p);
      hi[i] = h;
      if (lo != nullptr) lo[i] = __float2bfloat16_rn(p - __bfloat162float(h));
    }
  }
}


}  // namespace
}  // namespace vsx

using namespace vsx;
#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int vsx_soft_ce(const float* logits, long ld, const float* target, long ldt, int rows, int cols, float loss_scale,
                           float grad_scale, float* loss_sum, float* dlogits, long ldd, void* stream) {
  VSX_REQUIRE(cols > 0 && cols <= CE_MAX_PER_LANE * 32, "vsx_soft_ce: supports up to %d classes (got %d)", CE_MAX_PER_LANE * 32, cols);
  if (rows <= 0) return VSX_OK;
  const int grid = std::min(ceil_div(rows, CE_WARPS), num_sms() * 4);
  launch_pdl(soft_ce_kernel, dim3(grid), dim3(CE_WARPS * 32), 0, ST, logits, ld, target, ldt, rows, cols, loss_scale, grad_scale, loss_sum, dlogits, ldd);
  return check_launch("vsx_soft_ce");
}

extern "C" int vsx_scale_by_scalar(float* x, long n, const float* scalar_dev, void* stream) {
  if (n <= 0) return VSX_OK;
  const int grid = (int)std::min<long>(ceil_div_l(n, 1024), (long)num_sms() * 8);
  launch_pdl(scale_by_scalar_kernel, dim3(grid), dim3(256), 0, ST, x, n, scalar_dev);
  return check_launch("vsx_scale_by_scalar");
}

extern "C" int vsx_adamw_chunk_elems(void) { return ADAM_CHUNK; }

extern "C" int vsx_adamw(const vsx_adamw_tensor* tensors_dev, const int* chunk_tensor_dev, const int* chunk_index_dev, int num_chunks,
                         double lr, double beta1, double beta2, double eps, int step, const float* grad_scale_dev, const float* guard_loss_dev,
                         int* nonfinite_count_dev, void* stream) {
  VSX_REQUIRE(step >= 1, "vsx_adamw: step counts from 1");
  if (num_chunks <= 0) return VSX_OK;
  // hyper-parameters arrive as doubles so that 1 - beta is rounded ONCE, like torch.optim.AdamW (`value=1 - beta2` is evaluated in
  // Python's double arithmetic): 1.0f - 0.999f is 1.3e-5 away from float(0.001)
  const float bc1 = (float)(1.0 - pow(beta1, (double)step)), bc2 = (float)(1.0 - pow(beta2, (double)step));
  launch_pdl(adamw_kernel, dim3(num_chunks), dim3(256), 0, ST, tensors_dev, chunk_tensor_dev, chunk_index_dev, (float)lr, (float)beta1, (float)beta2,
                                             (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, bc1, bc2, grad_scale_dev, guard_loss_dev,
                                             nonfinite_count_dev);
  return check_launch("vsx_adamw");
}
