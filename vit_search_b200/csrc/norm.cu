// Masked LayerNorm forward / backward (HBM-bound; one warp per row, 16-byte loads, shuffle reductions).
//
// Restates MaskedLayerNormFunc (nets/masked_layer_norm.py:23-50 forward, :55-88 backward) plus the `x * mask`
// at :124 for PREFIX masks: statistics use the first `keep` channels only (mean = sum/keep, var = E[x^2]-mean^2,
// eps inside the sqrt), output channels >= keep are zero.  keep == C is torch's F.layer_norm (:121).
// The reference spends ~15 ATen passes over [B,N,C] per direction; here each direction is one pass:
//   forward : read x fp32 (4 B/ch), write y bf16 (2 B/ch) + 8 B/row of statistics
//   backward: read dy (2 B) + x (4 B) + g_in (4 B), write g_out (4 B); dgamma/dbeta via per-CTA partials + atomics
#include <algorithm>

#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace vsx {
namespace {

constexpr int LN_WARPS = 8;

// Row remap for the final norm (see vsx.h): returns the output row and selects y vs y2.
__device__ __forceinline__ long remap_row(long r, int rps, int split, bool& second) {
  second = false;
  if (rps <= 0) return r;
  const long b = r / rps;
  const int t = (int)(r - b * rps);
  if (t < split) return b * split + t;
  second = true;
  return b * (rps - split) + (t - split);
}

// NV = number of float4 chunks per lane: lane owns columns (i*32 + lane)*4 .. +3 for i < NV.
template <int NV, typename T>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const float* __restrict__ x, long ldx, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, T* __restrict__ y, T* __restrict__ y2,
                                                                long ldy, float* __restrict__ mean, float* __restrict__ rstd,
                                                                int rows, int C, int keep, float eps, int rps, int split) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_keep = 1.0f / (float)keep;
  for (long r = (long)blockIdx.x * LN_WARPS + warp; r < rows; r += (long)gridDim.x * LN_WARPS) {
    const float* xr = x + r * ldx;
    float4 v[NV];
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < keep) {
        v[i] = ld4(xr + c);
        if (c + 1 >= keep) v[i].y = 0.f;
        if (c + 2 >= keep) v[i].z = 0.f;
        if (c + 3 >= keep) v[i].w = 0.f;
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    const float mu = s * inv_keep;
    const float var = ss * inv_keep - mu * mu;
    const float rs = 1.0f / sqrtf(var + eps);
    if (lane == 0) {
      mean[r] = mu;
      rstd[r] = rs;
    }
    bool second;
    const long orow = remap_row(r, rps, split, second);
    T* yr = (second ? y2 : y) + orow * ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < keep) {
          const float4 gm = ld4(gamma + c), bt = ld4(beta + c);
          o.x = gm.x * ((v[i].x - mu) * rs) + bt.x;
          o.y = (c + 1 < keep) ? gm.y * ((v[i].y - mu) * rs) + bt.y : 0.f;
          o.z = (c + 2 < keep) ? gm.z * ((v[i].z - mu) * rs) + bt.z : 0.f;
          o.w = (c + 3 < keep) ? gm.w * ((v[i].w - mu) * rs) + bt.w : 0.f;
        }
        st4(yr + c, o);
      }
    }
  }
}

template <int NV, typename T>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ dy2, long lddy,
                                                                const float* __restrict__ x, long ldx, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                const float* __restrict__ g_in, float* __restrict__ g_out, long ldg,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int C,
                                                                int keep, int rps, int split) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_keep = 1.0f / (float)keep;
  float4 gm[NV], ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    gm[i] = (c < keep) ? ld4(gamma + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long r = (long)blockIdx.x * LN_WARPS + warp; r < rows; r += (long)gridDim.x * LN_WARPS) {
    bool second;
    const long irow = remap_row(r, rps, split, second);
    const T* dyr = (second ? dy2 : dy) + irow * lddy;
    const float* xr = x + r * ldx;
    const float mu = mean[r], rs = rstd[r];
    float4 d[NV], z[NV], gi4[NV];
    float s1 = 0.f, s2 = 0.f;
    const float* gi = g_in != nullptr ? g_in + r * ldg : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {   // all global loads of the row are issued before the first reduction
      const int c = (i * 32 + lane) * 4;
      gi4[i] = (gi != nullptr && c < C) ? ld4(gi + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < keep) {
        d[i] = ld4(dyr + c);
        const float4 xv = ld4(xr + c);
        z[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        if (c + 1 >= keep) d[i].y = 0.f, z[i].y = 0.f;
        if (c + 2 >= keep) d[i].z = 0.f, z[i].z = 0.f;
        if (c + 3 >= keep) d[i].w = 0.f, z[i].w = 0.f;
        // parameter gradients: g_gamma = sum dy*z, g_beta = sum dy   (masked_layer_norm.py:78-86)
        ag[i].x += d[i].x * z[i].x, ag[i].y += d[i].y * z[i].y, ag[i].z += d[i].z * z[i].z, ag[i].w += d[i].w * z[i].w;
        ab[i].x += d[i].x, ab[i].y += d[i].y, ab[i].z += d[i].z, ab[i].w += d[i].w;
        // dz = dy * gamma
        d[i].x *= gm[i].x, d[i].y *= gm[i].y, d[i].z *= gm[i].z, d[i].w *= gm[i].w;
      }
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * z[i].x + d[i].y * z[i].y) + (d[i].z * z[i].z + d[i].w * z[i].w);
    }
    s1 = warp_sum(s1) * inv_keep;   // mean(dz)/p
    s2 = warp_sum(s2) * inv_keep;   // mean(z*dz)/p
    float* go = g_out + r * ldg;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) {
        float4 o = gi4[i];
        if (c < keep) {
          o.x += (d[i].x - s1 - z[i].x * s2) * rs;
          if (c + 1 < keep) o.y += (d[i].y - s1 - z[i].y * s2) * rs;
          if (c + 2 < keep) o.z += (d[i].z - s1 - z[i].z * s2) * rs;
          if (c + 3 < keep) o.w += (d[i].w - s1 - z[i].w * s2) * rs;
        }
        st4(go + c, o);
      }
    }
  }
  // cross-warp reduction of the per-lane column partials (one barrier per statistic), then one atomic per column per CTA
  __shared__ float4 red[LN_WARPS][NV * 32];
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? dgamma : dbeta;
    if (pass == 1) __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) red[warp][i * 32 + lane] = pass == 0 ? ag[i] : ab[i];
    __syncthreads();
    for (int q = threadIdx.x; q < NV * 32; q += LN_WARPS * 32) {
      float4 t = red[0][q];
#pragma unroll
      for (int w = 1; w < LN_WARPS; ++w) {
        const float4 u = red[w][q];
        t.x += u.x, t.y += u.y, t.z += u.z, t.w += u.w;
      }
      red_add4(dst + q * 4, t, q * 4, keep);
    }
  }
}


// ---------------------------------------------------------------- backward with bulk-copy (TMA 1-D) row prefetch
// Same arithmetic as ln_bwd_kernel, different memory pipeline: every warp owns two shared-memory row slots; lane 0 issues
// cp.async.bulk copies of the NEXT row (dy, x, g_in: up to 10 B/channel) into one slot while the warp works on the other, completion
// on one mbarrier per slot.  The per-warp register tile (d, z, g_in) of the plain kernel disappears -- rows are re-read from shared
// memory -- so occupancy is no longer register-bound at C = 512 / 1024, and ~100 KB of loads are in flight per SM.
// CAST: the kernel also emits what the NEXT half block's backward starts with -- cast.out = T(cast.scale[row / rps] * g_out) masked to the first
// cast.keep channels, and its column sums (that block's proj / fc2 bias gradient) -- so the fp32 gradient is not read back by a separate
// vsx_scale_mask_cast launch.
struct LnCast {
  void* out;
  long ld;
  const float* scale;
  int rps, keep;
  float* colsum;
};

template <int NV, typename T, bool CAST>
__global__ void __launch_bounds__(LN_WARPS * 32, NV <= 2 ? 4 : 1) ln_bwd_bulk_kernel(const T* __restrict__ dy, long lddy, const float* __restrict__ x, long ldx,
                                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                     const float* __restrict__ gamma, const float* __restrict__ g_in,
                                                                     float* __restrict__ g_out, long ldg, float* __restrict__ dgamma,
                                                                     float* __restrict__ dbeta, int rows, int C, int keep, const LnCast cast, const RowSegs segs) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // `keep` is the LARGEST kept width of the launch (slot size, width of the column reductions); with segments every row uses its own
  const uint32_t xb = (uint32_t)keep * 4, gb = g_in != nullptr ? (uint32_t)C * 4 : 0u, db = (uint32_t)keep * (uint32_t)sizeof(T);
  const uint32_t slot = ((xb + gb + db) + 127u) & ~127u;
  uint8_t* my = ln_smem + (size_t)warp * 2 * slot;
  __shared__ __align__(8) unsigned long long bars[LN_WARPS][2];
  const uint32_t bar0 = smem_u32(&bars[warp][0]), bar1 = smem_u32(&bars[warp][1]);
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    fence_mbar_init();
  }
  __syncwarp();
  const long stride = (long)gridDim.x * LN_WARPS;
  auto issue = [&](long r, int sl) {     // lane 0 only
    const uint32_t dst = smem_u32(my + (size_t)sl * slot), bar = sl ? bar1 : bar0;
    const int kr = segs.count ? segs.keep[seg_of_row(segs, r)] : keep;
    if (kr == 0) {                       // skipped row (dropped layer): complete the phase so that the slot parities stay in step
      mbar_expect_tx(bar, 0);
      return;
    }
    const uint32_t xr = (uint32_t)kr * 4, dr = (uint32_t)kr * (uint32_t)sizeof(T);
    mbar_expect_tx(bar, xr + gb + dr);
    bulk_g2s(dst, x + r * ldx, xr, bar);
    if (gb) bulk_g2s(dst + xb, g_in + r * ldg, gb, bar);
    bulk_g2s(dst + xb + gb, dy + r * lddy, dr, bar);
  };
  float4 ag[NV], ab[NV], ac[CAST ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = make_float4(0.f, 0.f, 0.f, 0.f), ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < (CAST ? NV : 1); ++i) ac[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  long r = (long)blockIdx.x * LN_WARPS + warp;
  if (r < rows && lane == 0) issue(r, 0);
  int it = 0;
  for (; r < rows; r += stride, ++it) {
    const int sl = it & 1;
    if (r + stride < rows && lane == 0) issue(r + stride, sl ^ 1);      // the other slot was fully consumed one iteration ago
    float cs = 1.0f;
    if (CAST && cast.scale != nullptr) cs = __ldg(cast.scale + r / cast.rps);
    int keep_r = keep, cast_keep = cast.keep;
    if (segs.count) {
      const int si = seg_of_row(segs, r);
      keep_r = segs.keep[si], cast_keep = segs.keep2[si];
    }
    mbar_wait(sl ? bar1 : bar0, (uint32_t)(it >> 1) & 1u);
    if (keep_r == 0) {
      __syncwarp();
      continue;
    }
    const float inv_keep = 1.0f / (float)keep_r;
    const float* xs = reinterpret_cast<const float*>(my + (size_t)sl * slot);
    const float* gs = reinterpret_cast<const float*>(my + (size_t)sl * slot + xb);
    const T* ds = reinterpret_cast<const T*>(my + (size_t)sl * slot + xb + gb);
    const float mu = mean[r], rs = rstd[r];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < keep_r) {
        float4 d = ld4(ds + c);
        const float4 xv = ld4(xs + c), gm = ld4(gamma + c);
        const float4 z = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        ag[i].x += d.x * z.x, ag[i].y += d.y * z.y, ag[i].z += d.z * z.z, ag[i].w += d.w * z.w;
        ab[i].x += d.x, ab[i].y += d.y, ab[i].z += d.z, ab[i].w += d.w;
        d.x *= gm.x, d.y *= gm.y, d.z *= gm.z, d.w *= gm.w;
        s1 += (d.x + d.y) + (d.z + d.w);
        s2 += (d.x * z.x + d.y * z.y) + (d.z * z.z + d.w * z.w);
      }
    }
    s1 = warp_sum(s1) * inv_keep;
    s2 = warp_sum(s2) * inv_keep;
    float* go = g_out + r * ldg;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) {
        float4 o = gb ? ld4(gs + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < keep_r) {
          const float4 d0 = ld4(ds + c), xv = ld4(xs + c), gm = ld4(gamma + c);
          o.x += (d0.x * gm.x - s1 - (xv.x - mu) * rs * s2) * rs;
          o.y += (d0.y * gm.y - s1 - (xv.y - mu) * rs * s2) * rs;
          o.z += (d0.z * gm.z - s1 - (xv.z - mu) * rs * s2) * rs;
          o.w += (d0.w * gm.w - s1 - (xv.w - mu) * rs * s2) * rs;
        }
        st4(go + c, o);
        if (CAST) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < cast_keep) v = make_float4(o.x * cs, o.y * cs, o.z * cs, o.w * cs);
          st4(static_cast<T*>(cast.out) + r * cast.ld + c, v);
          ac[i].x += v.x, ac[i].y += v.y, ac[i].z += v.z, ac[i].w += v.w;
        }
      }
    }
    __syncwarp();      // every lane is done with this slot before lane 0 re-arms it (two iterations from now it is the target again)
  }
  __shared__ float4 red[LN_WARPS][NV * 32];
  const int npass = (CAST && cast.colsum != nullptr) ? 3 : 2;
  for (int pass = 0; pass < npass; ++pass) {
    float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : cast.colsum);
    if (pass >= 1) __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) red[warp][i * 32 + lane] = pass == 0 ? ag[i] : (pass == 1 ? ab[i] : ac[CAST ? i : 0]);
    __syncthreads();
    for (int q = threadIdx.x; q < NV * 32; q += LN_WARPS * 32) {
      float4 t = red[0][q];
#pragma unroll
      for (int w = 1; w < LN_WARPS; ++w) {
        const float4 u = red[w][q];
        t.x += u.x, t.y += u.y, t.z += u.z, t.w += u.w;
      }
      red_add4(dst + q * 4, t, q * 4, pass == 2 ? cast.keep : keep);
    }
  }
}


// Forward with the same bulk-copy row prefetch (see ln_bwd_bulk_kernel): x rows arrive in shared memory one iteration ahead.
template <int NV, typename T>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_bulk_kernel(const float* __restrict__ x, long ldx, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, T* __restrict__ y, long ldy,
                                                                     float* __restrict__ mean, float* __restrict__ rstd, int rows, int C,
                                                                     int keep, float eps, const RowSegs segs) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t xb = (uint32_t)keep * 4;       // largest kept width of the launch
  const uint32_t slot = (xb + 127u) & ~127u;
  uint8_t* my = ln_smem + (size_t)warp * 2 * slot;
  __shared__ __align__(8) unsigned long long bars[LN_WARPS][2];
  const uint32_t bar0 = smem_u32(&bars[warp][0]), bar1 = smem_u32(&bars[warp][1]);
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    fence_mbar_init();
  }
  __syncwarp();
  const long stride = (long)gridDim.x * LN_WARPS;
  auto issue = [&](long r, int sl) {
    const uint32_t bar = sl ? bar1 : bar0;
    const uint32_t xr = (uint32_t)(segs.count ? segs.keep[seg_of_row(segs, r)] : keep) * 4;
    mbar_expect_tx(bar, xr);             // 0 bytes for a skipped row: the phase completes at once, slot parities stay in step
    if (xr) bulk_g2s(smem_u32(my + (size_t)sl * slot), x + r * ldx, xr, bar);
  };
  long r = (long)blockIdx.x * LN_WARPS + warp;
  if (r < rows && lane == 0) issue(r, 0);
  int it = 0;
  for (; r < rows; r += stride, ++it) {
    const int sl = it & 1;
    if (r + stride < rows && lane == 0) issue(r + stride, sl ^ 1);
    const int keep_r = segs.count ? segs.keep[seg_of_row(segs, r)] : keep;
    mbar_wait(sl ? bar1 : bar0, (uint32_t)(it >> 1) & 1u);
    if (keep_r == 0) {
      __syncwarp();
      continue;
    }
    const float inv_keep = 1.0f / (float)keep_r;
    const float* xs = reinterpret_cast<const float*>(my + (size_t)sl * slot);
    float4 v[NV];
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      v[i] = c < keep_r ? ld4(xs + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    const float mu = s * inv_keep;
    const float var = ss * inv_keep - mu * mu;
    const float rs = 1.0f / sqrtf(var + eps);
    if (lane == 0) {
      mean[r] = mu;
      rstd[r] = rs;
    }
    T* yr = y + r * ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < keep_r) {
          const float4 gm = ld4(gamma + c), bt = ld4(beta + c);
          o.x = gm.x * ((v[i].x - mu) * rs) + bt.x;
          o.y = gm.y * ((v[i].y - mu) * rs) + bt.y;
          o.z = gm.z * ((v[i].z - mu) * rs) + bt.z;
          o.w = gm.w * ((v[i].w - mu) * rs) + bt.w;
        }
        st4(yr + c, o);
      }
    }
    __syncwarp();
  }
}

int ln_grid(int rows, int per_sm) {
  const int need = ceil_div(rows, LN_WARPS);
  const int cap = num_sms() * per_sm;
  return need < cap ? need : cap;
}

template <typename T>
int ln_fwd_dispatch(const float* x, long ldx, const float* gamma, const float* beta, void* y, void* y2, long ldy, float* mean,
                    float* rstd, int rows, int C, int keep, float eps, int rps, int split, cudaStream_t st, const RowSegs* segs = nullptr) {
  const int nv = ceil_div(C, 128);
  // segments (several extents in one launch) exist in the bulk variant only: returns 1 ("not handled") when it does not apply
  bool seg_ok = true;
  if (segs != nullptr) {
    keep = 0;
    for (int i = 0; i < segs->count; ++i) seg_ok = seg_ok && segs->keep[i] % 4 == 0, keep = segs->keep[i] > keep ? segs->keep[i] : keep;
    if (keep == 0) return VSX_OK;
  }
  // bulk-copy prefetch variant: kept prefix in whole float4s (the masked tail of a row is then exactly c >= keep), no row remap
  if (seg_ok && rps <= 0 && keep % 4 == 0 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const size_t slot = ((size_t)keep * 4 + 127) & ~(size_t)127;
    const size_t smem = slot * 2 * LN_WARPS;
    // CTAs per SM: 4 -> 6 (what 40 registers per thread allow to be resident) measured 32.9 -> 24.7 us at stage 1, 14.3 -> 13.9 us at stage 2
    // (tools/ln_bench.py), 13.30 -> 13.19 ms per train step; 8 requested CTAs are no faster and add a second wave at the small stages
    static const int fwd_per_sm = getenv("VSX_LN_FWD_PER_SM") ? atoi(getenv("VSX_LN_FWD_PER_SM")) : 6;
    const int gridb = ln_grid(rows, fwd_per_sm);
#define VSX_LN_FB(NV)                                                                                                            \
  case NV: {                                                                                                                     \
    static bool cfg = false;                                                                                                     \
    if (!cfg) {                                                                                                                  \
      cudaFuncSetAttribute(ln_fwd_bulk_kernel<NV, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);                   \
      cfg = true;                                                                                                                \
    }                                                                                                                            \
    launch_pdl(ln_fwd_bulk_kernel<NV, T>, dim3(gridb), dim3(LN_WARPS * 32), smem, st, x, ldx, gamma, beta, (T*)y, ldy, mean, rstd, rows, C, keep, eps, segs ? *segs : RowSegs{}); \
    return check_launch("vsx_masked_ln_fwd");                                                                                    \
  }
    switch (nv) {
      VSX_LN_FB(1) VSX_LN_FB(2) VSX_LN_FB(3) VSX_LN_FB(4) VSX_LN_FB(5) VSX_LN_FB(6) VSX_LN_FB(7) VSX_LN_FB(8) VSX_LN_FB(9) VSX_LN_FB(10)
      default: break;
    }
#undef VSX_LN_FB
  }
  if (segs != nullptr) return 1;
  const int grid = ln_grid(rows, 8);
#define VSX_LN_F(NV)                                                                                                        \
  case NV:                                                                                                                  \
    launch_pdl(ln_fwd_kernel<NV, T>, dim3(grid), dim3(LN_WARPS * 32), 0, st, x, ldx, gamma, beta, (T*)y, (T*)y2, ldy, mean, rstd, rows, C, keep, \
                                                          eps, rps, split);                                                  \
    break;
  switch (nv) {
    VSX_LN_F(1) VSX_LN_F(2) VSX_LN_F(3) VSX_LN_F(4) VSX_LN_F(5) VSX_LN_F(6) VSX_LN_F(7) VSX_LN_F(8) VSX_LN_F(9) VSX_LN_F(10)
    default:
      set_error("vsx_masked_ln_fwd: C=%d exceeds the supported 1280 channels", C);
      return VSX_ERR_ARG;
  }
#undef VSX_LN_F
  return check_launch("vsx_masked_ln_fwd");
}

template <typename T>
int ln_bwd_dispatch(const void* dy, const void* dy2, long lddy, const float* x, long ldx, const float* mean, const float* rstd,
                    const float* gamma, const float* g_in, float* g_out, long ldg, float* dgamma, float* dbeta, int rows, int C,
                    int keep, int rps, int split, cudaStream_t st, const LnCast* cast = nullptr, bool* cast_done = nullptr,
                    const RowSegs* segs = nullptr) {
  const int nv = ceil_div(C, 128);
  bool seg_ok = true;
  if (segs != nullptr) {        // several extents in one launch: bulk variant only; returns 1 ("not handled") when it does not apply
    keep = 0;
    for (int i = 0; i < segs->count; ++i) seg_ok = seg_ok && segs->keep[i] % 8 == 0, keep = segs->keep[i] > keep ? segs->keep[i] : keep;
    if (keep == 0) return VSX_OK;
  }
  // bulk-copy prefetch variant: whole kept prefix in 16-byte units, no final-norm row remap, 16-byte aligned rows
  const size_t esz = sizeof(T);
  const bool bulk = seg_ok && rps <= 0 && keep % 8 == 0 && C % 4 == 0 && lddy % 8 == 0 && ldx % 4 == 0 && ldg % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g_in)) & 15) == 0;
  if (bulk && (((size_t)keep * 4 + (g_in != nullptr ? (size_t)C * 4 : 0) + (size_t)keep * esz + 127) & ~(size_t)127) * 2 * LN_WARPS + 4096 * (size_t)nv + 512 <= 227 * 1024) {
    const size_t slot = (((size_t)keep * 4 + (g_in != nullptr ? (size_t)C * 4 : 0) + (size_t)keep * esz) + 127) & ~(size_t)127;
    const size_t smem = slot * 2 * LN_WARPS, stat = 4096 * (size_t)nv + 512;     // dynamic row slots + the static reduction buffer
    static const int bwd_per_sm = getenv("VSX_LN_BWD_PER_SM") ? atoi(getenv("VSX_LN_BWD_PER_SM")) : 4;
    const int per_sm = (int)std::min<size_t>(bwd_per_sm, (227 * 1024) / (smem + stat + 1024));
    const int gridb = ln_grid(rows, per_sm < 1 ? 1 : per_sm);
#define VSX_LN_BB(NV)                                                                                                               \
  case NV: {                                                                                                                        \
    static bool cfg = false;                                                                                                        \
    if (!cfg) {                                                                                                                     \
      cudaFuncSetAttribute(ln_bwd_bulk_kernel<NV, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096 * NV - 512);   \
      cfg = true;                                                                                                                   \
    }                                                                                                                               \
    if (cast != nullptr && cast->ld % 4 == 0) {                                                                                      \
      static bool cfg2 = false;                                                                                                     \
      if (!cfg2) {                                                                                                                  \
        cudaFuncSetAttribute(ln_bwd_bulk_kernel<NV, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096 * NV - 512); \
        cfg2 = true;                                                                                                                \
      }                                                                                                                             \
      launch_pdl(ln_bwd_bulk_kernel<NV, T, true>, dim3(gridb), dim3(LN_WARPS * 32), smem, st, (const T*)dy, lddy, x, ldx, mean, rstd, gamma, g_in, g_out, ldg, \
                                                                           dgamma, dbeta, rows, C, keep, *cast, segs ? *segs : RowSegs{});                    \
      if (cast_done != nullptr) *cast_done = true;                                                                                  \
      return check_launch("vsx_masked_ln_bwd");                                                                                     \
    }                                                                                                                               \
    launch_pdl(ln_bwd_bulk_kernel<NV, T, false>, dim3(gridb), dim3(LN_WARPS * 32), smem, st, (const T*)dy, lddy, x, ldx, mean, rstd, gamma, g_in, g_out, ldg, \
                                                                          dgamma, dbeta, rows, C, keep, LnCast{}, segs ? *segs : RowSegs{});                  \
    return check_launch("vsx_masked_ln_bwd");                                                                                       \
  }
    switch (nv) {
      VSX_LN_BB(1) VSX_LN_BB(2) VSX_LN_BB(3) VSX_LN_BB(4) VSX_LN_BB(5) VSX_LN_BB(6) VSX_LN_BB(7) VSX_LN_BB(8) VSX_LN_BB(9) VSX_LN_BB(10)
      default: break;
    }
#undef VSX_LN_BB
  }
  if (segs != nullptr) return 1;
  const int grid = ln_grid(rows, 4);
#define VSX_LN_B(NV)                                                                                                      \
  case NV:                                                                                                                \
    launch_pdl(ln_bwd_kernel<NV, T>, dim3(grid), dim3(LN_WARPS * 32), 0, st, (const T*)dy, (const T*)dy2, lddy, x, ldx, mean, rstd, gamma, g_in, \
                                                          g_out, ldg, dgamma, dbeta, rows, C, keep, rps, split);          \
    break;
  switch (nv) {
    VSX_LN_B(1) VSX_LN_B(2) VSX_LN_B(3) VSX_LN_B(4) VSX_LN_B(5) VSX_LN_B(6) VSX_LN_B(7) VSX_LN_B(8) VSX_LN_B(9) VSX_LN_B(10)
    default:
      set_error("vsx_masked_ln_bwd: C=%d exceeds the supported 1280 channels", C);
      return VSX_ERR_ARG;
  }
#undef VSX_LN_B
  return check_launch("vsx_masked_ln_bwd");
}

}  // namespace
}  // namespace vsx

using namespace vsx;

extern "C" int vsx_masked_ln_fwd(const float* x, long ldx, const float* gamma, const float* beta, void* y, void* y2, int dtype,
                                 long ldy, float* mean, float* rstd, int rows, int C, int keep, float eps, int rows_per_sample,
                                 int split_tokens, void* stream) {
  VSX_REQUIRE(rows >= 0 && C > 0 && keep > 0 && keep <= C, "vsx_masked_ln_fwd: need 0 < keep <= C (keep=%d C=%d)", keep, C);
  VSX_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "vsx_masked_ln_fwd: C and pitches must be multiples of 4");
  VSX_REQUIRE(rows_per_sample <= 0 || (y2 != nullptr && split_tokens > 0 && split_tokens < rows_per_sample),
              "vsx_masked_ln_fwd: row remap needs y2 and 0 < split_tokens < rows_per_sample");
  if (rows == 0) return VSX_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == VSX_BF16)
    return ln_fwd_dispatch<bf16>(x, ldx, gamma, beta, y, y2, ldy, mean, rstd, rows, C, keep, eps, rows_per_sample, split_tokens, st);
  if (dtype == VSX_F32)
    return ln_fwd_dispatch<float>(x, ldx, gamma, beta, y, y2, ldy, mean, rstd, rows, C, keep, eps, rows_per_sample, split_tokens, st);
  set_error("vsx_masked_ln_fwd: bad dtype %d", dtype);
  return VSX_ERR_ARG;
}

extern "C" int vsx_masked_ln_bwd(const void* dy, const void* dy2, int dtype, long lddy, const float* x, long ldx, const float* mean,
                                 const float* rstd, const float* gamma, const float* g_in, float* g_out, long ldg, float* dgamma,
                                 float* dbeta, int rows, int C, int keep, int rows_per_sample, int split_tokens, void* stream) {
  VSX_REQUIRE(rows >= 0 && C > 0 && keep > 0 && keep <= C, "vsx_masked_ln_bwd: need 0 < keep <= C (keep=%d C=%d)", keep, C);
  VSX_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && lddy % 4 == 0 && ldg % 4 == 0, "vsx_masked_ln_bwd: C and pitches must be multiples of 4");
  VSX_REQUIRE(rows_per_sample <= 0 || (dy2 != nullptr && split_tokens > 0 && split_tokens < rows_per_sample),
              "vsx_masked_ln_bwd: row remap needs dy2 and 0 < split_tokens < rows_per_sample");
  if (rows == 0) return VSX_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == VSX_BF16)
    return ln_bwd_dispatch<bf16>(dy, dy2, lddy, x, ldx, mean, rstd, gamma, g_in, g_out, ldg, dgamma, dbeta, rows, C, keep,
                                 rows_per_sample, split_tokens, st);
  if (dtype == VSX_F32)
    return ln_bwd_dispatch<float>(dy, dy2, lddy, x, ldx, mean, rstd, gamma, g_in, g_out, ldg, dgamma, dbeta, rows, C, keep,
                                  rows_per_sample, split_tokens, st);
  set_error("vsx_masked_ln_bwd: bad dtype %d", dtype);
  return VSX_ERR_ARG;
}

// LayerNorm backward that also emits the scaled / masked low-precision copy of g_out (and its column sums) the next half block's
// backward starts from.  Falls back to vsx_masked_ln_bwd + vsx_scale_mask_cast when the fused kernel does not apply.
extern "C" int vsx_masked_ln_bwd_cast(const void* dy, int dtype, long lddy, const float* x, long ldx, const float* mean, const float* rstd,
                                      const float* gamma, const float* g_in, float* g_out, long ldg, float* dgamma, float* dbeta, int rows, int C,
                                      int keep, void* cast_out, long ld_cast, const float* cast_scale, int cast_rows_per_sample, int cast_keep,
                                      float* cast_colsum, void* stream) {
  VSX_REQUIRE(rows >= 0 && C > 0 && keep > 0 && keep <= C, "vsx_masked_ln_bwd_cast: need 0 < keep <= C (keep=%d C=%d)", keep, C);
  VSX_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && lddy % 4 == 0 && ldg % 4 == 0 && ld_cast % 4 == 0, "vsx_masked_ln_bwd_cast: C and pitches must be multiples of 4");
  VSX_REQUIRE(cast_out != nullptr && cast_keep >= 0 && cast_keep <= C, "vsx_masked_ln_bwd_cast: bad cast output / keep");
  VSX_REQUIRE(dtype == VSX_BF16 || dtype == VSX_F32, "vsx_masked_ln_bwd_cast: bad dtype %d", dtype);
  if (rows == 0) return VSX_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rps = cast_rows_per_sample > 0 ? cast_rows_per_sample : 1;
  const LnCast cast{cast_out, ld_cast, cast_scale, rps, cast_keep, cast_colsum};
  bool done = false;
  int rc;
  if (dtype == VSX_BF16)
    rc = ln_bwd_dispatch<bf16>(dy, nullptr, lddy, x, ldx, mean, rstd, gamma, g_in, g_out, ldg, dgamma, dbeta, rows, C, keep, 0, 0, st, &cast, &done);
  else
    rc = ln_bwd_dispatch<float>(dy, nullptr, lddy, x, ldx, mean, rstd, gamma, g_in, g_out, ldg, dgamma, dbeta, rows, C, keep, 0, 0, st, &cast, &done);
  if (rc != VSX_OK || done) return rc;
  return vsx_scale_mask_cast(g_out, ldg, cast_scale, rps, cast_keep, cast_out, dtype, ld_cast, rows, C, cast_colsum, stream);
}

// ---------------------------------------------------------------- several extents in one launch (multi-architecture batches)
namespace {
int check_segs(const vsx_row_segments* sg, int rows, int C, const char* what) {
  VSX_REQUIRE(sg != nullptr && sg->count >= 1 && sg->count <= VSX_MAX_SEGMENTS, "%s: 1..%d segments", what, VSX_MAX_SEGMENTS);
  int prev = 0;
  for (int i = 0; i < sg->count; ++i) {
    VSX_REQUIRE(sg->row_end[i] >= prev && sg->keep[i] >= 0 && sg->keep[i] <= C && sg->keep2[i] >= 0 && sg->keep2[i] <= C,
                "%s: segment %d is not ordered or its extents exceed C=%d (row_end=%d keep=%d keep2=%d)", what, i, C, sg->row_end[i], sg->keep[i], sg->keep2[i]);
    prev = sg->row_end[i];
  }
  VSX_REQUIRE(prev == rows, "%s: the segments must cover the %d rows of the launch (last row_end = %d)", what, rows, prev);
  return VSX_OK;
}
inline RowSegs to_segs(const vsx_row_segments* sg) {
  RowSegs r;
  static_assert(sizeof(RowSegs) == sizeof(vsx_row_segments), "RowSegs mirrors vsx_row_segments");
  memcpy(&r, sg, sizeof(r));
  return r;
}
}  // namespace

extern "C" int vsx_masked_ln_fwd_segs(const float* x, long ldx, const float* gamma, const float* beta, void* y, int dtype, long ldy, float* mean,
                                      float* rstd, int rows, int C, const vsx_row_segments* segs, float eps, void* stream) {
  int rc = check_segs(segs, rows, C, "vsx_masked_ln_fwd_segs");
  if (rc) return rc;
  VSX_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "vsx_masked_ln_fwd_segs: C and pitches must be multiples of 4");
  VSX_REQUIRE(dtype == VSX_BF16 || dtype == VSX_F32, "vsx_masked_ln_fwd_segs: bad dtype %d", dtype);
  if (rows == 0) return VSX_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const RowSegs sg = to_segs(segs);
  rc = dtype == VSX_BF16 ? ln_fwd_dispatch<bf16>(x, ldx, gamma, beta, y, nullptr, ldy, mean, rstd, rows, C, 0, eps, 0, 0, st, &sg)
                         : ln_fwd_dispatch<float>(x, ldx, gamma, beta, y, nullptr, ldy, mean, rstd, rows, C, 0, eps, 0, 0, st, &sg);
  if (rc != 1) return rc;
  const size_t es = dtype == VSX_BF16 ? 2 : 4;      // not handled in one launch: one launch per segment
  for (int i = 0, r0 = 0; i < segs->count; r0 = segs->row_end[i], ++i) {
    if (segs->keep[i] == 0 || segs->row_end[i] == r0) continue;
    rc = vsx_masked_ln_fwd(x + (long)r0 * ldx, ldx, gamma, beta, static_cast<uint8_t*>(y) + (size_t)r0 * ldy * es, nullptr, dtype, ldy, mean + r0, rstd + r0,
                           segs->row_end[i] - r0, C, segs->keep[i], eps, 0, 0, stream);
    if (rc) return rc;
  }
  return VSX_OK;
}

// keep2 of a segment = the cast extent when cast_out != NULL (see vsx_masked_ln_bwd_cast)
extern "C" int vsx_masked_ln_bwd_segs(const void* dy, int dtype, long lddy, const float* x, long ldx, const float* mean, const float* rstd,
                                      const float* gamma, const float* g_in, float* g_out, long ldg, float* dgamma, float* dbeta, int rows, int C,
                                      const vsx_row_segments* segs, void* cast_out, long ld_cast, const float* cast_scale, int cast_rows_per_sample,
                                      float* cast_colsum, void* stream) {
  int rc = check_segs(segs, rows, C, "vsx_masked_ln_bwd_segs");
  if (rc) return rc;
  VSX_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && lddy % 4 == 0 && ldg % 4 == 0 && ld_cast % 4 == 0, "vsx_masked_ln_bwd_segs: C and pitches must be multiples of 4");
  VSX_REQUIRE(dtype == VSX_BF16 || dtype == VSX_F32, "vsx_masked_ln_bwd_segs: bad dtype %d", dtype);
  if (rows == 0) return VSX_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const RowSegs sg = to_segs(segs);
  const int rps = cast_rows_per_sample > 0 ? cast_rows_per_sample : 1;
  int keep2_max = 0;        // width of the cast's column sums: the largest cast extent of the launch
  for (int i = 0; i < segs->count; ++i) keep2_max = segs->keep2[i] > keep2_max ? segs->keep2[i] : keep2_max;
  const LnCast cast{cast_out, ld_cast, cast_scale, rps, keep2_max, cast_colsum};
  bool done = false;
  rc = dtype == VSX_BF16 ? ln_bwd_dispatch<bf16>(dy, nullptr, lddy, x, ldx, mean, rstd, gamma, g_in, g_out, ldg, dgamma, dbeta, rows, C, 0, 0, 0, st,
                                                 cast_out != nullptr ? &cast : nullptr, &done, &sg)
                         : ln_bwd_dispatch<float>(dy, nullptr, lddy, x, ldx, mean, rstd, gamma, g_in, g_out, ldg, dgamma, dbeta, rows, C, 0, 0, 0, st,
                                                  cast_out != nullptr ? &cast : nullptr, &done, &sg);
  if (rc < 0) return rc;
  const size_t es = dtype == VSX_BF16 ? 2 : 4;
  if (rc == 1) {                                      // not handled in one launch: one launch per segment
    for (int i = 0, r0 = 0; i < segs->count; r0 = segs->row_end[i], ++i) {
      if (segs->keep[i] == 0 || segs->row_end[i] == r0) continue;
      rc = vsx_masked_ln_bwd(static_cast<const uint8_t*>(dy) + (size_t)r0 * lddy * es, nullptr, dtype, lddy, x + (long)r0 * ldx, ldx, mean + r0, rstd + r0, gamma,
                             g_in != nullptr ? g_in + (long)r0 * ldg : nullptr, g_out + (long)r0 * ldg, ldg, dgamma, dbeta, segs->row_end[i] - r0, C,
                             segs->keep[i], 0, 0, stream);
      if (rc) return rc;
    }
  }
  if (cast_out != nullptr && !done) {                 // the cast did not ride on the LayerNorm backward: separate pass per segment
    for (int i = 0, r0 = 0; i < segs->count; r0 = segs->row_end[i], ++i) {
      if (segs->keep[i] == 0 || segs->keep2[i] == 0 || segs->row_end[i] == r0) continue;      // keep2 == 0: the consumer drops the layer there
      rc = vsx_scale_mask_cast(g_out + (long)r0 * ldg, ldg, cast_scale != nullptr ? cast_scale + r0 / rps : nullptr, rps, segs->keep2[i],
                               static_cast<uint8_t*>(cast_out) + (size_t)r0 * ld_cast * es, dtype, ld_cast, segs->row_end[i] - r0, C, cast_colsum, stream);
      if (rc) return rc;
    }
  }
  return VSX_OK;
}
