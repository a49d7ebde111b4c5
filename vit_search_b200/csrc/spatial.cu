// Spatial plumbing around the tensor-core GEMM for the conv layers of the path (HBM-bound, no reuse):
//   im2col / col2im      -- conv3x3 (stride 1|2, pad 1) and the patch projections (k = stride, pad 0) as GEMMs:
//                           PatchConvEmbed (nets/patch_conv.py:23-36, :63-74) and the SR block's patch_reduce
//                           (nets/vit_sr_supernet.py:139-143).  BatchNorm-apply + ReLU of the producing layer and the
//                           stem's residual add (:69-71) are fused into the im2col gather.
//   BatchNorm2d (train)  -- batch statistics, running-stat update, backward reductions and apply (patch_conv.py:28).
//   embed / SR assembly  -- class token + position embedding + embedding mask (vit_sr_supernet.py:399-407) and the SR
//                           block's pool / pad / pos / token / mask combine (:131-166), forward and backward.
// Feature maps are channels-last [B, H, W, C]; token tensors are [B, 1 + g*g, C].
#include "common.cuh"

namespace vsx {
namespace {

template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return Store<T>::ld(p); }

// ------------------------------------------------------------------------------------------------ im2col
// out[(b,oy,ox), (ky*k+kx)*C + c] = act1(in1[b, oy*s-p+ky, ox*s-p+kx, c]) (+ act2(in2[...])), zero outside the image.
// act(v) = relu(v*scale[c] + shift[c]) when scale != nullptr, identity otherwise.
// NCHW=true: in1 is fp32 [B, C, H, W] (the image); otherwise channels-last with `pix_pitch` elements between pixels,
// `batch_pitch` between samples (so token tensors [B, 1+g*g, C] can be read in place, skipping the class token).
template <typename TI, typename TO, bool NCHW>
__global__ void im2col_kernel(const TI* __restrict__ in1, const float* __restrict__ sc1, const float* __restrict__ sh1,
                              const TI* __restrict__ in2, const float* __restrict__ sc2, const float* __restrict__ sh2,
                              long batch_pitch, long pix_pitch, int B, int H, int W, int C, int k, int s, int p, int Ho, int Wo,
                              TO* __restrict__ out, long ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)B * Ho * Wo * k * k;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % (k * k));
    const long pix = i / (k * k);
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long)Wo * Ho));
    const int ky = tap / k, kx = tap % k;
    const int iy = oy * s - p + ky, ix = ox * s - p + kx;
    TO* o = out + pix * ldo + (long)tap * C;
    const bool inside = iy >= 0 && iy < H && ix >= 0 && ix < W;
    if (!inside) {
      for (int c = 0; c < C; ++c) Store<TO>::st(o + c, 0.f);
      continue;
    }
    if (NCHW) {
      const TI* src = in1 + (long)b * batch_pitch + (long)iy * W + ix;
      for (int c = 0; c < C; ++c) Store<TO>::st(o + c, ldf(src + (long)c * H * W));
    } else {
      const long off = (long)b * batch_pitch + ((long)iy * W + ix) * pix_pitch;
      const TI* src = in1 + off;
      for (int c = 0; c < C; c += 4) {
        float4 v = ld4(src + c);
        if (sc1 != nullptr) {
          const float4 a = ld4(sc1 + c), d = ld4(sh1 + c);
          v.x = fmaxf(v.x * a.x + d.x, 0.f), v.y = fmaxf(v.y * a.y + d.y, 0.f);
          v.z = fmaxf(v.z * a.z + d.z, 0.f), v.w = fmaxf(v.w * a.w + d.w, 0.f);
        }
        if (in2 != nullptr) {
          float4 w = ld4(in2 + off + c);
          if (sc2 != nullptr) {
            const float4 a = ld4(sc2 + c), d = ld4(sh2 + c);
            w.x = fmaxf(w.x * a.x + d.x, 0.f), w.y = fmaxf(w.y * a.y + d.y, 0.f);
            w.z = fmaxf(w.z * a.z + d.z, 0.f), w.w = fmaxf(w.w * a.w + d.w, 0.f);
          }
          v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;
        }
        st4(o + c, v);
      }
    }
  }
}

// conv1 of the stem (patch_conv.py:25-30): fp32 NCHW image, 3 channels, 3x3, stride 2, pad 1 -> bf16 rows of 27 (+5 zero) columns.
// One thread per output pixel: 27 image loads (neighbouring threads share them through L1) and ONE 64-byte row written with four
// 16-byte stores, instead of one thread per (pixel, tap) writing three 2-byte values.
__global__ void __launch_bounds__(256) im2col_conv1_kernel(const float* __restrict__ img, int B, int H, int W, int Ho, int Wo,
                                                           bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)B * Ho * Wo;
  for (long pix = (long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long)Wo * Ho));
    const float* src = img + (long)b * 3 * H * W;
    float v[32];
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy - 1 + ky;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = 2 * ox - 1 + kx;
        const bool in = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[(ky * 3 + kx) * 3 + c] = in ? __ldg(src + ((long)c * H + iy) * W + ix) : 0.f;
      }
    }
    uint4* o = reinterpret_cast<uint4*>(out + pix * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]), pack_bf16(v[8 * i + 4], v[8 * i + 5]),
                        pack_bf16(v[8 * i + 6], v[8 * i + 7]));
  }
}

// ------------------------------------------------------------------------------------------------ col2im
// din[b, iy, ix, c] = (add ? add[...] : 0) + sum over taps (ky,kx) with oy = (iy+p-ky)/s, ox = (ix+p-kx)/s integral and in
// range of dcol[(b,oy,ox), (ky*k+kx)*C + c].  One thread per (input pixel, 4 channels).
template <typename T>
__global__ void col2im_kernel(const T* __restrict__ dcol, long ldc, const T* __restrict__ add, int B, int H, int W, int C, int k, int s,
                              int p, int Ho, int Wo, T* __restrict__ din, long batch_pitch, long pix_pitch) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = C / 4;
  const long total = (long)B * H * W * c4n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long pix = i / c4n;
    const int ix = (int)(pix % W), iy = (int)((pix / W) % H), b = (int)(pix / ((long)W * H));
    const long off = (long)b * batch_pitch + ((long)iy * W + ix) * pix_pitch + c;
    float4 acc = add != nullptr ? ld4(add + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (k == s && p == 0) {   // non-overlapping patches (conv_proj): exactly one tap per pixel, a pure permutation
      const int oy = iy / k, ox = ix / k;
      if (oy < Ho && ox < Wo) {
        const float4 v = ld4(dcol + (((long)b * Ho + oy) * Wo + ox) * ldc + (long)((iy - oy * k) * k + (ix - ox * k)) * C + c);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
      st4(din + off, acc);
      continue;
    }
    for (int ky = 0; ky < k; ++ky) {
      const int ty = iy + p - ky;
      if (ty < 0 || ty % s != 0) continue;
      const int oy = ty / s;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int tx = ix + p - kx;
        if (tx < 0 || tx % s != 0) continue;
        const int ox = tx / s;
        if (ox >= Wo) continue;
        const float4 v = ld4(dcol + (((long)b * Ho + oy) * Wo + ox) * ldc + (long)(ky * k + kx) * C + c);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
    }
    st4(din + off, acc);
  }
}

// ------------------------------------------------------------------------------------------------ bf16 row kernels
// Same semantics as im2col_kernel / col2im_kernel for bf16 channels-last maps with C % 8 == 0, re-mapped for bandwidth: one CTA per
// output (im2col) / input (col2im) image row, 16-byte chunks, consecutive lanes on consecutive chunks of a contiguous run (the k taps
// of one kernel row are contiguous both in the source row and in the column matrix), 32-bit index math hoisted out of the chunk loop.
// The generic kernels above used one thread per (pixel, tap) with 8-byte accesses 2*C bytes apart and 64-bit divisions per element.
struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 unpack8(uint4 r) {
  F8 o;
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    o.v[2 * i] = f.x, o.v[2 * i + 1] = f.y;
  }
  return o;
}
__device__ __forceinline__ uint4 pack8(const F8& f) {
  return make_uint4(pack_bf16(f.v[0], f.v[1]), pack_bf16(f.v[2], f.v[3]), pack_bf16(f.v[4], f.v[5]), pack_bf16(f.v[6], f.v[7]));
}
__device__ __forceinline__ void bn_relu8(F8& f, const float* __restrict__ sc, const float* __restrict__ sh, int c) {
  const float4 a0 = ld4(sc + c), a1 = ld4(sc + c + 4), d0 = ld4(sh + c), d1 = ld4(sh + c + 4);
  const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) f.v[i] = fmaxf(f.v[i] * a[i] + d[i], 0.f);
}

__global__ void __launch_bounds__(256) im2col_rows_kernel(const bf16* __restrict__ in1, const float* __restrict__ sc1, const float* __restrict__ sh1,
                                                          const bf16* __restrict__ in2, const float* __restrict__ sc2,
                                                          const float* __restrict__ sh2, long batch_pitch, int pix_pitch, int H, int W, int C, int k,
                                                          int s, int p, int Ho, int Wo, bf16* __restrict__ out, long ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x / Ho, oy = blockIdx.x - b * Ho;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int cpc = C >> 3, run = k * cpc;
  const long in_b = (long)b * batch_pitch;
  bf16* orow = out + ((long)b * Ho + oy) * Wo * ldo;
  for (int u = warp; u < Wo * k; u += nwarps) {      // unit: (output pixel, kernel row)
    const int ox = u / k, ky = u - ox * k;
    const int iy = oy * s - p + ky, ix0 = ox * s - p;
    const bool row_in = iy >= 0 && iy < H;
    bf16* o = orow + (long)ox * ldo + (long)ky * k * C;
    const long src_row = in_b + ((long)iy * W + ix0) * pix_pitch;
    for (int j = lane; j < run; j += 32) {
      const int kx = j / cpc, c = (j - kx * cpc) << 3;
      const int ix = ix0 + kx;
      uint4 r = make_uint4(0u, 0u, 0u, 0u);
      if (row_in && ix >= 0 && ix < W) {
        const long off = src_row + (long)kx * pix_pitch + c;
        r = *reinterpret_cast<const uint4*>(in1 + off);
        if (sc1 != nullptr || in2 != nullptr) {
          F8 f = unpack8(r);
          if (sc1 != nullptr) bn_relu8(f, sc1, sh1, c);
          if (in2 != nullptr) {
            F8 g = unpack8(*reinterpret_cast<const uint4*>(in2 + off));
            if (sc2 != nullptr) bn_relu8(g, sc2, sh2, c);
#pragma unroll
            for (int i = 0; i < 8; ++i) f.v[i] += g.v[i];
          }
          r = pack8(f);
        }
      }
      *reinterpret_cast<uint4*>(o + (j << 3)) = r;
    }
  }
}

__global__ void __launch_bounds__(256) col2im_rows_kernel(const bf16* __restrict__ dcol, long ldc, const bf16* __restrict__ add, int H, int W, int C,
                                                          int k, int s, int p, int Ho, int Wo, bf16* __restrict__ din, long batch_pitch,
                                                          int pix_pitch) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x / H, iy = blockIdx.x - b * H;
  const int cpc = C >> 3;
  const long row_off = (long)b * batch_pitch + (long)iy * W * pix_pitch;
  const bf16* dc_b = dcol + (long)b * Ho * Wo * ldc;
  const bool patches = (k == s && p == 0);           // non-overlapping patches (conv_proj): exactly one tap per pixel, a pure permutation
  const int oy_p = iy / k, ky_p = iy - oy_p * k;
  for (int j = threadIdx.x; j < W * cpc; j += blockDim.x) {
    const int ix = j / cpc, c = (j - ix * cpc) << 3;
    const long off = row_off + (long)ix * pix_pitch + c;
    if (patches) {
      const int ox = ix / k, kx = ix - ox * k;
      uint4 r = make_uint4(0u, 0u, 0u, 0u);
      if (oy_p < Ho && ox < Wo) r = *reinterpret_cast<const uint4*>(dc_b + ((long)oy_p * Wo + ox) * ldc + (long)(ky_p * k + kx) * C + c);
      if (add != nullptr) {
        F8 f = unpack8(r), g = unpack8(*reinterpret_cast<const uint4*>(add + off));
#pragma unroll
        for (int i = 0; i < 8; ++i) f.v[i] = g.v[i] + f.v[i];
        r = pack8(f);
      }
      *reinterpret_cast<uint4*>(din + off) = r;
      continue;
    }
    F8 acc;
    if (add != nullptr) acc = unpack8(*reinterpret_cast<const uint4*>(add + off));
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
    }
    for (int ky = 0; ky < k; ++ky) {
      const int ty = iy + p - ky;
      if (ty < 0 || ty % s != 0) continue;
      const int oy = ty / s;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int tx = ix + p - kx;
        if (tx < 0 || tx % s != 0) continue;
        const int ox = tx / s;
        if (ox >= Wo) continue;
        const F8 v = unpack8(*reinterpret_cast<const uint4*>(dc_b + ((long)oy * Wo + ox) * ldc + (long)(ky * k + kx) * C + c));
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] += v.v[i];
      }
    }
    *reinterpret_cast<uint4*>(din + off) = pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm (training)
// Per-channel reductions over a channels-last map y[P, C] (C <= 32, C % 4 == 0).  Thread = pixel; per-thread channel
// accumulators, shared-memory tree over the CTA, one double atomic per channel per CTA.
constexpr int BN_THREADS = 256;
constexpr int BN_MAXC = 32;

template <int NACC>
__device__ __forceinline__ void bn_block_reduce(float (&acc)[NACC][BN_MAXC], int C, double* out /*[NACC][C]*/) {
  __shared__ float red[BN_THREADS / 32][BN_MAXC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    for (int c = 0; c < C; ++c) {
      const float v = warp_sum(acc[a][c]);
      if (lane == 0) red[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < C) {
      float t = 0.f;
      for (int w = 0; w < BN_THREADS / 32; ++w) t += red[w][threadIdx.x];
      atomicAdd(out + a * C + threadIdx.x, (double)t);
    }
    __syncthreads();
  }
}

// sums[0][c] = sum y, sums[1][c] = sum y^2
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const T* __restrict__ y, long P, int C, double* __restrict__ sums) {
  pdl_launch_dependents();
  pdl_wait();
  float acc[2][BN_MAXC];
#pragma unroll
  for (int c = 0; c < BN_MAXC; ++c) acc[0][c] = acc[1][c] = 0.f;
  for (long pidx = (long)blockIdx.x * BN_THREADS + threadIdx.x; pidx < P; pidx += (long)gridDim.x * BN_THREADS) {
    const T* r = y + pidx * C;
#pragma unroll
    for (int c = 0; c < BN_MAXC; c += 4) {
      if (c < C) {
        const float4 v = ld4(r + c);
        acc[0][c] += v.x, acc[0][c + 1] += v.y, acc[0][c + 2] += v.z, acc[0][c + 3] += v.w;
        acc[1][c] += v.x * v.x, acc[1][c + 1] += v.y * v.y, acc[1][c + 2] += v.z * v.z, acc[1][c + 3] += v.w * v.w;
      }
    }
  }
  bn_block_reduce<2>(acc, C, sums);
}

// mean/var -> (scale, shift) for the fused apply, saved (mean, rstd) for the backward, running-stat update
// (momentum 0.1, unbiased variance), num_batches_tracked += 1.   One thread per channel.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, long P, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ tracked) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = threadIdx.x;
  if (c < C) {
    const double m = sums[c] / (double)P;
    double var = sums[C + c] / (double)P - m * m;
    if (var < 0) var = 0;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    scale[c] = gamma[c] * rs;
    shift[c] = beta[c] - (float)m * gamma[c] * rs;
    mean_out[c] = (float)m;
    rstd_out[c] = rs;
    if (running_mean != nullptr) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * (double)P / (double)(P - 1));
    }
  }
  if (c == 0 && tracked != nullptr) *tracked += 1;
}

// Backward reductions through a = relu(gamma*zhat + beta), zhat = (y-mean)*rstd:  dz = da * [a > 0];
// sums[0][c] = sum dz (= dbeta), sums[1][c] = sum dz*zhat (= dgamma)
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_stats_kernel(const T* __restrict__ da, const T* __restrict__ y, long P, int C,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  double* __restrict__ sums) {
  pdl_launch_dependents();
  pdl_wait();
  float acc[2][BN_MAXC];
#pragma unroll
  for (int c = 0; c < BN_MAXC; ++c) acc[0][c] = acc[1][c] = 0.f;
  for (long pidx = (long)blockIdx.x * BN_THREADS + threadIdx.x; pidx < P; pidx += (long)gridDim.x * BN_THREADS) {
#pragma unroll
    for (int c = 0; c < BN_MAXC; c += 4) {
      if (c < C) {
        const float4 yv = ld4(y + pidx * C + c), dv = ld4(da + pidx * C + c);
        const float yy[4] = {yv.x, yv.y, yv.z, yv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float zh = (yy[j] - mean[c + j]) * rstd[c + j];
          const float dz = (gamma[c + j] * zh + beta[c + j] > 0.f) ? dd[j] : 0.f;
          acc[0][c + j] += dz;
          acc[1][c + j] += dz * zh;
        }
      }
    }
  }
  bn_block_reduce<2>(acc, C, sums);
}

// dy = gamma*rstd * (dz - sum(dz)/P - zhat * sum(dz*zhat)/P); also emits dgamma/dbeta (fp32) from the sums (block 0).
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ da, const T* __restrict__ y, long P, int C, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const double* __restrict__ sums, T* __restrict__ dy, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta) {
  pdl_launch_dependents();
  pdl_wait();
  // per-channel constants once per CTA: zhat = y*a + b, pre-activation = gamma*zhat + beta, dy = s*(dz - m1 - zhat*m2)
  __shared__ float ca[BN_MAXC], cb[BN_MAXC], cg[BN_MAXC], cbeta[BN_MAXC], cs[BN_MAXC], cm1[BN_MAXC], cm2[BN_MAXC];
  const int c4n = C / 4;
  const long total = P * c4n;
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    const double invP = 1.0 / (double)P;
    ca[c] = rstd[c], cb[c] = -mean[c] * rstd[c];
    cg[c] = gamma[c], cbeta[c] = beta[c], cs[c] = gamma[c] * rstd[c];
    cm1[c] = (float)(sums[c] * invP), cm2[c] = (float)(sums[C + c] * invP);
    if (blockIdx.x == 0) {
      dbeta[c] += (float)sums[c];
      dgamma[c] += (float)sums[C + c];
    }
  }
  __syncthreads();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long pidx = i / c4n;
    const float4 yv = ld4(y + pidx * C + c), dv = ld4(da + pidx * C + c);
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float zh = fmaf(yy[j], ca[c + j], cb[c + j]);
      const float dz = fmaf(cg[c + j], zh, cbeta[c + j]) > 0.f ? dd[j] : 0.f;
      o[j] = cs[c + j] * (dz - cm1[c + j] - zh * cm2[c + j]);
    }
    st4(dy + pidx * C + c, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// ------------------------------------------------------------------------------------------------ embedding assembly
// x0[b, t, c] = [c < keep] * ((t == 0 ? tokens[c] : patches[b, t-1, c]) + pos[t, c])        (vit_sr_supernet.py:399-407)
__global__ void embed_assemble_kernel(const float* __restrict__ patches, const float* __restrict__ tokens, const float* __restrict__ pos,
                                      float* __restrict__ x0, int nb, int N, int C, int keep, int T) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = C / 4;
  const long total = (long)nb * N * c4n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long row = i / c4n;
    const int t = (int)(row % N);
    const long b = row / N;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < keep) {
      const float4 a = t < T ? ld4(tokens + (long)t * C + c) : ld4(patches + (b * (N - T) + (t - T)) * C + c);
      const float4 q = ld4(pos + (long)t * C + c);
      v = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
      if (c + 1 >= keep) v.y = 0.f;
      if (c + 2 >= keep) v.z = 0.f;
      if (c + 3 >= keep) v.w = 0.f;
    }
    st4(x0 + row * C + c, v);
  }
}

// Backward: dpatches[b,p,c] = [c<keep] g[b,NT+p,c] (cast to T);  dpos[t,c] += sum_b [c<keep] g[b,t,c];  dtokens[t,c] += sum_b g[b,t,c], t < NT.
// One thread per (t, 4 channels) loops over the segment's samples -> no atomics within a segment.
template <typename T>
__global__ void embed_assemble_bwd_kernel(const float* __restrict__ g, T* __restrict__ dpatches, float* __restrict__ dpos,
                                          float* __restrict__ dtokens, int nb, int N, int C, int keep, int NT) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = C / 4;
  const long total = (long)N * c4n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const int t = (int)(i / c4n);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int bper = (nb + gridDim.y - 1) / gridDim.y;       // blockIdx.y splits the batch; one atomic per chunk for dpos / dtokens
    const int b_lo = blockIdx.y * bper, b_hi = min(nb, b_lo + bper);
    if (b_lo >= b_hi) continue;
    for (int b = b_lo; b < b_hi; ++b) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < keep) {
        v = ld4(g + ((long)b * N + t) * C + c);
        if (c + 1 >= keep) v.y = 0.f;
        if (c + 2 >= keep) v.z = 0.f;
        if (c + 3 >= keep) v.w = 0.f;
      }
      if (t >= NT) st4(dpatches + ((long)b * (N - NT) + (t - NT)) * C + c, v);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    float* dp = dpos + (long)t * C + c;
    atomicAdd(dp, acc.x), atomicAdd(dp + 1, acc.y), atomicAdd(dp + 2, acc.z), atomicAdd(dp + 3, acc.w);
    if (t < NT) {
      float* dt_ = dtokens + (long)t * C + c;
      atomicAdd(dt_, acc.x), atomicAdd(dt_ + 1, acc.y), atomicAdd(dt_ + 2, acc.z), atomicAdd(dt_ + 3, acc.w);
    }
  }
}

// ------------------------------------------------------------------------------------------------ SR combine
// y[b,0,c]   = [c<keep2] * (tok[b,c] + (c<C1 ? x[b,0,c] : 0))
// y[b,1+p,c] = [c<keep2] * (conv[b,p,c] + pos[p,c] + (c<C1 ? mean of the 2x2 block of x patch rows : 0))     (vit_sr_supernet.py:131-166)
__global__ void sr_combine_kernel(const float* __restrict__ conv, const float* __restrict__ tok, const float* __restrict__ pos,
                                  const float* __restrict__ x, float* __restrict__ y, int nb, int g, int C1, int C2, int keep2) {
  pdl_launch_dependents();
  pdl_wait();
  const int g2 = g / 2, N2 = 1 + g2 * g2, N1 = 1 + g * g;
  const int c4n = C2 / 4;
  const long total = (long)nb * N2 * c4n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long row = i / c4n;
    const int t = (int)(row % N2);
    const long b = row / N2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < keep2) {
      if (t == 0) {
        v = ld4(tok + b * C2 + c);
        if (c < C1) {
          const float4 r = ld4(x + b * N1 * C1 + c);
          v.x += r.x, v.y += r.y, v.z += r.z, v.w += r.w;
        }
      } else {
        const int pidx = t - 1, oy = pidx / g2, ox = pidx % g2;
        v = ld4(conv + (b * (N2 - 1) + pidx) * C2 + c);
        const float4 q = ld4(pos + (long)pidx * C2 + c);
        v.x += q.x, v.y += q.y, v.z += q.z, v.w += q.w;
        if (c < C1) {
          float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const float4 r = ld4(x + (b * N1 + 1 + (2 * oy + dy) * g + (2 * ox + dx)) * C1 + c);
              s4.x += r.x, s4.y += r.y, s4.z += r.z, s4.w += r.w;
            }
          v.x += 0.25f * s4.x, v.y += 0.25f * s4.y, v.z += 0.25f * s4.z, v.w += 0.25f * s4.w;
        }
      }
      if (c + 1 >= keep2) v.y = 0.f;
      if (c + 2 >= keep2) v.z = 0.f;
      if (c + 3 >= keep2) v.w = 0.f;
    }
    st4(y + row * C2 + c, v);
  }
}

// Backward of the combine.  gm = [c<keep2] * g.   dconv[b,p,:] = gm[b,1+p,:] (T), dtok[b,:] = gm[b,0,:] (T),
// dpos[p,:] += sum_b gm[b,1+p,:], and the residual path into gres[b, 1+g*g, C1] (fp32):
// gres[b,0,c] = gm[b,0,c]; gres[b,1+(y,x),c] = 0.25*gm[b,1+(y/2,x/2),c]  for c < C1.
template <typename T>
__global__ void sr_combine_bwd_kernel(const float* __restrict__ gy, T* __restrict__ dconv, T* __restrict__ dtok, float* __restrict__ dpos,
                                      float* __restrict__ gres, int nb, int g, int C1, int C2, int keep2) {
  pdl_launch_dependents();
  pdl_wait();
  const int g2 = g / 2, N2 = 1 + g2 * g2, N1 = 1 + g * g;
  const int c4n = C2 / 4;
  const long total = (long)N2 * c4n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const int t = (int)(i / c4n);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // blockIdx.y splits the batch (the position-embedding gradient is the only cross-sample reduction: one atomic per chunk)
    const int bper = (nb + gridDim.y - 1) / gridDim.y;
    const int b_lo = blockIdx.y * bper, b_hi = min(nb, b_lo + bper);
    for (int b = b_lo; b < b_hi; ++b) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < keep2) {
        v = ld4(gy + ((long)b * N2 + t) * C2 + c);
        if (c + 1 >= keep2) v.y = 0.f;
        if (c + 2 >= keep2) v.z = 0.f;
        if (c + 3 >= keep2) v.w = 0.f;
      }
      if (t == 0) {
        st4(dtok + (long)b * C2 + c, v);
        if (c < C1) st4(gres + (long)b * N1 * C1 + c, v);
      } else {
        const int pidx = t - 1, oy = pidx / g2, ox = pidx % g2;
        st4(dconv + ((long)b * (N2 - 1) + pidx) * C2 + c, v);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        if (c < C1) {
          const float4 q = make_float4(0.25f * v.x, 0.25f * v.y, 0.25f * v.z, 0.25f * v.w);
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) st4(gres + ((long)b * N1 + 1 + (2 * oy + dy) * g + (2 * ox + dx)) * C1 + c, q);
        }
      }
    }
    if (t > 0 && b_hi > b_lo) {
      float* dp = dpos + (long)(t - 1) * C2 + c;
      atomicAdd(dp, acc.x), atomicAdd(dp + 1, acc.y), atomicAdd(dp + 2, acc.z), atomicAdd(dp + 3, acc.w);
    }
  }
}

int grid_for(long items) {
  long gsz = ceil_div_l(items, 256);
  const long cap = (long)num_sms() * 16;
  return (int)(gsz < cap ? (gsz > 0 ? gsz : 1) : cap);
}

}  // namespace
}  // namespace vsx

using namespace vsx;
#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int vsx_im2col(const void* in1, const float* scale1, const float* shift1, const void* in2, const float* scale2,
                          const float* shift2, int in_dtype, int nchw, long batch_pitch, long pix_pitch, int B, int H, int W, int C,
                          int k, int stride, int pad, void* out, int out_dtype, long ldo, void* stream) {
  VSX_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0, "vsx_im2col: bad shape");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  VSX_REQUIRE(ldo >= (long)k * k * C, "vsx_im2col: ldo too small");
  if (B == 0) return VSX_OK;
  const int grid = grid_for((long)B * Ho * Wo * k * k);
  if (nchw) {
    VSX_REQUIRE(in_dtype == VSX_F32 && in2 == nullptr && scale1 == nullptr, "vsx_im2col: NCHW input is the fp32 image, no fused activation");
    if (out_dtype == VSX_BF16 && C == 3 && k == 3 && stride == 2 && pad == 1 && ldo == 32 && batch_pitch == 3L * H * W &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      launch_pdl(im2col_conv1_kernel, dim3(grid_for((long)B * Ho * Wo)), dim3(256), 0, ST, (const float*)in1, B, H, W, Ho, Wo, (bf16*)out);
      return check_launch("vsx_im2col");
    }
    if (out_dtype == VSX_BF16)
      launch_pdl(im2col_kernel<float, bf16, true>, dim3(grid), dim3(256), 0, ST, (const float*)in1, nullptr, nullptr, nullptr, nullptr, nullptr, batch_pitch,
                                                              pix_pitch, B, H, W, C, k, stride, pad, Ho, Wo, (bf16*)out, ldo);
    else
      launch_pdl(im2col_kernel<float, float, true>, dim3(grid), dim3(256), 0, ST, (const float*)in1, nullptr, nullptr, nullptr, nullptr, nullptr, batch_pitch,
                                                               pix_pitch, B, H, W, C, k, stride, pad, Ho, Wo, (float*)out, ldo);
  } else {
    VSX_REQUIRE(C % 4 == 0 && pix_pitch % 4 == 0 && batch_pitch % 4 == 0 && ldo % 4 == 0, "vsx_im2col: channels-last needs C, pitches %% 4 == 0");
    VSX_REQUIRE(in_dtype == out_dtype, "vsx_im2col: channels-last input and output share the activation dtype");
    const bool al16 = ((reinterpret_cast<uintptr_t>(in1) | reinterpret_cast<uintptr_t>(in2) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    // narrow maps (the 24-channel stem: conv_proj's 7x7 patches) keep the generic kernel: one thread per (pixel, tap) gives ~3 M independent
    // 48-byte copies, which measured faster inside the step (119 us) than one warp per run of taps (180 us)
    if (out_dtype == VSX_BF16 && C % 8 == 0 && C > 32 && pix_pitch % 8 == 0 && batch_pitch % 8 == 0 && ldo % 8 == 0 && al16 && pix_pitch < (1L << 30))
      launch_pdl(im2col_rows_kernel, dim3(B * Ho), dim3(256), 0, ST, (const bf16*)in1, scale1, shift1, (const bf16*)in2, scale2, shift2, batch_pitch, (int)pix_pitch, H,
                                                 W, C, k, stride, pad, Ho, Wo, (bf16*)out, ldo);
    else if (out_dtype == VSX_BF16)
      launch_pdl(im2col_kernel<bf16, bf16, false>, dim3(grid), dim3(256), 0, ST, (const bf16*)in1, scale1, shift1, (const bf16*)in2, scale2, shift2, batch_pitch,
                                                              pix_pitch, B, H, W, C, k, stride, pad, Ho, Wo, (bf16*)out, ldo);
    else
      launch_pdl(im2col_kernel<float, float, false>, dim3(grid), dim3(256), 0, ST, (const float*)in1, scale1, shift1, (const float*)in2, scale2, shift2,
                                                                batch_pitch, pix_pitch, B, H, W, C, k, stride, pad, Ho, Wo, (float*)out, ldo);
  }
  return check_launch("vsx_im2col");
}

extern "C" int vsx_col2im(const void* dcol, long ldc, const void* add, int dtype, int B, int H, int W, int C, int k, int stride, int pad,
                          void* din, long batch_pitch, long pix_pitch, void* stream) {
  VSX_REQUIRE(C % 4 == 0 && ldc % 4 == 0 && pix_pitch % 4 == 0 && batch_pitch % 4 == 0, "vsx_col2im: C and pitches must be multiples of 4");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  if (B == 0) return VSX_OK;
  const int grid = grid_for((long)B * H * W * C / 4);
  const bool al16 = ((reinterpret_cast<uintptr_t>(dcol) | reinterpret_cast<uintptr_t>(add) | reinterpret_cast<uintptr_t>(din)) & 15) == 0;
  if (dtype == VSX_BF16 && C % 8 == 0 && ldc % 8 == 0 && pix_pitch % 8 == 0 && batch_pitch % 8 == 0 && al16 && pix_pitch < (1L << 30))
    launch_pdl(col2im_rows_kernel, dim3(B * H), dim3(256), 0, ST, (const bf16*)dcol, ldc, (const bf16*)add, H, W, C, k, stride, pad, Ho, Wo, (bf16*)din, batch_pitch,
                                              (int)pix_pitch);
  else if (dtype == VSX_BF16)
    launch_pdl(col2im_kernel<bf16>, dim3(grid), dim3(256), 0, ST, (const bf16*)dcol, ldc, (const bf16*)add, B, H, W, C, k, stride, pad, Ho, Wo, (bf16*)din,
                                              batch_pitch, pix_pitch);
  else
    launch_pdl(col2im_kernel<float>, dim3(grid), dim3(256), 0, ST, (const float*)dcol, ldc, (const float*)add, B, H, W, C, k, stride, pad, Ho, Wo, (float*)din,
                                               batch_pitch, pix_pitch);
  return check_launch("vsx_col2im");
}

extern "C" int vsx_bn_stats(const void* y, int dtype, long P, int C, double* sums, void* stream) {
  VSX_REQUIRE(C % 4 == 0 && C <= BN_MAXC, "vsx_bn_stats: C must be a multiple of 4 and <= 32 (got %d)", C);
  const int grid = (int)std::min<long>(ceil_div_l(P, BN_THREADS), (long)num_sms() * 4);
  if (dtype == VSX_BF16) launch_pdl(bn_stats_kernel<bf16>, dim3(grid), dim3(BN_THREADS), 0, ST, (const bf16*)y, P, C, sums);
  else launch_pdl(bn_stats_kernel<float>, dim3(grid), dim3(BN_THREADS), 0, ST, (const float*)y, P, C, sums);
  return check_launch("vsx_bn_stats");
}

extern "C" int vsx_bn_finalize(const double* sums, long P, int C, const float* gamma, const float* beta, float eps, float momentum,
                               float* scale, float* shift, float* mean, float* rstd, float* running_mean, float* running_var,
                               long long* num_batches_tracked, void* stream) {
  VSX_REQUIRE(C <= BN_MAXC && P > 1, "vsx_bn_finalize: bad C/P");
  launch_pdl(bn_finalize_kernel, dim3(1), dim3(32), 0, ST, sums, P, C, gamma, beta, eps, momentum, scale, shift, mean, rstd, running_mean, running_var,
                                      num_batches_tracked);
  return check_launch("vsx_bn_finalize");
}

extern "C" int vsx_bn_bwd_stats(const void* da, const void* y, int dtype, long P, int C, const float* gamma, const float* beta,
                                const float* mean, const float* rstd, double* sums, void* stream) {
  VSX_REQUIRE(C % 4 == 0 && C <= BN_MAXC, "vsx_bn_bwd_stats: C must be a multiple of 4 and <= 32 (got %d)", C);
  const int grid = (int)std::min<long>(ceil_div_l(P, BN_THREADS), (long)num_sms() * 4);
  if (dtype == VSX_BF16)
    launch_pdl(bn_bwd_stats_kernel<bf16>, dim3(grid), dim3(BN_THREADS), 0, ST, (const bf16*)da, (const bf16*)y, P, C, gamma, beta, mean, rstd, sums);
  else
    launch_pdl(bn_bwd_stats_kernel<float>, dim3(grid), dim3(BN_THREADS), 0, ST, (const float*)da, (const float*)y, P, C, gamma, beta, mean, rstd, sums);
  return check_launch("vsx_bn_bwd_stats");
}

extern "C" int vsx_bn_bwd_apply(const void* da, const void* y, int dtype, long P, int C, const float* gamma, const float* beta,
                                const float* mean, const float* rstd, const double* sums, void* dy, float* dgamma, float* dbeta,
                                void* stream) {
  VSX_REQUIRE(C % 4 == 0 && C <= BN_MAXC, "vsx_bn_bwd_apply: C must be a multiple of 4 and <= 32 (got %d)", C);
  const int grid = grid_for(P * C / 4);
  if (dtype == VSX_BF16)
    launch_pdl(bn_bwd_apply_kernel<bf16>, dim3(grid), dim3(256), 0, ST, (const bf16*)da, (const bf16*)y, P, C, gamma, beta, mean, rstd, sums, (bf16*)dy, dgamma, dbeta);
  else
    launch_pdl(bn_bwd_apply_kernel<float>, dim3(grid), dim3(256), 0, ST, (const float*)da, (const float*)y, P, C, gamma, beta, mean, rstd, sums, (float*)dy, dgamma, dbeta);
  return check_launch("vsx_bn_bwd_apply");
}

extern "C" int vsx_embed_assemble(const float* patches, const float* tokens, const float* pos, float* x0, int batch, int tokens_per_sample,
                                  int C, int keep, int num_tokens, void* stream) {
  VSX_REQUIRE(C % 4 == 0 && keep >= 0 && keep <= C, "vsx_embed_assemble: bad C/keep");
  VSX_REQUIRE(num_tokens >= 1 && num_tokens < tokens_per_sample, "vsx_embed_assemble: bad num_tokens %d", num_tokens);
  if (batch == 0) return VSX_OK;
  launch_pdl(embed_assemble_kernel, dim3(grid_for((long)batch * tokens_per_sample * C / 4)), dim3(256), 0, ST, patches, tokens, pos, x0, batch, tokens_per_sample, C, keep,
                                                                                            num_tokens);
  return check_launch("vsx_embed_assemble");
}

extern "C" int vsx_embed_assemble_bwd(const float* g, void* dpatches, int dtype, float* dpos, float* dtokens, int batch,
                                      int tokens_per_sample, int C, int keep, int num_tokens, void* stream) {
  VSX_REQUIRE(C % 4 == 0 && keep >= 0 && keep <= C, "vsx_embed_assemble_bwd: bad C/keep");
  VSX_REQUIRE(num_tokens >= 1 && num_tokens < tokens_per_sample, "vsx_embed_assemble_bwd: bad num_tokens %d", num_tokens);
  if (batch == 0) return VSX_OK;
  const int gx = grid_for((long)tokens_per_sample * C / 4);
  int gyy = (4 * num_sms() + gx - 1) / gx;
  gyy = gyy > batch ? batch : (gyy < 1 ? 1 : gyy);
  const dim3 grid(gx, gyy);
  if (dtype == VSX_BF16) launch_pdl(embed_assemble_bwd_kernel<bf16>, dim3(grid), dim3(256), 0, ST, g, (bf16*)dpatches, dpos, dtokens, batch, tokens_per_sample, C, keep, num_tokens);
  else launch_pdl(embed_assemble_bwd_kernel<float>, dim3(grid), dim3(256), 0, ST, g, (float*)dpatches, dpos, dtokens, batch, tokens_per_sample, C, keep, num_tokens);
  return check_launch("vsx_embed_assemble_bwd");
}

extern "C" int vsx_sr_combine(const float* conv, const float* tok, const float* pos, const float* x, float* y, int batch, int grid_in,
                              int C1, int C2, int keep2, void* stream) {
  VSX_REQUIRE(C1 % 4 == 0 && C2 % 4 == 0 && C2 >= C1 && grid_in % 2 == 0 && keep2 >= 0 && keep2 <= C2, "vsx_sr_combine: bad shape");
  if (batch == 0) return VSX_OK;
  const int N2 = 1 + (grid_in / 2) * (grid_in / 2);
  launch_pdl(sr_combine_kernel, dim3(grid_for((long)batch * N2 * C2 / 4)), dim3(256), 0, ST, conv, tok, pos, x, y, batch, grid_in, C1, C2, keep2);
  return check_launch("vsx_sr_combine");
}

extern "C" int vsx_sr_combine_bwd(const float* gy, void* dconv, void* dtok, int dtype, float* dpos, float* gres, int batch, int grid_in,
                                  int C1, int C2, int keep2, void* stream) {
  VSX_REQUIRE(C1 % 4 == 0 && C2 % 4 == 0 && C2 >= C1 && grid_in % 2 == 0 && keep2 >= 0 && keep2 <= C2, "vsx_sr_combine_bwd: bad shape");
  if (batch == 0) return VSX_OK;
  const int N2 = 1 + (grid_in / 2) * (grid_in / 2);
  const int gx = grid_for((long)N2 * C2 / 4);
  int gy_ = (4 * num_sms() + gx - 1) / gx;          // enough batch chunks for ~4 CTAs per SM
  if (gy_ > batch) gy_ = batch;
  if (gy_ < 1) gy_ = 1;
  const dim3 grid(gx, gy_);
  if (dtype == VSX_BF16) launch_pdl(sr_combine_bwd_kernel<bf16>, dim3(grid), dim3(256), 0, ST, gy, (bf16*)dconv, (bf16*)dtok, dpos, gres, batch, grid_in, C1, C2, keep2);
  else launch_pdl(sr_combine_bwd_kernel<float>, dim3(grid), dim3(256), 0, ST, gy, (float*)dconv, (float*)dtok, dpos, gres, batch, grid_in, C1, C2, keep2);
  return check_launch("vsx_sr_combine_bwd");
}
