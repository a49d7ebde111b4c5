// Small HBM-bound helper kernels around the GEMMs: operand casts / bf16 hi-lo splits, the masked + drop-path-scaled
// gradient cast that opens each branch's backward, and bias-gradient column sums.
#include <string.h>

#include "common.cuh"

namespace vsx {
namespace {

// hi = bf16(src); lo = bf16(src - hi) (optional); lo2 = bf16(src - hi - lo) (optional).  4 elements per thread.
__global__ void split_kernel(const float* __restrict__ src, long lds, bf16* __restrict__ hi, bf16* __restrict__ lo, bf16* __restrict__ lo2,
                             long ldd, int rows, int cols4) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)rows * cols4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / cols4;
    const int c = (int)(i - r * cols4) * 4;
    const float4 v4 = ld4(src + r * lds + c);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    float h[4], l[4] = {0.f, 0.f, 0.f, 0.f}, l2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
      if (lo != nullptr) l[j] = __bfloat162float(__float2bfloat16_rn(v[j] - h[j]));
      l2[j] = (v[j] - h[j]) - l[j];
    }
    st4(hi + r * ldd + c, make_float4(h[0], h[1], h[2], h[3]));
    if (lo != nullptr) st4(lo + r * ldd + c, make_float4(l[0], l[1], l[2], l[3]));
    if (lo2 != nullptr) st4(lo2 + r * ldd + c, make_float4(l2[0], l2[1], l2[2], l2[3]));
  }
}

// out[m, n] = n < n_keep ? g[m, n] * row_scale[m / rows_per_sample] : 0     (Block backward, nets/supernet_blocks.py:243,251
// together with nets/drop.py:25: the branch output was scaled by drop-path and masked before the residual add).
// Optionally colsum[n] += sum_m out[m, n] -- the bias gradient of the Linear that produced the branch output -- from the same pass.
// One warp per row, lane owns columns (i*32 + lane)*4..+3, i < NV (like the LayerNorm kernels).
constexpr int SMC_WARPS = 8;
template <int NV, typename T>
__global__ void __launch_bounds__(SMC_WARPS * 32) scale_mask_cast_kernel(const float* __restrict__ g, long ldg, const float* __restrict__ row_scale,
                                                                        int rps, int n_keep, T* __restrict__ out, long ldo, int rows, int cols,
                                                                        float* __restrict__ colsum) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long r = (long)blockIdx.x * SMC_WARPS + warp; r < rows; r += (long)gridDim.x * SMC_WARPS) {
    const float s = row_scale != nullptr ? __ldg(row_scale + r / rps) : 1.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < cols) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < n_keep) {
          v = ld4(g + r * ldg + c);
          v.x *= s;
          v.y = (c + 1 < n_keep) ? v.y * s : 0.f;
          v.z = (c + 2 < n_keep) ? v.z * s : 0.f;
          v.w = (c + 3 < n_keep) ? v.w * s : 0.f;
        }
        st4(out + r * ldo + c, v);
        acc[i].x += v.x, acc[i].y += v.y, acc[i].z += v.z, acc[i].w += v.w;
      }
    }
  }
  if (colsum == nullptr) return;
  // one barrier: every warp parks its partials, then thread j sums column-quad j over the warps
  __shared__ float4 red[SMC_WARPS][NV * 32];
#pragma unroll
  for (int i = 0; i < NV; ++i) red[warp][i * 32 + lane] = acc[i];
  __syncthreads();
  for (int q = threadIdx.x; q < NV * 32; q += SMC_WARPS * 32) {
    float4 t = red[0][q];
#pragma unroll
    for (int w = 1; w < SMC_WARPS; ++w) {
      const float4 u = red[w][q];
      t.x += u.x, t.y += u.y, t.z += u.z, t.w += u.w;
    }
    red_add4(colsum + q * 4, t, q * 4, n_keep);
  }
}


// Same operation with bulk-copy row prefetch (see ln_bwd_bulk_kernel in norm.cu): g rows arrive in shared memory one iteration ahead.
template <int NV, typename T>
__global__ void __launch_bounds__(SMC_WARPS * 32) scale_mask_cast_bulk_kernel(const float* __restrict__ g, long ldg, const float* __restrict__ row_scale,
                                                                             int rps, int n_keep, T* __restrict__ out, long ldo, int rows, int cols,
                                                                             float* __restrict__ colsum, const RowSegs segs) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t smc_smem[];      // n_keep = the largest kept width of the launch when there are segments
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gb = (uint32_t)n_keep * 4;
  const uint32_t slot = (gb + 127u) & ~127u;
  uint8_t* my = smc_smem + (size_t)warp * 2 * slot;
  __shared__ __align__(8) unsigned long long bars[SMC_WARPS][2];
  const uint32_t bar0 = smem_u32(&bars[warp][0]), bar1 = smem_u32(&bars[warp][1]);
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    fence_mbar_init();
  }
  __syncwarp();
  const long stride = (long)gridDim.x * SMC_WARPS;
  auto issue = [&](long r, int sl) {
    const uint32_t bar = sl ? bar1 : bar0;
    const uint32_t gr = (uint32_t)(segs.count ? segs.keep[seg_of_row(segs, r)] : n_keep) * 4;
    mbar_expect_tx(bar, gr);             // 0 bytes for a skipped row: the phase completes at once
    if (gr) bulk_g2s(smem_u32(my + (size_t)sl * slot), g + r * ldg, gr, bar);
  };
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  long r = (long)blockIdx.x * SMC_WARPS + warp;
  if (r < rows && lane == 0) issue(r, 0);
  int it = 0;
  for (; r < rows; r += stride, ++it) {
    const int sl = it & 1;
    if (r + stride < rows && lane == 0) issue(r + stride, sl ^ 1);
    const float s = row_scale != nullptr ? __ldg(row_scale + r / rps) : 1.0f;
    const int keep_r = segs.count ? segs.keep[seg_of_row(segs, r)] : n_keep;
    mbar_wait(sl ? bar1 : bar0, (uint32_t)(it >> 1) & 1u);
    if (keep_r == 0) {
      __syncwarp();
      continue;
    }
    const float* gs = reinterpret_cast<const float*>(my + (size_t)sl * slot);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < cols) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < keep_r) {
          v = ld4(gs + c);
          v.x *= s, v.y *= s, v.z *= s, v.w *= s;
        }
        st4(out + r * ldo + c, v);
        acc[i].x += v.x, acc[i].y += v.y, acc[i].z += v.z, acc[i].w += v.w;
      }
    }
    __syncwarp();
  }
  if (colsum == nullptr) return;
  __shared__ float4 red[SMC_WARPS][NV * 32];
#pragma unroll
  for (int i = 0; i < NV; ++i) red[warp][i * 32 + lane] = acc[i];
  __syncthreads();
  for (int q = threadIdx.x; q < NV * 32; q += SMC_WARPS * 32) {
    float4 t = red[0][q];
#pragma unroll
    for (int w = 1; w < SMC_WARPS; ++w) {
      const float4 u = red[w][q];
      t.x += u.x, t.y += u.y, t.z += u.z, t.w += u.w;
    }
    red_add4(colsum + q * 4, t, q * 4, n_keep);
  }
}

// out[c] += sum_r x[r, c].  CTA = 256 threads = 32 column-quads x 8 row lanes, covers 128 columns x ROWS_PER_CTA rows.
// uint8 image batch [images, channels, H*W] -> fp32 (v / 255 - mean[c]) / std[c]: the ToTensor + Normalize of the data pipeline done on
// the device, so that a step uploads 1 byte per pixel instead of 4.  16 pixels per thread (one 16-byte load, four 16-byte stores).
__global__ void __launch_bounds__(256) image_normalize_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, long plane16, int channels,
                                                                 long planes, float m0, float m1, float m2, float m3, float s0, float s1, float s2,
                                                                 float s3) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = planes * plane16;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)((i / plane16) % channels);
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : (c == 2 ? m2 : m3)), inv = c == 0 ? s0 : (c == 1 ? s1 : (c == 2 ? s2 : s3));
    const float a = inv * (1.0f / 255.0f), b = -mean * inv;
    const uint4 v = *reinterpret_cast<const uint4*>(in + i * 16);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 o;
      o.x = fmaf((float)(w[k] & 255u), a, b), o.y = fmaf((float)((w[k] >> 8) & 255u), a, b);
      o.z = fmaf((float)((w[k] >> 16) & 255u), a, b), o.w = fmaf((float)(w[k] >> 24), a, b);
      *reinterpret_cast<float4*>(out + i * 16 + k * 4) = o;
    }
  }
}

// 128 rows per CTA (16 per warp, all loads of a thread in flight at once): with 512 rows per CTA a [16384 x 512] sum ran 128 CTAs whose
// warps each walked 64 dependent loads -- 22 us for 17 MB (tools: ncu launch list r2y); now ~4 us.
constexpr int CS_ROWS = 128;
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long ldx, int rows, int cols, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + cq * 4;
  const long r0 = (long)blockIdx.y * CS_ROWS;
  const long r1 = r0 + CS_ROWS < rows ? r0 + CS_ROWS : rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols) {
#pragma unroll 8
    for (long r = r0 + rl; r < r1; r += 8) {
      const float4 v = ld4(x + r * ldx + c);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  }
  __shared__ float4 red[8][32];
  red[rl][cq] = acc;
  __syncthreads();
  if (rl == 0 && c < cols) {
    for (int k = 1; k < 8; ++k) {
      const float4 u = red[k][cq];
      acc.x += u.x, acc.y += u.y, acc.z += u.z, acc.w += u.w;
    }
    atomicAdd(out + c, acc.x);
    if (c + 1 < cols) atomicAdd(out + c + 1, acc.y);
    if (c + 2 < cols) atomicAdd(out + c + 2, acc.z);
    if (c + 3 < cols) atomicAdd(out + c + 3, acc.w);
  }
}

int ew_grid(long work_items) {
  long g = ceil_div_l(work_items, 256);
  const long cap = (long)num_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace
}  // namespace vsx


using namespace vsx;

extern "C" int vsx_split_bf16(const float* src, long lds, void* hi, void* lo, void* lo2, long ldd, int rows, int cols, void* stream) {
  VSX_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "vsx_split_bf16: cols and pitches must be multiples of 4");
  if (rows <= 0 || cols <= 0) return VSX_OK;
  launch_pdl(split_kernel, dim3(ew_grid((long)rows * cols / 4)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), src, lds, (bf16*)hi, (bf16*)lo, (bf16*)lo2,
                                                                                                 ldd, rows, cols / 4);
  return check_launch("vsx_split_bf16");
}

template <typename T>
static int smc_dispatch(const float* g, long ldg, const float* row_scale, int rps, int n_keep, void* out, long ldo, int rows, int cols,
                        float* colsum, cudaStream_t st, const RowSegs* segs = nullptr) {
  const int nv = ceil_div(cols, 128);
  bool seg_ok = true;
  if (segs != nullptr) {        // several extents in one launch: bulk variant only; returns 1 ("not handled") when it does not apply
    n_keep = 0;
    for (int i = 0; i < segs->count; ++i) seg_ok = seg_ok && segs->keep[i] % 4 == 0, n_keep = segs->keep[i] > n_keep ? segs->keep[i] : n_keep;
  }
  const int need = ceil_div(rows, SMC_WARPS), cap = num_sms() * (colsum != nullptr ? 4 : 8);
  const int grid = need < cap ? need : cap;
  if (seg_ok && n_keep > 0 && n_keep % 4 == 0 && ldg % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 && nv <= 10) {
    const size_t smem = (((size_t)n_keep * 4 + 127) & ~(size_t)127) * 2 * SMC_WARPS;
    const int cap4 = num_sms() * 4, gridb = need < cap4 ? need : cap4;
#define VSX_SMCB(NV)                                                                                                                       \
  case NV: {                                                                                                                               \
    static bool cfg = false;                                                                                                               \
    if (!cfg) {                                                                                                                            \
      cudaFuncSetAttribute(scale_mask_cast_bulk_kernel<NV, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);                    \
      cfg = true;                                                                                                                          \
    }                                                                                                                                      \
    launch_pdl(scale_mask_cast_bulk_kernel<NV, T>, dim3(gridb), dim3(SMC_WARPS * 32), smem, st, g, ldg, row_scale, rps, n_keep, (T*)out, ldo, rows, cols, colsum, segs ? *segs : RowSegs{}); \
    return check_launch("vsx_scale_mask_cast");                                                                                            \
  }
    switch (nv) {
      VSX_SMCB(1) VSX_SMCB(2) VSX_SMCB(3) VSX_SMCB(4) VSX_SMCB(5) VSX_SMCB(6) VSX_SMCB(7) VSX_SMCB(8) VSX_SMCB(9) VSX_SMCB(10)
      default: break;
    }
#undef VSX_SMCB
  }
  if (segs != nullptr) return 1;
#define VSX_SMC(NV)                                                                                                                 \
  case NV:                                                                                                                          \
    launch_pdl(scale_mask_cast_kernel<NV, T>, dim3(grid), dim3(SMC_WARPS * 32), 0, st, g, ldg, row_scale, rps, n_keep, (T*)out, ldo, rows, cols, colsum); \
    break;
  switch (nv) {
    VSX_SMC(1) VSX_SMC(2) VSX_SMC(3) VSX_SMC(4) VSX_SMC(5) VSX_SMC(6) VSX_SMC(7) VSX_SMC(8) VSX_SMC(9) VSX_SMC(10)
    default:
      set_error("vsx_scale_mask_cast: cols=%d exceeds the supported 1280", cols);
      return VSX_ERR_ARG;
  }
#undef VSX_SMC
  return check_launch("vsx_scale_mask_cast");
}

extern "C" int vsx_scale_mask_cast(const float* g, long ldg, const float* row_scale, int rows_per_sample, int n_keep, void* out,
                                   int dtype, long ldo, int rows, int cols, float* colsum, void* stream) {
  VSX_REQUIRE(cols % 4 == 0 && ldg % 4 == 0 && ldo % 4 == 0, "vsx_scale_mask_cast: cols and pitches must be multiples of 4");
  if (rows <= 0 || cols <= 0) return VSX_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rps = rows_per_sample > 0 ? rows_per_sample : 1;
  if (dtype == VSX_BF16) return smc_dispatch<bf16>(g, ldg, row_scale, rps, n_keep, out, ldo, rows, cols, colsum, st);
  if (dtype == VSX_F32) return smc_dispatch<float>(g, ldg, row_scale, rps, n_keep, out, ldo, rows, cols, colsum, st);
  set_error("vsx_scale_mask_cast: bad dtype %d", dtype);
  return VSX_ERR_ARG;
}

extern "C" int vsx_colsum(const void* x, int dtype, long ldx, int rows, int cols, float* out, void* stream) {
  VSX_REQUIRE(ldx % 4 == 0, "vsx_colsum: pitch must be a multiple of 4");
  if (rows <= 0 || cols <= 0) return VSX_OK;
  VSX_REQUIRE(cols % 4 == 0 || cols <= ldx, "vsx_colsum: bad cols");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(ceil_div(cols, 128), ceil_div(rows, CS_ROWS));
  if (dtype == VSX_BF16)
    launch_pdl(colsum_kernel<bf16>, dim3(grid), dim3(256), 0, st, (const bf16*)x, ldx, rows, cols, out);
  else if (dtype == VSX_F32)
    launch_pdl(colsum_kernel<float>, dim3(grid), dim3(256), 0, st, (const float*)x, ldx, rows, cols, out);
  else {
    set_error("vsx_colsum: bad dtype %d", dtype);
    return VSX_ERR_ARG;
  }
  return check_launch("vsx_colsum");
}

// Several extents in one launch (multi-architecture batches): rows of segment i keep segs->keep[i] channels; keep 0 = rows untouched.
extern "C" int vsx_scale_mask_cast_segs(const float* g, long ldg, const float* row_scale, int rows_per_sample, void* out, int dtype, long ldo,
                                        int rows, int cols, const vsx_row_segments* segs, float* colsum, void* stream) {
  VSX_REQUIRE(cols % 4 == 0 && ldg % 4 == 0 && ldo % 4 == 0, "vsx_scale_mask_cast_segs: cols and pitches must be multiples of 4");
  VSX_REQUIRE(segs != nullptr && segs->count >= 1 && segs->count <= VSX_MAX_SEGMENTS && segs->row_end[segs->count - 1] == rows,
              "vsx_scale_mask_cast_segs: 1..%d segments covering the %d rows", VSX_MAX_SEGMENTS, rows);
  VSX_REQUIRE(dtype == VSX_BF16 || dtype == VSX_F32, "vsx_scale_mask_cast_segs: bad dtype %d", dtype);
  if (rows <= 0 || cols <= 0) return VSX_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rps = rows_per_sample > 0 ? rows_per_sample : 1;
  RowSegs sg;
  static_assert(sizeof(RowSegs) == sizeof(vsx_row_segments), "RowSegs mirrors vsx_row_segments");
  memcpy(&sg, segs, sizeof(sg));
  int rc = dtype == VSX_BF16 ? smc_dispatch<bf16>(g, ldg, row_scale, rps, 0, out, ldo, rows, cols, colsum, st, &sg)
                             : smc_dispatch<float>(g, ldg, row_scale, rps, 0, out, ldo, rows, cols, colsum, st, &sg);
  if (rc != 1) return rc;
  const size_t es = dtype == VSX_BF16 ? 2 : 4;
  for (int i = 0, r0 = 0; i < segs->count; r0 = segs->row_end[i], ++i) {
    if (segs->keep[i] == 0 || segs->row_end[i] == r0) continue;
    rc = vsx_scale_mask_cast(g + (long)r0 * ldg, ldg, row_scale != nullptr ? row_scale + r0 / rps : nullptr, rps, segs->keep[i],
                             static_cast<uint8_t*>(out) + (size_t)r0 * ldo * es, dtype, ldo, segs->row_end[i] - r0, cols, colsum, stream);
    if (rc) return rc;
  }
  return VSX_OK;
}

// Device-side ToTensor + Normalize of a uint8 image batch [images, channels <= 4, pixels] (pixels % 16 == 0; 16-byte aligned pointers).
extern "C" int vsx_image_normalize_u8(const void* in_u8, float* out, int images, int channels, long pixels, const float* mean, const float* stdv,
                                      void* stream) {
  VSX_REQUIRE(images >= 0 && channels >= 1 && channels <= 4 && pixels > 0 && pixels % 16 == 0, "vsx_image_normalize_u8: channels 1..4, pixels %% 16 == 0");
  VSX_REQUIRE(mean != nullptr && stdv != nullptr, "vsx_image_normalize_u8: mean / std are host arrays of `channels` floats");
  VSX_REQUIRE(((reinterpret_cast<uintptr_t>(in_u8) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "vsx_image_normalize_u8: 16-byte aligned buffers");
  if (images == 0) return VSX_OK;
  float m[4] = {0, 0, 0, 0}, s[4] = {1, 1, 1, 1};
  for (int c = 0; c < channels; ++c) m[c] = mean[c], s[c] = 1.0f / stdv[c];
  const long planes = (long)images * channels, plane16 = pixels / 16;
  launch_pdl(image_normalize_u8_kernel, dim3(ew_grid(planes * plane16)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
             static_cast<const uint8_t*>(in_u8), out, plane16, channels, planes, m[0], m[1], m[2], m[3], s[0], s[1], s[2], s[3]);
  return check_launch("vsx_image_normalize_u8");
}
