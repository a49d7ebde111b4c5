// First convolution of the conv stem (nets/patch_conv.py:25-30, ConvBnAct(3 -> 24, 3x3, stride 2, pad 1)) as ONE kernel per direction:
//   conv1_fwd   : fp32 NCHW image -> bf16 channels-last map y[B, H/2, W/2, 24] (+ the BatchNorm batch statistics of y)
//   conv1_wgrad : dW[co][(ky*3+kx)*3 + c] += sum_p dy[p][co] * img[p, tap, c]
// Before: im2col (image -> 64-byte bf16 rows in HBM, 205 MB) + a 128x128-tile tcgen05 GEMM whose tiles are 24 / 27 wide (TMA fills the
// full boxes, so it ran at the shared-memory fill rate: 188 us forward, 165 us weight gradient) + a separate statistics pass.  Here a
// warp builds the 27-value rows of 32 consecutive output pixels in shared memory, runs them through mma.sync (K = 27 -> 32, N = 24)
// and writes 1536 contiguous bytes: HBM traffic is the image once and the map once.
#include "common.cuh"

namespace vsx {
namespace {

constexpr int C1_OUT = 24, C1_K = 27, C1_KP = 32;
constexpr int C1_WARPS = 8;
constexpr int A_PITCH = 80;                // bytes per staged im2col row (64 used): conflict-free ldmatrix
constexpr int D_PITCH = 48;                // bytes per staged output / gradient row (24 bf16)

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// The 27 image values under output pixel `pix` (zero outside the image / beyond `total`), as 16 packed bf16 pairs (k = 27..31 zero),
// written as one 64-byte row of the warp's staging tile.
__device__ __forceinline__ void stage_row(const float* __restrict__ img, long pix, long total, int H, int W, int Ho, int Wo, uint8_t* row) {
  float v[C1_KP];
#pragma unroll
  for (int i = 0; i < C1_KP; ++i) v[i] = 0.f;
  if (pix < total) {
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho);
    const long b = pix / ((long)Wo * Ho);
    const float* src = img + b * 3 * H * W;
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
  }
  uint4* o = reinterpret_cast<uint4*>(row);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    o[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]), pack_bf16(v[8 * i + 4], v[8 * i + 5]),
                      pack_bf16(v[8 * i + 6], v[8 * i + 7]));
}

__device__ __forceinline__ float2 bf2(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w)); }

template <bool STATS>
__global__ void __launch_bounds__(C1_WARPS * 32) conv1_fwd_kernel(const float* __restrict__ img, const bf16* __restrict__ wt, long ldw,
                                                                  bf16* __restrict__ y, int B, int H, int W, int Ho, int Wo,
                                                                  double* __restrict__ sums) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) uint8_t stage[C1_WARPS][32 * A_PITCH];
  __shared__ float red[C1_WARPS][2][C1_KP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, lj = lane >> 3, li = lane & 7;
  // B fragments (weights [n = out channel][k], k contiguous, zero padded to 32): resident for the whole kernel
  uint32_t bw[2][3][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      const bf16* wr = wt + (long)(nt * 8 + g) * ldw + ks * 16 + 2 * t;
      bw[ks][nt][0] = *reinterpret_cast<const uint32_t*>(wr);
      bw[ks][nt][1] = *reinterpret_cast<const uint32_t*>(wr + 8);
    }
  float st[3][2][2];
#pragma unroll
  for (int i = 0; i < 3; ++i) st[i][0][0] = st[i][0][1] = st[i][1][0] = st[i][1][1] = 0.f;
  uint8_t* my = stage[warp];
  const uint32_t my_a = smem_u32(my);
  const long total = (long)B * Ho * Wo;
  const long groups = (total + 31) / 32;
  for (long grp = (long)blockIdx.x * C1_WARPS + warp; grp < groups; grp += (long)gridDim.x * C1_WARPS) {
    const long p0 = grp * 32;
    stage_row(img, p0 + lane, total, H, W, Ho, Wo, my + lane * A_PITCH);
    __syncwarp();
    float acc[2][3][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) acc[m][nt][0] = acc[m][nt][1] = acc[m][nt][2] = acc[m][nt][3] = 0.f;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t a[4];
        ldsm4(a, my_a + (uint32_t)((m * 16 + li + (lj & 1) * 8) * A_PITCH + ks * 32 + (lj >> 1) * 16));
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) mma16816(acc[m][nt], a, bw[ks][nt][0], bw[ks][nt][1]);
      }
    __syncwarp();                                        // all fragment loads done: the staging rows become the output rows
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      uint32_t pk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {                      // q = m*2 + pixel half
        pk[q] = pack_bf16(acc[q >> 1][nt][2 * (q & 1)], acc[q >> 1][nt][2 * (q & 1) + 1]);
        if (STATS) {
          const float2 r2 = bf2(pk[q]);                  // what the next kernel reads
          st[nt][0][0] += r2.x, st[nt][0][1] += r2.x * r2.x;
          st[nt][1][0] += r2.y, st[nt][1][1] += r2.y * r2.y;
        }
      }
      asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(my_a + (uint32_t)(((lj >> 1) * 16 + (lj & 1) * 8 + li) * D_PITCH + nt * 16)),
                   "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                   : "memory");
    }
    __syncwarp();
    // 32 pixels x 48 bytes are contiguous in y: 96 16-byte chunks, 3 per lane
    uint8_t* dst = reinterpret_cast<uint8_t*>(y + p0 * C1_OUT);
    const long valid = (total - p0 < 32 ? total - p0 : 32) * D_PITCH;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int off = (i * 32 + lane) * 16;
      if (off < valid) *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(my + off);
    }
    __syncwarp();
  }
  if (STATS) {
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float v = st[nt][c][q];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) red[warp][q][nt * 8 + 2 * t + c] = v;
        }
    __syncthreads();
    if (threadIdx.x < 2 * C1_KP) {
      const int q = threadIdx.x / C1_KP, ch = threadIdx.x % C1_KP;
      if (ch < C1_OUT) {
        float v = 0.f;
        for (int w = 0; w < C1_WARPS; ++w) v += red[w][q][ch];
        atomicAdd(sums + q * C1_OUT + ch, (double)v);
      }
    }
  }
}

__global__ void __launch_bounds__(C1_WARPS * 32) conv1_wgrad_kernel(const float* __restrict__ img, const bf16* __restrict__ dy, float* __restrict__ dw,
                                                                    long ldw, int B, int H, int W, int Ho, int Wo) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int WSTAGE = 32 * A_PITCH + 32 * D_PITCH + 64;                  // +64: the second m-tile's loads of the last row overhang
  __shared__ __align__(16) uint8_t smem[C1_WARPS * WSTAGE];                  // 33 KB; re-used for the final reduction
  static_assert(C1_WARPS * C1_OUT * 28 * 4 <= C1_WARPS * WSTAGE, "reduction buffer must fit in the staging area");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, lj = lane >> 3, li = lane & 7;
  float acc[2][4][4];                                    // [m-tile: 16 output channels][n-tile: 8 reduction columns]
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[m][nt][0] = acc[m][nt][1] = acc[m][nt][2] = acc[m][nt][3] = 0.f;
  uint8_t *ma = smem + warp * WSTAGE, *md = ma + 32 * A_PITCH;
  const uint32_t ma_a = smem_u32(ma), md_a = smem_u32(md);
  if (lane < 4) *reinterpret_cast<uint4*>(md + 32 * D_PITCH + lane * 16) = make_uint4(0u, 0u, 0u, 0u);
  const long total = (long)B * Ho * Wo;
  const long groups = (total + 31) / 32;
  for (long grp = (long)blockIdx.x * C1_WARPS + warp; grp < groups; grp += (long)gridDim.x * C1_WARPS) {
    const long p0 = grp * 32;
    stage_row(img, p0 + lane, total, H, W, Ho, Wo, ma + lane * A_PITCH);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(dy + p0 * C1_OUT);
    const long valid = (total - p0 < 32 ? total - p0 : 32) * D_PITCH;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int off = (i * 32 + lane) * 16;
      *reinterpret_cast<uint4*>(md + off) = off < valid ? *reinterpret_cast<const uint4*>(src + off) : make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {                     // k-step = 16 pixels
      uint32_t a0[4], a1[4], b01[4], b23[4];
      // A[m = co][k = pixel] from dy[pixel][co] (transposed): blocks (k lo, co lo), (k lo, co hi), (k hi, co lo), (k hi, co hi)
      const uint32_t da = md_a + (uint32_t)((ks * 16 + (lj >> 1) * 8 + li) * D_PITCH + (lj & 1) * 16);
      ldsm4t(a0, da);
      ldsm4t(a1, da + 32);                               // co 16..31: 24..31 are the next pixel's data -> rows that are never written back
      // B[k = pixel][n = reduction column] from the im2col rows (transposed): (k lo, n0), (k hi, n0), (k lo, n1), (k hi, n1)
      const uint32_t ba = ma_a + (uint32_t)((ks * 16 + (lj & 1) * 8 + li) * A_PITCH + (lj >> 1) * 16);
      ldsm4t(b01, ba);
      ldsm4t(b23, ba + 32);
      mma16816(acc[0][0], a0, b01[0], b01[1]);
      mma16816(acc[0][1], a0, b01[2], b01[3]);
      mma16816(acc[0][2], a0, b23[0], b23[1]);
      mma16816(acc[0][3], a0, b23[2], b23[3]);
      mma16816(acc[1][0], a1, b01[0], b01[1]);
      mma16816(acc[1][1], a1, b01[2], b01[3]);
      mma16816(acc[1][2], a1, b23[0], b23[1]);
      mma16816(acc[1][3], a1, b23[2], b23[3]);
    }
    __syncwarp();
  }
  // CTA reduction (in the staging area, once every warp has left the loop), then one atomic per element per CTA
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem);           // [warp][24][28]
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int co = m * 16 + g + (e >> 1) * 8, k = nt * 8 + 2 * t + (e & 1);
        if (co < C1_OUT && k < 28) red[(warp * C1_OUT + co) * 28 + k] = acc[m][nt][e];
      }
  __syncthreads();
  for (int i = threadIdx.x; i < C1_OUT * C1_K; i += C1_WARPS * 32) {
    const int co = i / C1_K, k = i - co * C1_K;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < C1_WARPS; ++w) v += red[(w * C1_OUT + co) * 28 + k];
    atomicAdd(dw + (long)co * ldw + k, v);
  }
}

}  // namespace
}  // namespace vsx

using namespace vsx;

extern "C" int vsx_conv1_fwd(const float* image, const void* weight, long ldw, void* y, int B, int H, int W, double* sums, void* stream) {
  VSX_REQUIRE(H % 2 == 0 && W % 2 == 0 && ldw >= C1_KP && ldw % 2 == 0, "vsx_conv1_fwd: need even image sides and a weight pitch >= 32 (H=%d W=%d ldw=%ld)", H, W, ldw);
  VSX_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 3) == 0, "vsx_conv1_fwd: unaligned output / weight");
  if (B <= 0) return VSX_OK;
  const int Ho = H / 2, Wo = W / 2;
  const long groups = ((long)B * Ho * Wo + 31) / 32;
  const int grid = (int)std::min<long>((groups + C1_WARPS - 1) / C1_WARPS, (long)num_sms() * 6);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (sums != nullptr) launch_pdl(conv1_fwd_kernel<true>, dim3(grid), dim3(C1_WARPS * 32), 0, st, image, (const bf16*)weight, ldw, (bf16*)y, B, H, W, Ho, Wo, sums);
  else launch_pdl(conv1_fwd_kernel<false>, dim3(grid), dim3(C1_WARPS * 32), 0, st, image, (const bf16*)weight, ldw, (bf16*)y, B, H, W, Ho, Wo, nullptr);
  return check_launch("vsx_conv1_fwd");
}

extern "C" int vsx_conv1_wgrad(const float* image, const void* dy, float* dw, long ldw, int B, int H, int W, void* stream) {
  VSX_REQUIRE(H % 2 == 0 && W % 2 == 0 && ldw >= C1_K, "vsx_conv1_wgrad: need even image sides and ldw >= 27");
  VSX_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0, "vsx_conv1_wgrad: unaligned gradient map");
  if (B <= 0) return VSX_OK;
  const int Ho = H / 2, Wo = W / 2;
  const long groups = ((long)B * Ho * Wo + 31) / 32;
  const int grid = (int)std::min<long>((groups + C1_WARPS - 1) / C1_WARPS, (long)num_sms() * 4);
  launch_pdl(conv1_wgrad_kernel, dim3(grid), dim3(C1_WARPS * 32), 0, reinterpret_cast<cudaStream_t>(stream), image, (const bf16*)dy, dw, ldw, B, H, W, Ho, Wo);
  return check_launch("vsx_conv1_wgrad");
}
