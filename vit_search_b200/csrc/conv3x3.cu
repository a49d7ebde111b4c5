// Direct 3x3 convolution (stride 1, pad 1, C_in = C_out = C <= 32) on channels-last bf16 maps -- the two 24->24 convs of
// the conv stem (nets/patch_conv.py:53-54, ConvBnAct :23-36), forward, data gradient and weight gradient, without ever
// materialising an im2col matrix.  HBM-bound by design: each map is read once and written once.
//
//   conv3x3_fwd : out[p][n] = sum_{tap,k} act(in[p + tap - 1][k]) * wt[tap][n][k]
//                 act = ReLU(BatchNorm-apply) of the PRODUCING layer fused into the halo load (scale/shift per channel),
//                 optional `add` (residual gradient), optional per-channel statistics fused into the epilogue:
//                   STATS_FWD : sum(out), sum(out^2)                          (BatchNorm batch statistics of this layer)
//                   STATS_BWD : sum(dz), sum(dz * zhat), dz = out * [bn(y) > 0]  (reductions of the next BN backward)
//                 The data gradient is the same kernel with flipped/transposed weights and no input activation.
//   conv3x3_wgrad : dW[n][tap][k] += sum_p dy[p][n] * act(in[p + tap - 1][k])
//
// Tiling: a CTA walks 8x16-pixel tiles of one image (persistent, grid = SMs x 2); the activated 10x18 halo tile and the
// weights sit in shared memory with an 80-byte pixel pitch (conflict-free ldmatrix); mma.sync m16n8k16, fp32 accumulate.
// Channels are padded to 32 inside shared memory only.
#include "common.cuh"

namespace vsx {
namespace {

constexpr int TH = 8, TW = 16;             // output tile (pixels)
constexpr int HH = TH + 2, HW = TW + 2;    // halo tile
constexpr int CP = 32;                     // padded channels
constexpr int PITCH = 40;                  // bf16 elements per pixel / weight row in shared memory (80 bytes)

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

// The halo tile of the NEXT tile is fetched into registers while the current tile is in the tensor cores (software
// pipelining), then activated (BatchNorm-apply + ReLU of the producing layer) and committed to shared memory.
// OOB pixels and channels >= C are zero (zero padding applies to the ACTIVATED map).
constexpr int HALO_ITEMS = HH * HW * (CP / 8);   // 16-byte chunks

template <int NT>
struct HaloRegs {
  static constexpr int N = (HALO_ITEMS + NT - 1) / NT;
  uint4 v[N];
  uint32_t valid;
};

template <int NT>
__device__ __forceinline__ void halo_fetch(HaloRegs<NT>& r, const bf16* __restrict__ in, int b, int H, int W, int C, int ty0, int tx0) {
  r.valid = 0;
#pragma unroll
  for (int j = 0; j < HaloRegs<NT>::N; ++j) {
    const int idx = threadIdx.x + j * NT;
    r.v[j] = make_uint4(0u, 0u, 0u, 0u);
    if (idx < HALO_ITEMS) {
      const int pix = idx / (CP / 8), ch = (idx % (CP / 8)) * 8;
      const int iy = ty0 - 1 + pix / HW, ix = tx0 - 1 + pix % HW;
      if (ch < C && iy >= 0 && iy < H && ix >= 0 && ix < W) {
        r.v[j] = *reinterpret_cast<const uint4*>(in + (((long)b * H + iy) * W + ix) * C + ch);
        r.valid |= 1u << j;
      }
    }
  }
}

template <int NT>
__device__ __forceinline__ void halo_commit(bf16* halo, const HaloRegs<NT>& r, const float* sc_s, const float* sh_s, bool act) {
#pragma unroll
  for (int j = 0; j < HaloRegs<NT>::N; ++j) {
    const int idx = threadIdx.x + j * NT;
    if (idx < HALO_ITEMS) {
      const int pix = idx / (CP / 8), ch = (idx % (CP / 8)) * 8;
      uint4 v = r.v[j];
      if (act && (r.valid >> j & 1u)) {
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w[k]));
          f.x = fmaxf(f.x * sc_s[ch + 2 * k] + sh_s[ch + 2 * k], 0.f);
          f.y = fmaxf(f.y * sc_s[ch + 2 * k + 1] + sh_s[ch + 2 * k + 1], 0.f);
          w[k] = pack_bf16(f.x, f.y);
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
      }
      *reinterpret_cast<uint4*>(halo + pix * PITCH + ch) = v;
    }
  }
}

constexpr int FWD_WARPS = 8;   // one warp per tile row

// STATS: 0 none, 1 forward BN statistics of the output, 2 backward reductions through relu(bn(y_prev)) of the output gradient
template <int STATS>
__global__ void __launch_bounds__(FWD_WARPS * 32) conv3x3_fwd_kernel(const bf16* __restrict__ in, const float* __restrict__ in_scale,
                                                                     const float* __restrict__ in_shift, const bf16* __restrict__ wt,
                                                                     const bf16* __restrict__ add, bf16* __restrict__ out, int B, int H, int W,
                                                                     int C, const bf16* __restrict__ y_prev, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd, double* __restrict__ sums) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) bf16 halo[HH * HW * PITCH];
  __shared__ __align__(16) bf16 wts[9 * CP * PITCH];
  __shared__ float sc_s[CP], sh_s[CP], g_s[CP], b_s[CP], m_s[CP], r_s[CP];
  __shared__ float red[FWD_WARPS][2][CP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bool act = in_scale != nullptr;
  for (int i = threadIdx.x; i < CP; i += blockDim.x) {
    sc_s[i] = (act && i < C) ? in_scale[i] : 0.f;
    sh_s[i] = (act && i < C) ? in_shift[i] : 0.f;
    if (STATS == 2) {
      g_s[i] = i < C ? gamma[i] : 0.f, b_s[i] = i < C ? beta[i] : 0.f, m_s[i] = i < C ? mean[i] : 0.f, r_s[i] = i < C ? rstd[i] : 0.f;
    }
  }
  // weights: global [9][CP][CP] bf16 (n rows, k contiguous, zero padded) -> shared [9*CP][PITCH]
  for (int idx = threadIdx.x; idx < 9 * CP * (CP / 8); idx += blockDim.x) {
    const int row = idx / (CP / 8), ch = (idx % (CP / 8)) * 8;
    *reinterpret_cast<uint4*>(wts + row * PITCH + ch) = *reinterpret_cast<const uint4*>(wt + row * CP + ch);
  }
  float st[4][2][2];   // [n-tile][channel of the pair][stat]
#pragma unroll
  for (int i = 0; i < 4; ++i) st[i][0][0] = st[i][0][1] = st[i][1][0] = st[i][1][1] = 0.f;
  const uint32_t halo_a = smem_u32(halo), wts_a = smem_u32(wts);
  const int tiles_x = W / TW, tiles_y = H / TH, tiles = B * tiles_x * tiles_y;
  const int lj = lane >> 3, li = lane & 7;
  HaloRegs<FWD_WARPS * 32> pre;
  if ((int)blockIdx.x < tiles) {
    const int rem0 = blockIdx.x % (tiles_x * tiles_y);
    halo_fetch(pre, in, blockIdx.x / (tiles_x * tiles_y), H, W, C, (rem0 / tiles_x) * TH, (rem0 % tiles_x) * TW);
  }
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int b = tile / (tiles_x * tiles_y), rem = tile % (tiles_x * tiles_y);
    const int ty0 = (rem / tiles_x) * TH, tx0 = (rem % tiles_x) * TW;
    __syncthreads();                       // previous tile's MMAs are done reading the halo (also orders the weight load)
    halo_commit(halo, pre, sc_s, sh_s, act);
    __syncthreads();
    {
      const int nxt = tile + gridDim.x;    // next tile's loads fly while this tile is in the tensor cores
      if (nxt < tiles) {
        const int remn = nxt % (tiles_x * tiles_y);
        halo_fetch(pre, in, nxt / (tiles_x * tiles_y), H, W, C, (remn / tiles_x) * TH, (remn % tiles_x) * TW);
      }
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    if (C <= 24) {
      // Packed reduction for the 24-channel stem: K = 9 taps x 3 channel groups of 8 = 27 groups (+1 zero group) = 14 k16 steps instead of
      // 18, and 3 output-channel tiles instead of 4: 42 MMAs per warp and tile instead of 72.  Every ldmatrix row address is per lane, so
      // the two k8 halves of a step may come from different taps; group 27 reads the zero-padded channels 24..31 of a pixel / weight row.
#pragma unroll
      for (int ks = 0; ks < 14; ++ks) {
        constexpr int NG = 27;
        const int q0 = 2 * ks, q1 = 2 * ks + 1;
        const int t0 = q0 < NG ? q0 / 3 : 0, c0 = q0 < NG ? q0 % 3 : 3;
        const int t1 = q1 < NG ? q1 / 3 : 0, c1 = q1 < NG ? q1 % 3 : 3;
        // A: matrix lj covers pixels li + (lj & 1) * 8, k group (lj >> 1)
        const int a0 = (((t0 / 3) * HW + t0 % 3) * PITCH + c0 * 8), a1 = (((t1 / 3) * HW + t1 % 3) * PITCH + c1 * 8);
        // B: matrix lj covers n rows li + (lj >> 1) * 8, k group (lj & 1)
        const int b0 = (t0 * CP * PITCH + c0 * 8), b1 = (t1 * CP * PITCH + c1 * 8);
        uint32_t a[4], w01[4], w23[4];
        ldsm4(a, halo_a + (uint32_t)((warp * HW + li + (lj & 1) * 8) * PITCH + ((lj >> 1) ? a1 : a0)) * 2u);
        ldsm4(w01, wts_a + (uint32_t)((li + (lj >> 1) * 8) * PITCH + ((lj & 1) ? b1 : b0)) * 2u);
        ldsm4(w23, wts_a + (uint32_t)((16 + li + (lj >> 1) * 8) * PITCH + ((lj & 1) ? b1 : b0)) * 2u);
        mma16816(acc[0], a, w01[0], w01[1]);
        mma16816(acc[1], a, w01[2], w01[3]);
        mma16816(acc[2], a, w23[0], w23[1]);
      }
    } else {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          uint32_t a[4], w01[4], w23[4];
          // A: 16 pixels of tile row `warp` (x = 0..15) shifted by the tap, 16 channels
          ldsm4(a, halo_a + (uint32_t)(((warp + ky) * HW + kx + li + (lj & 1) * 8) * PITCH + kc * 16 + (lj >> 1) * 8) * 2u);
          // B: weights [n][k]: n-tiles (0,1) and (2,3)
          ldsm4(w01, wts_a + (uint32_t)((tap * CP + li + (lj >> 1) * 8) * PITCH + kc * 16 + (lj & 1) * 8) * 2u);
          ldsm4(w23, wts_a + (uint32_t)((tap * CP + 16 + li + (lj >> 1) * 8) * PITCH + kc * 16 + (lj & 1) * 8) * 2u);
          mma16816(acc[0], a, w01[0], w01[1]);
          mma16816(acc[1], a, w01[2], w01[3]);
          mma16816(acc[2], a, w23[0], w23[1]);
          mma16816(acc[3], a, w23[2], w23[3]);
        }
      }
    }
    // epilogue: thread holds pixels x = g, g+8 of row `warp`, channels nt*8 + 2t + {0,1}
    const long prow = ((long)b * H + ty0 + warp) * W + tx0;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int ch = nt * 8 + 2 * t;
      if (ch < C) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const long off = (prow + g + hh * 8) * C + ch;
          float v0 = acc[nt][2 * hh], v1 = acc[nt][2 * hh + 1];
          if (add != nullptr) {
            const float2 a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(add + off));
            v0 += a2.x, v1 += a2.y;
          }
          const uint32_t packed = pack_bf16(v0, v1);
          *reinterpret_cast<uint32_t*>(out + off) = packed;
          if (STATS != 0) {
            const float2 r2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&packed));   // what the next kernel reads
            if (STATS == 1) {
              st[nt][0][0] += r2.x, st[nt][0][1] += r2.x * r2.x;
              st[nt][1][0] += r2.y, st[nt][1][1] += r2.y * r2.y;
            } else {
              const float2 y2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(y_prev + off));
              const float z0 = (y2.x - m_s[ch]) * r_s[ch], z1 = (y2.y - m_s[ch + 1]) * r_s[ch + 1];
              const float d0 = (g_s[ch] * z0 + b_s[ch] > 0.f) ? r2.x : 0.f, d1 = (g_s[ch + 1] * z1 + b_s[ch + 1] > 0.f) ? r2.y : 0.f;
              st[nt][0][0] += d0, st[nt][0][1] += d0 * z0;
              st[nt][1][0] += d1, st[nt][1][1] += d1 * z1;
            }
          }
        }
      }
    }
  }
  if (STATS != 0) {
    // reduce over the 8 pixel groups of the warp (lanes with equal t), then over warps, then one fp64 atomic per channel per CTA
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          float v = st[nt][c][s];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) red[warp][s][nt * 8 + 2 * t + c] = v;
        }
    __syncthreads();
    if (threadIdx.x < 2 * CP) {
      const int s = threadIdx.x / CP, ch = threadIdx.x % CP;
      if (ch < C) {
        float v = 0.f;
        for (int w = 0; w < FWD_WARPS; ++w) v += red[w][s][ch];
        atomicAdd(sums + s * C + ch, (double)v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
constexpr int WG_WARPS = 9;   // one warp per tap

__global__ void __launch_bounds__(WG_WARPS * 32) conv3x3_wgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ in,
                                                                      const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                                                                      float* __restrict__ dw, int B, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) bf16 halo[HH * HW * PITCH];
  __shared__ __align__(16) bf16 dys[TH * TW * PITCH];
  __shared__ float sc_s[CP], sh_s[CP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bool act = in_scale != nullptr;
  for (int i = threadIdx.x; i < CP; i += blockDim.x) {
    sc_s[i] = (act && i < C) ? in_scale[i] : 0.f;
    sh_s[i] = (act && i < C) ? in_shift[i] : 0.f;
  }
  float acc[2][4][4];   // [m-tile: 16 output channels][n-tile: 8 input channels]
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
  const uint32_t halo_a = smem_u32(halo), dys_a = smem_u32(dys);
  const int tiles_x = W / TW, tiles_y = H / TH, tiles = B * tiles_x * tiles_y;
  const int lj = lane >> 3, li = lane & 7;
  const int ky = warp / 3, kx = warp % 3;
  constexpr int NT = WG_WARPS * 32, DY_ITEMS = TH * TW * (CP / 8), DY_N = (DY_ITEMS + NT - 1) / NT;
  HaloRegs<NT> pre;
  uint4 dpre[DY_N];
  auto fetch = [&](int tile) {
    const int b = tile / (tiles_x * tiles_y), rem = tile % (tiles_x * tiles_y);
    const int ty0 = (rem / tiles_x) * TH, tx0 = (rem % tiles_x) * TW;
    halo_fetch(pre, in, b, H, W, C, ty0, tx0);
#pragma unroll
    for (int j = 0; j < DY_N; ++j) {
      const int idx = threadIdx.x + j * NT;
      dpre[j] = make_uint4(0u, 0u, 0u, 0u);
      if (idx < DY_ITEMS) {
        const int pix = idx / (CP / 8), ch = (idx % (CP / 8)) * 8;
        if (ch < C) dpre[j] = *reinterpret_cast<const uint4*>(dy + (((long)b * H + ty0 + pix / TW) * W + tx0 + pix % TW) * C + ch);
      }
    }
  };
  if ((int)blockIdx.x < tiles) fetch(blockIdx.x);
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    __syncthreads();
    halo_commit(halo, pre, sc_s, sh_s, act);
#pragma unroll
    for (int j = 0; j < DY_N; ++j) {
      const int idx = threadIdx.x + j * NT;
      if (idx < DY_ITEMS) *reinterpret_cast<uint4*>(dys + (idx / (CP / 8)) * PITCH + (idx % (CP / 8)) * 8) = dpre[j];
    }
    __syncthreads();
    if (tile + (int)gridDim.x < tiles) fetch(tile + gridDim.x);
#pragma unroll
    for (int r = 0; r < TH; ++r) {   // k-step = the 16 pixels of tile row r
      uint32_t a0[4], a1[4], b01[4], b23[4];
      // A[m = co][k = pixel] from dys[pixel][co] (transposed): blocks (k0,co0), (k0,co0+8), (k0+8,co0), (k0+8,co0+8)
      ldsm4t(a0, dys_a + (uint32_t)((r * TW + (lj >> 1) * 8 + li) * PITCH + (lj & 1) * 8) * 2u);
      ldsm4t(a1, dys_a + (uint32_t)((r * TW + (lj >> 1) * 8 + li) * PITCH + 16 + (lj & 1) * 8) * 2u);
      // B[k = pixel][n = ci] from the halo row shifted by this warp's tap (transposed load), n-tiles (0,1) and (2,3)
      const int hp = (r + ky) * HW + kx;
      ldsm4t(b01, halo_a + (uint32_t)((hp + (lj & 1) * 8 + li) * PITCH + (lj >> 1) * 8) * 2u);
      ldsm4t(b23, halo_a + (uint32_t)((hp + (lj & 1) * 8 + li) * PITCH + 16 + (lj >> 1) * 8) * 2u);
      mma16816(acc[0][0], a0, b01[0], b01[1]);
      mma16816(acc[0][1], a0, b01[2], b01[3]);
      mma16816(acc[0][2], a0, b23[0], b23[1]);
      mma16816(acc[1][0], a1, b01[0], b01[1]);
      mma16816(acc[1][1], a1, b01[2], b01[3]);
      mma16816(acc[1][2], a1, b23[0], b23[1]);
      if (C > 24) {                                  // input channels 24..31 exist only for C = 32
        mma16816(acc[0][3], a0, b23[2], b23[3]);
        mma16816(acc[1][3], a1, b23[2], b23[3]);
      }
    }
  }
  // dw layout: [co][tap][ci] fp32 (the GEMM-side 'ohwi' layout), accumulated with atomics (one per element per CTA)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int co = mt * 16 + g + (e >> 1) * 8, ci = nt * 8 + 2 * t + (e & 1);
        if (co < C && ci < C) atomicAdd(dw + ((long)co * 9 + warp) * C + ci, acc[mt][nt][e]);
      }
}

}  // namespace
}  // namespace vsx

namespace vsx {
// csrc/conv3x3_tma.cu: TMA + warp-specialised kernel for the 24-channel stem
bool conv3x3_tma_supported(int H, int W, int C);
int conv3x3_tma_launch(const void* in, const float* in_scale, const float* in_shift, const void* wt, const void* add, void* out, int B, int H, int W,
                       int stats_mode, const void* y_prev, const float* gamma, const float* beta, const float* mean, const float* rstd, double* sums,
                       cudaStream_t st);
int conv3x3_wgrad_tma_launch(const void* dy, const void* in, const float* in_scale, const float* in_shift, float* dw, int B, int H, int W,
                             cudaStream_t st);
static int g_conv_impl = 0;   // 0 auto, 1 legacy direct kernel, 2 TMA kernel (error if unsupported)
}  // namespace vsx

using namespace vsx;

extern "C" int vsx_conv3x3_force_impl(int impl) {
  VSX_REQUIRE(impl >= 0 && impl <= 2, "vsx_conv3x3_force_impl: 0 auto, 1 legacy, 2 tma (got %d)", impl);
  g_conv_impl = impl;
  return VSX_OK;
}

extern "C" int vsx_conv3x3(const void* in, const float* in_scale, const float* in_shift, const void* wt, const void* add, void* out, int B,
                           int H, int W, int C, int stats_mode, const void* y_prev, const float* gamma, const float* beta, const float* mean,
                           const float* rstd, double* sums, void* stream) {
  VSX_REQUIRE(C % 8 == 0 && C <= CP && H % TH == 0 && W % TW == 0, "vsx_conv3x3: needs C %% 8 == 0, C <= 32, H %% 8 == 0, W %% 16 == 0 (C=%d H=%d W=%d)", C, H, W);
  VSX_REQUIRE(stats_mode >= 0 && stats_mode <= 2 && (stats_mode == 0 || sums != nullptr), "vsx_conv3x3: bad stats_mode / sums");
  VSX_REQUIRE(stats_mode != 2 || (y_prev && gamma && beta && mean && rstd), "vsx_conv3x3: stats_mode 2 needs y_prev and the BN parameters");
  if (B <= 0) return VSX_OK;
  const int tiles = B * (H / TH) * (W / TW);
  const int grid = std::min(tiles, num_sms() * 2);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VSX_REQUIRE(g_conv_impl != 2 || conv3x3_tma_supported(H, W, C), "vsx_conv3x3: the TMA kernel needs C == 24, H %% 16 == 0, W %% 16 == 0");
  if (g_conv_impl != 1 && conv3x3_tma_supported(H, W, C))
    return conv3x3_tma_launch(in, in_scale, in_shift, wt, add, out, B, H, W, stats_mode, y_prev, gamma, beta, mean, rstd, sums, st);
#define VSX_CONV_ARGS (const bf16*)in, in_scale, in_shift, (const bf16*)wt, (const bf16*)add, (bf16*)out, B, H, W, C, (const bf16*)y_prev, gamma, beta, mean, rstd, sums
  if (stats_mode == 0) launch_pdl(conv3x3_fwd_kernel<0>, dim3(grid), dim3(FWD_WARPS * 32), 0, st, VSX_CONV_ARGS);
  else if (stats_mode == 1) launch_pdl(conv3x3_fwd_kernel<1>, dim3(grid), dim3(FWD_WARPS * 32), 0, st, VSX_CONV_ARGS);
  else launch_pdl(conv3x3_fwd_kernel<2>, dim3(grid), dim3(FWD_WARPS * 32), 0, st, VSX_CONV_ARGS);
#undef VSX_CONV_ARGS
  return check_launch("vsx_conv3x3");
}

extern "C" int vsx_conv3x3_wgrad(const void* dy, const void* in, const float* in_scale, const float* in_shift, float* dw, int B, int H, int W,
                                 int C, void* stream) {
  VSX_REQUIRE(C % 8 == 0 && C <= CP && H % TH == 0 && W % TW == 0, "vsx_conv3x3_wgrad: needs C %% 8 == 0, C <= 32, H %% 8 == 0, W %% 16 == 0");
  if (B <= 0) return VSX_OK;
  VSX_REQUIRE(g_conv_impl != 2 || conv3x3_tma_supported(H, W, C), "vsx_conv3x3_wgrad: the TMA kernel needs C == 24, H %% 16 == 0, W %% 16 == 0");
  if (g_conv_impl != 1 && conv3x3_tma_supported(H, W, C))
    return conv3x3_wgrad_tma_launch(dy, in, in_scale, in_shift, dw, B, H, W, reinterpret_cast<cudaStream_t>(stream));
  const int tiles = B * (H / TH) * (W / TW);
  const int grid = std::min(tiles, num_sms() * 2);
  launch_pdl(conv3x3_wgrad_kernel, dim3(grid), dim3(WG_WARPS * 32), 0, reinterpret_cast<cudaStream_t>(stream), (const bf16*)dy, (const bf16*)in, in_scale, in_shift, dw, B,
                                                                                          H, W, C);
  return check_launch("vsx_conv3x3_wgrad");
}
