// 3x3 convolution of the 24-channel conv stem (nets/patch_conv.py:53-54), TMA + warp-specialised version of csrc/conv3x3.cu -- same
// arithmetic, same operands, same fused statistics; forward and data gradient.
//
// Why: the per-tile phases of conv3x3.cu (global loads into registers -> activate -> shared memory -> MMA -> 4-byte global stores)
// ran in lock step behind two CTA barriers, 16 warps per SM: ~300 us per launch against an HBM floor of 45-90 us.  Here every global
// access is a bulk tensor copy and the three jobs overlap:
//   producer warp : one elected thread walks the CTA's tiles and issues the TMA loads of a 4-stage ring -- the 18x18-pixel halo tile
//                   (4-D map [B][H][W][24], negative / overflowing coordinates are zero-filled by the hardware = the conv padding), the
//                   16x16 tile of the residual gradient (`add`) and of the producing layer's pre-activation (`y_prev`, STATS_BWD)
//   8 activation warps : BatchNorm-apply + ReLU of the PRODUCING layer in place in shared memory, once per halo pixel (border pixels
//                   outside the image stay zero: padding applies to the activated map)
//   8 MMA warps   : two tile rows each: mma.sync m16n8k16 over the packed reduction 9 taps x 24 channels (14 k-steps, 3 n-tiles),
//                   epilogue out of shared memory (add / y_prev), statistics in registers, bf16 output rows staged per warp and
//                   written with one TMA store per warp and tile
// Shared-memory pixel pitch is 48 bytes (24 channels, unpadded): 8 consecutive pixels fall into 8 distinct 16-byte bank groups, so
// ldmatrix and the 4-byte fragment accesses are conflict free without padding or swizzle.
#include "common.cuh"

namespace vsx {
namespace {

constexpr int T2 = 16, HL = T2 + 2;                  // tile / halo side (pixels)
constexpr int CH = 24, PIXB = CH * 2;                // channels, bytes per pixel
constexpr int WCP = 32, WPITCH = 40;                 // weight rows: padded channels / bf16 pitch (same operand layout as conv3x3.cu)
constexpr int HALO_BYTES = HL * HL * PIXB;           // 15552
constexpr int HALO_SLOT = (HALO_BYTES + 127) / 128 * 128;
constexpr int TILE_BYTES = T2 * T2 * PIXB;           // 12288
constexpr int NSTAGE = 4;
constexpr int ACT_WARPS = 3, MMA_WARPS = 8;          // 12 warps = 384 threads: up to 168 registers per thread
constexpr int THREADS = (ACT_WARPS + MMA_WARPS + 1) * 32;
constexpr int STG_BYTES = 2 * T2 * PIXB;             // per MMA warp: two output rows

struct ConvMaps {
  CUtensorMap in, add, yprev, out;
};

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float2 bf2(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w)); }

// dynamic shared memory layout (byte offsets from a 128-byte aligned base)
struct Layout {
  int stage_bytes, off_add, off_yprev;   // inside a stage
  int off_wts, off_stg, off_red, off_bn, off_bar, total;
};
__host__ __device__ inline Layout make_layout(bool has_add, bool has_yprev) {
  Layout l;
  l.off_add = HALO_SLOT;
  l.off_yprev = l.off_add + (has_add ? TILE_BYTES : 0);
  l.stage_bytes = l.off_yprev + (has_yprev ? TILE_BYTES : 0);
  l.off_wts = NSTAGE * l.stage_bytes;
  l.off_stg = l.off_wts + 9 * WCP * WPITCH * 2;
  l.off_red = l.off_stg + MMA_WARPS * STG_BYTES;
  l.off_bn = l.off_red + MMA_WARPS * 2 * WCP * 4;
  l.off_bar = l.off_bn + 4 * WCP * 4;
  l.total = l.off_bar + 3 * NSTAGE * 8;
  return l;
}

// STATS: 0 none, 1 forward BN statistics of the output, 2 backward reductions through relu(bn(y_prev)) of the output gradient
template <int STATS>
__global__ void __launch_bounds__(THREADS, 1) conv3x3_tma_kernel(const __grid_constant__ ConvMaps maps, const float* __restrict__ in_scale,
                                                                 const float* __restrict__ in_shift, const bf16* __restrict__ wt, int has_add,
                                                                 int B, int H, int W, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, double* __restrict__ sums) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const Layout L = make_layout(has_add != 0, STATS == 2);
  const bool act = in_scale != nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = base + L.off_bar;
  auto bar_full = [&](int s) { return bar0 + (uint32_t)s * 8u; };
  auto bar_act = [&](int s) { return bar0 + (uint32_t)(NSTAGE + s) * 8u; };
  auto bar_empty = [&](int s) { return bar0 + (uint32_t)(2 * NSTAGE + s) * 8u; };

  // weights: global [9][32][32] bf16 (n rows, k contiguous, zero padded) -> shared [9*32][WPITCH]; BN parameters of the STATS_BWD epilogue
  {
    bf16* wts = reinterpret_cast<bf16*>(gbase + L.off_wts);
    for (int idx = threadIdx.x; idx < 9 * WCP * (WCP / 8); idx += THREADS) {
      const int row = idx / (WCP / 8), ch = (idx % (WCP / 8)) * 8;
      *reinterpret_cast<uint4*>(wts + row * WPITCH + ch) = *reinterpret_cast<const uint4*>(wt + row * WCP + ch);
    }
    if (STATS == 2) {
      float* bn = reinterpret_cast<float*>(gbase + L.off_bn);
      for (int i = threadIdx.x; i < WCP; i += THREADS) {
        bn[i] = i < CH ? gamma[i] : 0.f, bn[WCP + i] = i < CH ? beta[i] : 0.f;
        bn[2 * WCP + i] = i < CH ? mean[i] : 0.f, bn[3 * WCP + i] = i < CH ? rstd[i] : 0.f;
      }
    }
    if (threadIdx.x == 0) {
      for (int s = 0; s < NSTAGE; ++s) {
        mbar_init(bar_full(s), 1);
        mbar_init(bar_act(s), ACT_WARPS);
        mbar_init(bar_empty(s), MMA_WARPS);
      }
      fence_mbar_init();
    }
  }
  __syncthreads();

  const int tiles_x = W / T2, tiles_y = H / T2, per_img = tiles_x * tiles_y, tiles = B * per_img;

  if (warp == ACT_WARPS + MMA_WARPS) {
    // ------------------------------------------------------------------------------------------ producer
    if (lane == 0) {
      tma_prefetch_desc(&maps.in);
      tma_prefetch_desc(&maps.out);
      const uint32_t tx = HALO_BYTES + (has_add ? TILE_BYTES : 0) + (STATS == 2 ? TILE_BYTES : 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int s = it % NSTAGE, n = it / NSTAGE;
        if (n > 0) mbar_wait(bar_empty(s), (uint32_t)(n - 1) & 1u);
        const int b = tile / per_img, rem = tile % per_img;
        const int ty0 = (rem / tiles_x) * T2, tx0 = (rem % tiles_x) * T2;
        const uint32_t st = base + (uint32_t)(s * L.stage_bytes);
        mbar_expect_tx(bar_full(s), tx);
        tma_load_4d(st, &maps.in, bar_full(s), 0, tx0 - 1, ty0 - 1, b);
        if (has_add) tma_load_4d(st + L.off_add, &maps.add, bar_full(s), 0, tx0, ty0, b);
        if (STATS == 2) tma_load_4d(st + L.off_yprev, &maps.yprev, bar_full(s), 0, tx0, ty0, b);
      }
    }
    return;
  }

  if (warp < ACT_WARPS) {
    // ------------------------------------------------------------------------------------------ activation warps
    if (!act) return;
    const int a = threadIdx.x;                       // 0..95: thread a owns the 16-byte channel group a % 3 of every 32nd pixel
    const int cg = a % 3, p0 = a / 3;
    constexpr int ACT_PIX = (ACT_WARPS * 32) / 3, ACT_ITERS = (HL * HL + ACT_PIX - 1) / ACT_PIX;
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sc[k] = in_scale[cg * 8 + k], sh[k] = in_shift[cg * 8 + k];
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int s = it % NSTAGE, n = it / NSTAGE;
      const int rem = tile % per_img;
      const int ty0 = (rem / tiles_x) * T2 - 1, tx0 = (rem % tiles_x) * T2 - 1;
      const uint32_t st = base + (uint32_t)(s * L.stage_bytes);
      mbar_wait(bar_full(s), (uint32_t)n & 1u);
      if (a < ACT_PIX * 3) {
#pragma unroll
        for (int j = 0; j < ACT_ITERS; ++j) {
          const int pix = p0 + ACT_PIX * j;
          if (pix < HL * HL) {
            const int hy = pix / HL, hx = pix - hy * HL;
            const int iy = ty0 + hy, ix = tx0 + hx;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
              const uint32_t addr = st + (uint32_t)(pix * PIXB + cg * 16);
              uint4 v = lds128(addr);
              uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float2 f = bf2(w[k]);
                f.x = fmaxf(fmaf(f.x, sc[2 * k], sh[2 * k]), 0.f);
                f.y = fmaxf(fmaf(f.y, sc[2 * k + 1], sh[2 * k + 1]), 0.f);
                w[k] = pack_bf16(f.x, f.y);
              }
              sts128(addr, make_uint4(w[0], w[1], w[2], w[3]));
            }
          }
        }
      }
      fence_proxy_async();                           // these generic-proxy writes precede the next TMA load into this stage
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_act(s));
    }
    return;
  }

  // -------------------------------------------------------------------------------------------- MMA + epilogue warps
  const int mw = warp - ACT_WARPS;                   // tile rows 2*mw, 2*mw + 1
  const int g = lane >> 2, t = lane & 3, lj = lane >> 3, li = lane & 7;
  const uint32_t wts_a = base + L.off_wts, stg = base + L.off_stg + (uint32_t)(mw * STG_BYTES);
  const float* bn = reinterpret_cast<const float*>(gbase + L.off_bn);
  float st_acc[3][2][2];
#pragma unroll
  for (int i = 0; i < 3; ++i) st_acc[i][0][0] = st_acc[i][0][1] = st_acc[i][1][0] = st_acc[i][1][1] = 0.f;
  float gm[3][2], bt[3][2], mn[3][2], rs[3][2];
  if (STATS == 2) {
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ch = nt * 8 + 2 * t + c;
        gm[nt][c] = bn[ch], bt[nt][c] = bn[WCP + ch], mn[nt][c] = bn[2 * WCP + ch], rs[nt][c] = bn[3 * WCP + ch];
      }
  }
  // the weight fragments of the first KREG k-steps never change: keep them in registers (ldmatrix traffic on the weights was 36 % of the
  // shared-memory wavefronts of a tile, and the shared-memory pipe is the busiest unit of this kernel)
  constexpr int KREG = 12;
  uint32_t bw[KREG][6];
#pragma unroll
  for (int ks = 0; ks < KREG; ++ks) {
    constexpr int NG = 27;
    const int q0 = 2 * ks, q1 = 2 * ks + 1;
    const int t0 = q0 < NG ? q0 / 3 : 0, c0 = q0 < NG ? q0 % 3 : 3;
    const int t1 = q1 < NG ? q1 / 3 : 0, c1 = q1 < NG ? q1 % 3 : 3;
    const int b0 = (t0 * WCP * WPITCH + c0 * 8), b1 = (t1 * WCP * WPITCH + c1 * 8);
    uint32_t w01[4], w23[4];
    ldsm4(w01, wts_a + (uint32_t)((li + (lj >> 1) * 8) * WPITCH + ((lj & 1) ? b1 : b0)) * 2u);
    ldsm4(w23, wts_a + (uint32_t)((16 + li + (lj >> 1) * 8) * WPITCH + ((lj & 1) ? b1 : b0)) * 2u);
    bw[ks][0] = w01[0], bw[ks][1] = w01[1], bw[ks][2] = w01[2], bw[ks][3] = w01[3], bw[ks][4] = w23[0], bw[ks][5] = w23[1];
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    const int s = it % NSTAGE, n = it / NSTAGE;
    const int b = tile / per_img, rem = tile % per_img;
    const int ty0 = (rem / tiles_x) * T2, tx0 = (rem % tiles_x) * T2;
    const uint32_t sbase = base + (uint32_t)(s * L.stage_bytes);
    mbar_wait(bar_full(s), (uint32_t)n & 1u);
    if (act) mbar_wait(bar_act(s), (uint32_t)n & 1u);
    float acc[2][3][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int i = 0; i < 3; ++i) acc[m][i][0] = acc[m][i][1] = acc[m][i][2] = acc[m][i][3] = 0.f;
    // packed reduction: 27 groups of 8 channels (9 taps x 3) + one zero group = 14 k16 steps; the two k8 halves of a step may come from
    // different taps because every ldmatrix row address is per lane.  Group 27: the A side reads the next pixel's first group (finite
    // data), the B side the zero-padded channels 24..31 of a weight row.
    const uint32_t a_lane = (uint32_t)(((2 * mw) * HL + li + (lj & 1) * 8) * PIXB);
#pragma unroll
    for (int ks = 0; ks < 14; ++ks) {
      constexpr int NG = 27;
      const int q0 = 2 * ks, q1 = 2 * ks + 1;
      const int t0 = q0 < NG ? q0 / 3 : 0, c0 = q0 < NG ? q0 % 3 : 3;
      const int t1 = q1 < NG ? q1 / 3 : 0, c1 = q1 < NG ? q1 % 3 : 3;
      const int a0 = ((t0 / 3) * HL + t0 % 3) * PIXB + c0 * 16, a1 = ((t1 / 3) * HL + t1 % 3) * PIXB + c1 * 16;      // bytes
      const int b0 = (t0 * WCP * WPITCH + c0 * 8), b1 = (t1 * WCP * WPITCH + c1 * 8);                              // elements
      uint32_t ar0[4], ar1[4], wf[6];
      const uint32_t aoff = sbase + a_lane + (uint32_t)((lj >> 1) ? a1 : a0);
      ldsm4(ar0, aoff);
      ldsm4(ar1, aoff + HL * PIXB);
      if (ks < KREG) {
#pragma unroll
        for (int i = 0; i < 6; ++i) wf[i] = bw[ks][i];
      } else {
        uint32_t w01[4], w23[4];
        ldsm4(w01, wts_a + (uint32_t)((li + (lj >> 1) * 8) * WPITCH + ((lj & 1) ? b1 : b0)) * 2u);
        ldsm4(w23, wts_a + (uint32_t)((16 + li + (lj >> 1) * 8) * WPITCH + ((lj & 1) ? b1 : b0)) * 2u);
        wf[0] = w01[0], wf[1] = w01[1], wf[2] = w01[2], wf[3] = w01[3], wf[4] = w23[0], wf[5] = w23[1];
      }
      mma16816(acc[0][0], ar0, wf[0], wf[1]);
      mma16816(acc[0][1], ar0, wf[2], wf[3]);
      mma16816(acc[0][2], ar0, wf[4], wf[5]);
      mma16816(acc[1][0], ar1, wf[0], wf[1]);
      mma16816(acc[1][1], ar1, wf[2], wf[3]);
      mma16816(acc[1][2], ar1, wf[4], wf[5]);
    }
    // the previous tile's output rows have left the staging buffer?
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    // epilogue: thread holds pixels x = g, g+8 of rows 2*mw + m, channels nt*8 + 2t + {0,1}.  For a fixed n-tile the four 8x8 blocks
    // (m, pixel half) are one ldmatrix / stmatrix x4: lane l addresses row l%8 of block l/8 = pixel (l/8 & 1)*8 + l%8 of row l/16.
    const uint32_t frag_off = (uint32_t)((((lj >> 1) * T2 + (lj & 1) * 8 + li) * CH) * 2);          // inside a two-row slab
    const uint32_t slab = (uint32_t)(2 * mw * T2 * CH * 2);
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      uint32_t av[4], yv[4], pk[4];
      if (has_add) ldsm4(av, sbase + L.off_add + slab + frag_off + nt * 16);
      if (STATS == 2) ldsm4(yv, sbase + L.off_yprev + slab + frag_off + nt * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {                  // q = m*2 + hh
        const int m = q >> 1, hh = q & 1;
        float v0 = acc[m][nt][2 * hh], v1 = acc[m][nt][2 * hh + 1];
        if (has_add) {
          const float2 a2 = bf2(av[q]);
          v0 += a2.x, v1 += a2.y;
        }
        pk[q] = pack_bf16(v0, v1);
        if (STATS != 0) {
          const float2 r2 = bf2(pk[q]);                // what the next kernel reads
          if (STATS == 1) {
            st_acc[nt][0][0] += r2.x, st_acc[nt][0][1] += r2.x * r2.x;
            st_acc[nt][1][0] += r2.y, st_acc[nt][1][1] += r2.y * r2.y;
          } else {
            const float2 y2 = bf2(yv[q]);
            const float z0 = (y2.x - mn[nt][0]) * rs[nt][0], z1 = (y2.y - mn[nt][1]) * rs[nt][1];
            const float d0 = (gm[nt][0] * z0 + bt[nt][0] > 0.f) ? r2.x : 0.f, d1 = (gm[nt][1] * z1 + bt[nt][1] > 0.f) ? r2.y : 0.f;
            st_acc[nt][0][0] += d0, st_acc[nt][0][1] += d0 * z0;
            st_acc[nt][1][0] += d1, st_acc[nt][1][1] += d1 * z1;
          }
        }
      }
      asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(stg + frag_off + nt * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                   "r"(pk[3])
                   : "memory");
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(bar_empty(s));                     // every lane of this warp is done with the stage (halo, add, y_prev)
      tma_store_4d(&maps.out, stg, 0, tx0, ty0 + 2 * mw, b);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (STATS != 0) {
    float* red = reinterpret_cast<float*>(gbase + L.off_red);      // [MMA_WARPS][2][WCP]
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float v = st_acc[nt][c][q];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) red[(mw * 2 + q) * WCP + nt * 8 + 2 * t + c] = v;
        }
    named_bar_sync(1, MMA_WARPS * 32);
    const int i = threadIdx.x - ACT_WARPS * 32;
    if (i < 2 * WCP) {
      const int q = i / WCP, ch = i % WCP;
      if (ch < CH) {
        float v = 0.f;
        for (int w = 0; w < MMA_WARPS; ++w) v += red[(w * 2 + q) * WCP + ch];
        atomicAdd(sums + q * CH + ch, (double)v);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------ weight gradient
// dW[co][tap][ci] += sum_p dy[p][co] * act(in[p + tap - 1][ci]).  Same ring / producer / activation warps; 12 MMA warps = 3 kernel rows (ky) x 4
// row quarters of the tile: a warp accumulates its three taps (kx = 0..2) x 2 m-tiles (co) x 3 n-tiles (ci) over the whole kernel, so the
// dy fragments are loaded once per three taps.  Partial sums of the four row quarters are added in shared memory at the end: one atomic
// per weight-gradient element per CTA.
constexpr int WG_MMA_WARPS = 12, WG_ACT_WARPS = 3;
constexpr int WG_THREADS = (WG_ACT_WARPS + WG_MMA_WARPS + 1) * 32;      // 512
constexpr int WG_STAGE = HALO_SLOT + TILE_BYTES;

__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2t(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

struct WgradMaps {
  CUtensorMap in, dy;
};

__global__ void __launch_bounds__(WG_THREADS, 1) conv3x3_wgrad_tma_kernel(const __grid_constant__ WgradMaps maps, const float* __restrict__ in_scale,
                                                                          const float* __restrict__ in_shift, float* __restrict__ dw, int B, int H,
                                                                          int W) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const bool act = in_scale != nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = base + NSTAGE * WG_STAGE;
  auto bar_full = [&](int s) { return bar0 + (uint32_t)s * 8u; };
  auto bar_act = [&](int s) { return bar0 + (uint32_t)(NSTAGE + s) * 8u; };
  auto bar_empty = [&](int s) { return bar0 + (uint32_t)(2 * NSTAGE + s) * 8u; };
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_act(s), WG_ACT_WARPS);
      mbar_init(bar_empty(s), WG_MMA_WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int tiles_x = W / T2, tiles_y = H / T2, per_img = tiles_x * tiles_y, tiles = B * per_img;

  if (warp == WG_ACT_WARPS + WG_MMA_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.in);
      tma_prefetch_desc(&maps.dy);
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int s = it % NSTAGE, n = it / NSTAGE;
        if (n > 0) mbar_wait(bar_empty(s), (uint32_t)(n - 1) & 1u);
        const int b = tile / per_img, rem = tile % per_img;
        const int ty0 = (rem / tiles_x) * T2, tx0 = (rem % tiles_x) * T2;
        const uint32_t st = base + (uint32_t)(s * WG_STAGE);
        mbar_expect_tx(bar_full(s), HALO_BYTES + TILE_BYTES);
        tma_load_4d(st, &maps.in, bar_full(s), 0, tx0 - 1, ty0 - 1, b);
        tma_load_4d(st + HALO_SLOT, &maps.dy, bar_full(s), 0, tx0, ty0, b);
      }
    }
    return;
  }

  if (warp < WG_ACT_WARPS) {
    if (!act) return;
    const int a = threadIdx.x;                       // 0..95: thread a owns the 16-byte channel group a % 3 of every 32nd pixel
    const int cg = a % 3, p0 = a / 3;
    constexpr int ACT_PIX = (WG_ACT_WARPS * 32) / 3, ACT_ITERS = (HL * HL + ACT_PIX - 1) / ACT_PIX;
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sc[k] = in_scale[cg * 8 + k], sh[k] = in_shift[cg * 8 + k];
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int s = it % NSTAGE, n = it / NSTAGE;
      const int rem = tile % per_img;
      const int ty0 = (rem / tiles_x) * T2 - 1, tx0 = (rem % tiles_x) * T2 - 1;
      const uint32_t st = base + (uint32_t)(s * WG_STAGE);
      mbar_wait(bar_full(s), (uint32_t)n & 1u);
#pragma unroll
      for (int j = 0; j < ACT_ITERS; ++j) {
        const int pix = p0 + ACT_PIX * j;
        if (pix < HL * HL) {
          const int hy = pix / HL, hx = pix - hy * HL;
          const int iy = ty0 + hy, ix = tx0 + hx;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
            const uint32_t addr = st + (uint32_t)(pix * PIXB + cg * 16);
            uint4 v = lds128(addr);
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float2 f = bf2(w[k]);
              f.x = fmaxf(fmaf(f.x, sc[2 * k], sh[2 * k]), 0.f);
              f.y = fmaxf(fmaf(f.y, sc[2 * k + 1], sh[2 * k + 1]), 0.f);
              w[k] = pack_bf16(f.x, f.y);
            }
            sts128(addr, make_uint4(w[0], w[1], w[2], w[3]));
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_act(s));
    }
    return;
  }

  // -------------------------------------------------------------------------------------------- MMA warps
  const int mw = warp - WG_ACT_WARPS;                // 0..11
  const int ky = mw % 3, quarter = mw / 3;           // tile rows 4*quarter .. 4*quarter + 3
  const int g = lane >> 2, t = lane & 3, lj = lane >> 3, li = lane & 7;
  float acc[3][2][3][4];                             // [kx][m-tile: 16 output channels][n-tile: 8 input channels]
#pragma unroll
  for (int kx = 0; kx < 3; ++kx)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) acc[kx][m][nt][0] = acc[kx][m][nt][1] = acc[kx][m][nt][2] = acc[kx][m][nt][3] = 0.f;
  // A[m = co][k = pixel] from dy[pixel][co] (transposed): blocks (k lo, co lo), (k lo, co hi), (k hi, co lo), (k hi, co hi)
  const uint32_t a_lane = (uint32_t)((((lj >> 1) * 8 + li) * CH) * 2 + (lj & 1) * 16);
  // B[k = pixel][n = ci] from the halo row shifted by the tap (transposed): (k lo, n0), (k hi, n0), (k lo, n1), (k hi, n1); x2: n2 only
  const uint32_t b_lane = (uint32_t)((((lj & 1) * 8 + li) * CH) * 2 + (lj >> 1) * 16);
  const uint32_t b2_lane = (uint32_t)(((((lane >> 3) & 1) * 8 + li) * CH) * 2 + 32);
  int it = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    const int s = it % NSTAGE, n = it / NSTAGE;
    const uint32_t halo = base + (uint32_t)(s * WG_STAGE), dys = halo + HALO_SLOT;
    mbar_wait(bar_full(s), (uint32_t)n & 1u);
    if (act) mbar_wait(bar_act(s), (uint32_t)n & 1u);
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {                 // k-step = the 16 pixels of tile row r
      const int r = quarter * 4 + rr;
      uint32_t a0[4], a1[4];
      ldsm4t(a0, dys + (uint32_t)(r * T2 * PIXB) + a_lane);
      ldsm4t(a1, dys + (uint32_t)(r * T2 * PIXB) + a_lane + 32);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint32_t hp = halo + (uint32_t)(((r + ky) * HL + kx) * PIXB);
        uint32_t b01[4], b2[2];
        ldsm4t(b01, hp + b_lane);
        ldsm2t(b2, hp + b2_lane);
        mma16816(acc[kx][0][0], a0, b01[0], b01[1]);
        mma16816(acc[kx][0][1], a0, b01[2], b01[3]);
        mma16816(acc[kx][0][2], a0, b2[0], b2[1]);
        mma16816(acc[kx][1][0], a1, b01[0], b01[1]);
        mma16816(acc[kx][1][1], a1, b01[2], b01[3]);
        mma16816(acc[kx][1][2], a1, b2[0], b2[1]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty(s));
  }
  // add the four row quarters in shared memory (the ring is idle now), then one atomic per element per CTA
  named_bar_sync(1, WG_MMA_WARPS * 32);
  float* red = reinterpret_cast<float*>(gbase);      // [quarter 1..3][ky][72][32 lanes]
  if (quarter > 0) {
    float* dst = red + (((quarter - 1) * 3 + ky) * 72) * 32 + lane;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) dst[(((kx * 2 + m) * 3 + nt) * 4 + e) * 32] = acc[kx][m][nt][e];
  }
  named_bar_sync(1, WG_MMA_WARPS * 32);
  if (quarter == 0) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = (((kx * 2 + m) * 3 + nt) * 4 + e) * 32 + lane;
            const float v = acc[kx][m][nt][e] + red[(0 * 3 + ky) * 72 * 32 + i] + red[(1 * 3 + ky) * 72 * 32 + i] + red[(2 * 3 + ky) * 72 * 32 + i];
            const int co = m * 16 + g + (e >> 1) * 8, ci = nt * 8 + 2 * t + (e & 1);
            if (co < CH) atomicAdd(dw + ((long)co * 9 + ky * 3 + kx) * CH + ci, v);
          }
  }
}

}  // namespace

bool conv3x3_tma_supported(int H, int W, int C) { return C == CH && H % T2 == 0 && W % T2 == 0; }

int conv3x3_tma_launch(const void* in, const float* in_scale, const float* in_shift, const void* wt, const void* add, void* out, int B, int H, int W,
                       int stats_mode, const void* y_prev, const float* gamma, const float* beta, const float* mean, const float* rstd, double* sums,
                       cudaStream_t st) {
  ConvMaps maps;
  int rc;
  if ((rc = make_tmap_nhwc(&maps.in, in, CH, W, H, B, HL, HL))) return rc;
  if ((rc = make_tmap_nhwc(&maps.out, out, CH, W, H, B, T2, 2))) return rc;
  if ((rc = make_tmap_nhwc(&maps.add, add ? add : in, CH, W, H, B, T2, T2))) return rc;
  if ((rc = make_tmap_nhwc(&maps.yprev, y_prev ? y_prev : in, CH, W, H, B, T2, T2))) return rc;
  const Layout L = make_layout(add != nullptr, stats_mode == 2);
  const int smem = L.total + 128;
  const int tiles = B * (H / T2) * (W / T2);
  const int grid = std::min(tiles, num_sms());
#define VSX_CONV2_ARGS maps, in_scale, in_shift, (const bf16*)wt, add != nullptr ? 1 : 0, B, H, W, gamma, beta, mean, rstd, sums
#define VSX_CONV2_LAUNCH(S)                                                                                      \
  do {                                                                                                           \
    static bool attr = false;                                                                                    \
    if (!attr) {                                                                                                 \
      if (cudaFuncSetAttribute(conv3x3_tma_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) { \
        set_error("vsx_conv3x3: cannot raise the dynamic shared memory limit");                                   \
        return VSX_ERR_CUDA;                                                                                     \
      }                                                                                                          \
      attr = true;                                                                                               \
    }                                                                                                            \
    launch_pdl(conv3x3_tma_kernel<S>, dim3(grid), dim3(THREADS), smem, st, VSX_CONV2_ARGS);                                           \
  } while (0)
  if (stats_mode == 0) VSX_CONV2_LAUNCH(0);
  else if (stats_mode == 1) VSX_CONV2_LAUNCH(1);
  else VSX_CONV2_LAUNCH(2);
#undef VSX_CONV2_LAUNCH
#undef VSX_CONV2_ARGS
  return check_launch("vsx_conv3x3");
}

int conv3x3_wgrad_tma_launch(const void* dy, const void* in, const float* in_scale, const float* in_shift, float* dw, int B, int H, int W,
                             cudaStream_t st) {
  WgradMaps maps;
  int rc;
  if ((rc = make_tmap_nhwc(&maps.in, in, CH, W, H, B, HL, HL))) return rc;
  if ((rc = make_tmap_nhwc(&maps.dy, dy, CH, W, H, B, T2, T2))) return rc;
  const int smem = NSTAGE * WG_STAGE + 3 * NSTAGE * 8 + 128;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(conv3x3_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
      set_error("vsx_conv3x3_wgrad: cannot raise the dynamic shared memory limit");
      return VSX_ERR_CUDA;
    }
    attr = true;
  }
  const int tiles = B * (H / T2) * (W / T2);
  launch_pdl(conv3x3_wgrad_tma_kernel, dim3(std::min(tiles, num_sms())), dim3(WG_THREADS), smem, st, maps, in_scale, in_shift, dw, B, H, W);
  return check_launch("vsx_conv3x3_wgrad");
}

}  // namespace vsx
