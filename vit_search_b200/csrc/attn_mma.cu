// Tensor-core attention for the ViT-Res shapes (N = 257 / 65 / 17 tokens, head_dim 32 / 48 / 64), bf16 in / fp32 softmax.
//
// One CTA per (sample, head).  Forward: K and V of the head in shared memory (78 KB, two CTAs per SM), Q fragments straight
// from global memory.  Backward: Q, K, V, dO in shared memory (<= 158 KB, one 12-warp CTA per SM);
// every warp owns 16-row blocks and streams over the other dimension in 32/64-column chunks with mma.sync m16n8k16
// (ldmatrix operand loads, online softmax in the forward, recomputation from the saved log-sum-exp in the backward).
// Nothing of size N x N is written to HBM -- the reference materialises [B,H,N,N] scores three times
// (nets/supernet_blocks.py:105-109).  Heads >= heads_keep are never computed; their slices are zero-filled.
//
//   forward : S = Q K^T * scale, P = softmax(S), O = P V, lse = logsumexp(S)
//   backward: phase A (warp = 16 keys)   S^T = K Q^T, P^T, dV += P^T dO, dP^T = V dO^T, dS^T = P^T o (dP^T - delta), dK += dS^T Q
//             phase B (warp = 16 queries) S = Q K^T, P, dP = dO V^T, dS = P o (dP - delta), dQ += dS K
// Each phase keeps its accumulators in registers, so there are no atomics and no cross-warp reductions.
// TODO(next round): move the two contractions per chunk onto tcgen05 with S/P in tensor memory.
#include "common.cuh"

namespace vsx {
namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int STAT_PAD = 32;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Shared-memory tile [rows][D] bf16 with row pitch D+8 elements (conflict-free ldmatrix).
template <int D> struct Tile {
  static constexpr int PITCH = D + 8;
  uint32_t base;   // shared address
  __device__ __forceinline__ uint32_t addr(int row, int col) const { return base + (uint32_t)(row * PITCH + col) * 2u; }
  // A operand (16 rows x 16 k) at (row0, k0): matrices (rows 0-7,k 0-7), (rows 8-15,k 0-7), (rows 0-7,k 8-15), (rows 8-15,k 8-15)
  __device__ __forceinline__ void load_a(uint32_t (&r)[4], int row0, int k0, int lane) const {
    const int j = lane >> 3, i = lane & 7;
    ldsm_x4(r, addr(row0 + i + (j & 1) * 8, k0 + (j >> 1) * 8));
  }
  // B operand from rows = n, contiguous = k (e.g. K for Q K^T): two n-tiles (n0..n0+15) x 16 k:
  // r[0],r[1] = (b0,b1) of n-tile 0, r[2],r[3] = (b0,b1) of n-tile 1
  __device__ __forceinline__ void load_b_nk(uint32_t (&r)[4], int n0, int k0, int lane) const {
    const int j = lane >> 3, i = lane & 7;
    ldsm_x4(r, addr(n0 + i + (j >> 1) * 8, k0 + (j & 1) * 8));
  }
  // B operand from rows = k, contiguous = n (e.g. V for P V): 16 k (k0..) x two n-tiles (n0..n0+15), transposed load
  __device__ __forceinline__ void load_b_kn(uint32_t (&r)[4], int k0, int n0, int lane) const {
    const int j = lane >> 3, i = lane & 7;
    ldsm_x4_t(r, addr(k0 + i + (j & 1) * 8, n0 + (j >> 1) * 8));
  }
};

// Cooperative copy of a [N x D] head slice (row pitch ld elements in global) into a tile; rows N..rows_pad are zeroed.
template <int D>
__device__ __forceinline__ void stage_tile(uint8_t* smem, const Tile<D>& t, const bf16* __restrict__ src, long ld, int N, int rows_pad,
                                           uint32_t smem_base) {
  constexpr int CH = D / 8;   // 16-byte chunks per row
  for (int idx = threadIdx.x; idx < rows_pad * CH; idx += blockDim.x) {
    const int r = idx / CH, c = (idx % CH) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < N) v = *reinterpret_cast<const uint4*>(src + (long)r * ld + c);
    *reinterpret_cast<uint4*>(smem + (t.addr(r, c) - smem_base)) = v;
  }
}

// A-operand fragments (16 rows x D) of the rows [row0, row0+16) straight from global memory (rows >= N read as zero):
// the row-owner operands are touched once per row block, so they do not need a shared-memory copy.
template <int D>
__device__ __forceinline__ void load_a_global(uint32_t (&f)[D / 16][4], const bf16* __restrict__ src, long ld, int row0, int N, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const bool v0 = row0 + g < N, v1 = row0 + g + 8 < N;
  const bf16* p0 = src + (long)(row0 + g) * ld + 2 * t;
  const bf16* p1 = p0 + 8 * ld;
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    f[ks][0] = v0 ? *reinterpret_cast<const uint32_t*>(p0 + ks * 16) : 0u;
    f[ks][1] = v1 ? *reinterpret_cast<const uint32_t*>(p1 + ks * 16) : 0u;
    f[ks][2] = v0 ? *reinterpret_cast<const uint32_t*>(p0 + ks * 16 + 8) : 0u;
    f[ks][3] = v1 ? *reinterpret_cast<const uint32_t*>(p1 + ks * 16 + 8) : 0u;
  }
}

template <typename T>
__device__ __forceinline__ void zero_slice(T* dst, long ld, int N, int D) {
  for (int idx = threadIdx.x; idx < N * (D / 8); idx += blockDim.x) {
    const int j = idx / (D / 8), d8 = (idx % (D / 8)) * 8;
    *reinterpret_cast<uint4*>(dst + (long)j * ld + d8) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <int D>
__global__ void __launch_bounds__(256, 2) attn_fwd_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, float* __restrict__ lse, int N, int H, int Hk,
                                    float scale) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  const int h = blockIdx.x, b = blockIdx.y;
  const long ldq = 3L * H * D, ldo = (long)H * D;
  bf16* ob = o + (long)b * N * ldo + h * D;
  if (h >= Hk) {
    zero_slice(ob, ldo, N, D);
    return;
  }
  const int Np = (N + 15) / 16 * 16;
  const uint32_t sbase = smem_u32(smem);
  constexpr int TB = (D + 8) * 2;   // bytes per tile row
  Tile<D> Ks{sbase}, Vs{sbase + (uint32_t)Np * TB};
  const bf16* base = qkv + (long)b * N * ldq + h * D;
  stage_tile<D>(smem, Ks, base + (long)H * D, ldq, N, Np, sbase);
  stage_tile<D>(smem, Vs, base + 2L * H * D, ldq, N, Np, sbase);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float c = scale * LOG2E;
  const int ntiles = Np / 8;
  for (int rb = warp; rb * 16 < N; rb += nwarps) {
    const int row0 = rb * 16;
    uint32_t qf[D / 16][4];
    load_a_global<D>(qf, base, ldq, row0, N, lane);
    float oacc[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int nt0 = 0; nt0 < ntiles; nt0 += 4) {     // 32 keys per step (ntiles is even)
      float s[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        if (nt0 + 2 * p < ntiles) {
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t kb[4];
            Ks.load_b_nk(kb, (nt0 + 2 * p) * 8, ks * 16, lane);
            mma_bf16(s[2 * p], qf[ks], kb[0], kb[1]);
            mma_bf16(s[2 * p + 1], qf[ks], kb[2], kb[3]);
          }
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = (nt0 + i) * 8 + 2 * t;
        s[i][0] = col < N ? s[i][0] * c : -INFINITY;
        s[i][1] = col + 1 < N ? s[i][1] * c : -INFINITY;
        s[i][2] = col < N ? s[i][2] * c : -INFINITY;
        s[i][3] = col + 1 < N ? s[i][3] * c : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
        mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // every chunk starts below N, so mn is finite
      const float cr0 = exp2f(m0 - mn0), cr1 = exp2f(m1 - mn1);
      m0 = mn0, m1 = mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[i][0] = exp2f(s[i][0] - mn0), s[i][1] = exp2f(s[i][1] - mn0);
        s[i][2] = exp2f(s[i][2] - mn1), s[i][3] = exp2f(s[i][3] - mn1);
        rs0 += s[i][0] + s[i][1];
        rs1 += s[i][2] + s[i][3];
      }
      rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1);
      rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
      rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1);
      rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
      l0 = l0 * cr0 + rs0, l1 = l1 * cr1 + rs1;
#pragma unroll
      for (int i = 0; i < D / 8; ++i) oacc[i][0] *= cr0, oacc[i][1] *= cr0, oacc[i][2] *= cr1, oacc[i][3] *= cr1;
#pragma unroll
      for (int j = 0; j < 2; ++j) {   // k-steps of 16 keys
        if (nt0 + 2 * j < ntiles) {
          uint32_t pa[4] = {pack_bf16(s[2 * j][0], s[2 * j][1]), pack_bf16(s[2 * j][2], s[2 * j][3]),
                            pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]), pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3])};
#pragma unroll
          for (int dp = 0; dp < D / 16; ++dp) {
            uint32_t vb[4];
            Vs.load_b_kn(vb, (nt0 + 2 * j) * 8, dp * 16, lane);
            mma_bf16(oacc[2 * dp], pa, vb[0], vb[1]);
            mma_bf16(oacc[2 * dp + 1], pa, vb[2], vb[3]);
          }
        }
      }
    }
    const float il0 = 1.f / l0, il1 = 1.f / l1;
    const int r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      const int col = i * 8 + 2 * t;
      if (r0 < N) *reinterpret_cast<uint32_t*>(ob + (long)r0 * ldo + col) = pack_bf16(oacc[i][0] * il0, oacc[i][1] * il0);
      if (r1 < N) *reinterpret_cast<uint32_t*>(ob + (long)r1 * ldo + col) = pack_bf16(oacc[i][2] * il1, oacc[i][3] * il1);
    }
    if (t == 0) {
      float* lp = lse + ((long)b * H + h) * N;
      if (r0 < N) lp[r0] = (m0 + log2f(l0)) * LN2;
      if (r1 < N) lp[r1] = (m1 + log2f(l1)) * LN2;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// One 16-row block of the "owner" operand against all columns of the other, producing two accumulators:
//   (xf, yf are the owner rows' A-operand fragments, loaded from global memory by the caller)
//   TRANSPOSED = true  (phase A): owner rows = keys.    X = K, Y = V (A operands);  cols = queries: Bs = Q (for S^T), Bd = dO (for dP^T)
//        acc1 = dV += P^T dO, acc2 = dK += dS^T Q;     per-COLUMN statistics lse2[q], delta[q]
//   TRANSPOSED = false (phase B): owner rows = queries. X = Q, Y = dO;  cols = keys: Bs = K, Bd = V
//        acc2 = dQ += dS K (acc1 unused);               per-ROW statistics
template <int D, bool TRANSPOSED>
__device__ __forceinline__ void bwd_block(const uint32_t (&xf)[D / 16][4], const uint32_t (&yf)[D / 16][4], const Tile<D>& Bs, const Tile<D>& Bd,
                                          const float* __restrict__ lse2_s, const float* __restrict__ delta_s, int row0, int ntiles,
                                          float c, int lane, float (&acc1)[D / 8][4], float (&acc2)[D / 8][4]) {
  const int g = lane >> 2, t = lane & 3;
  float rl0 = 0.f, rl1 = 0.f, rd0 = 0.f, rd1 = 0.f;
  if (!TRANSPOSED) {
    rl0 = lse2_s[row0 + g], rl1 = lse2_s[row0 + g + 8];
    rd0 = delta_s[row0 + g], rd1 = delta_s[row0 + g + 8];
  }
  for (int nt0 = 0; nt0 < ntiles; nt0 += 4) {
    float s[4][4], dp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      if (nt0 + 2 * p < ntiles) {
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          uint32_t bb[4];
          Bs.load_b_nk(bb, (nt0 + 2 * p) * 8, ks * 16, lane);
          mma_bf16(s[2 * p], xf[ks], bb[0], bb[1]);
          mma_bf16(s[2 * p + 1], xf[ks], bb[2], bb[3]);
          Bd.load_b_nk(bb, (nt0 + 2 * p) * 8, ks * 16, lane);
          mma_bf16(dp[2 * p], yf[ks], bb[0], bb[1]);
          mma_bf16(dp[2 * p + 1], yf[ks], bb[2], bb[3]);
        }
      }
    }
    // P = exp2(S*c - lse2), dS = P * (dP - delta)   (the trailing `scale` is applied once to the accumulators)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float la, lb, da, db;   // statistics of the two columns (transposed) -- or of the two rows (not transposed)
      if (TRANSPOSED) {
        const int col = (nt0 + i) * 8 + 2 * t;
        la = lse2_s[col], lb = lse2_s[col + 1], da = delta_s[col], db = delta_s[col + 1];
        const float p0 = exp2f(s[i][0] * c - la), p1 = exp2f(s[i][1] * c - lb), p2 = exp2f(s[i][2] * c - la), p3 = exp2f(s[i][3] * c - lb);
        dp[i][0] = p0 * (dp[i][0] - da), dp[i][1] = p1 * (dp[i][1] - db), dp[i][2] = p2 * (dp[i][2] - da), dp[i][3] = p3 * (dp[i][3] - db);
        s[i][0] = p0, s[i][1] = p1, s[i][2] = p2, s[i][3] = p3;
      } else {
        const float p0 = exp2f(s[i][0] * c - rl0), p1 = exp2f(s[i][1] * c - rl0), p2 = exp2f(s[i][2] * c - rl1), p3 = exp2f(s[i][3] * c - rl1);
        dp[i][0] = p0 * (dp[i][0] - rd0), dp[i][1] = p1 * (dp[i][1] - rd0), dp[i][2] = p2 * (dp[i][2] - rd1), dp[i][3] = p3 * (dp[i][3] - rd1);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (nt0 + 2 * j < ntiles) {
        uint32_t pa[4], da[4];
        if (TRANSPOSED) {
          pa[0] = pack_bf16(s[2 * j][0], s[2 * j][1]), pa[1] = pack_bf16(s[2 * j][2], s[2 * j][3]);
          pa[2] = pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]), pa[3] = pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3]);
        }
        da[0] = pack_bf16(dp[2 * j][0], dp[2 * j][1]), da[1] = pack_bf16(dp[2 * j][2], dp[2 * j][3]);
        da[2] = pack_bf16(dp[2 * j + 1][0], dp[2 * j + 1][1]), da[3] = pack_bf16(dp[2 * j + 1][2], dp[2 * j + 1][3]);
#pragma unroll
        for (int dd = 0; dd < D / 16; ++dd) {
          uint32_t bb[4];
          if (TRANSPOSED) {
            Bd.load_b_kn(bb, (nt0 + 2 * j) * 8, dd * 16, lane);     // dV += P^T dO
            mma_bf16(acc1[2 * dd], pa, bb[0], bb[1]);
            mma_bf16(acc1[2 * dd + 1], pa, bb[2], bb[3]);
          }
          Bs.load_b_kn(bb, (nt0 + 2 * j) * 8, dd * 16, lane);       // dK += dS^T Q   |   dQ += dS K
          mma_bf16(acc2[2 * dd], da, bb[0], bb[1]);
          mma_bf16(acc2[2 * dd + 1], da, bb[2], bb[3]);
        }
      }
    }
  }
}

template <int D>
__device__ __forceinline__ void store_block(bf16* dst, long ld, int row0, int N, int lane, const float (&acc)[D / 8][4], float mul) {
  const int g = lane >> 2, t = lane & 3;
  const int r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    const int col = i * 8 + 2 * t;
    if (r0 < N) *reinterpret_cast<uint32_t*>(dst + (long)r0 * ld + col) = pack_bf16(acc[i][0] * mul, acc[i][1] * mul);
    if (r1 < N) *reinterpret_cast<uint32_t*>(dst + (long)r1 * ld + col) = pack_bf16(acc[i][2] * mul, acc[i][3] * mul);
  }
}

// Column sums of a row block's gradient (rows < N only) into per-CTA shared accumulators: the qkv bias gradient.
template <int D>
__device__ __forceinline__ void colsum_block(float* cs, int row0, int N, int lane, const float (&acc)[D / 8][4], float mul) {
  const int g = lane >> 2, t = lane & 3;
  const bool v0 = row0 + g < N, v1 = row0 + g + 8 < N;
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    float a = (v0 ? acc[i][0] : 0.f) + (v1 ? acc[i][2] : 0.f), b = (v0 ? acc[i][1] : 0.f) + (v1 ? acc[i][3] : 0.f);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (g == 0) {
      atomicAdd(cs + i * 8 + 2 * t, a * mul);
      atomicAdd(cs + i * 8 + 2 * t + 1, b * mul);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(384, 1) attn_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                                    const float* __restrict__ lse, bf16* __restrict__ dqkv, int N, int H, int Hk, float scale,
                                    float* __restrict__ dbias) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float cs[3 * D];
  const int h = blockIdx.x, b = blockIdx.y;
  const long ldq = 3L * H * D, ldo = (long)H * D;
  bf16* dbase = dqkv + (long)b * N * ldq + h * D;
  if (h >= Hk) {
    zero_slice(dbase, ldq, N, D);
    zero_slice(dbase + (long)H * D, ldq, N, D);
    zero_slice(dbase + 2L * H * D, ldq, N, D);
    return;
  }
  const int Np = (N + 15) / 16 * 16;
  const uint32_t sbase = smem_u32(smem);
  constexpr int TB = (D + 8) * 2;
  Tile<D> Qs{sbase}, Ks{sbase + (uint32_t)Np * TB}, Vs{sbase + 2u * Np * TB}, Os{sbase + 3u * Np * TB};   // Os holds dO
  float* lse2_s = reinterpret_cast<float*>(smem + 4 * (size_t)Np * TB);
  float* delta_s = lse2_s + Np + STAT_PAD;
  const bf16* base = qkv + (long)b * N * ldq + h * D;
  const bf16* ob = o + (long)b * N * ldo + h * D;
  const bf16* dob = d_o + (long)b * N * ldo + h * D;
  stage_tile<D>(smem, Qs, base, ldq, N, Np, sbase);
  stage_tile<D>(smem, Ks, base + (long)H * D, ldq, N, Np, sbase);
  stage_tile<D>(smem, Vs, base + 2L * H * D, ldq, N, Np, sbase);
  stage_tile<D>(smem, Os, dob, ldo, N, Np, sbase);
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) cs[i] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // delta_i = dO_i . O_i and lse in log2 units (padded rows get finite zeros; the padding keeps the 32-column chunk loads in
  // bounds).  One row per thread: all 2*D/8 16-byte loads of a row are independent and in flight together.
  for (int r = threadIdx.x; r < Np + STAT_PAD; r += blockDim.x) {
    float acc = 0.f;
    if (r < N) {
#pragma unroll
      for (int cc = 0; cc < D; cc += 8) {
        const uint4 ov = *reinterpret_cast<const uint4*>(ob + (long)r * ldo + cc);
        const uint4 dv = *reinterpret_cast<const uint4*>(dob + (long)r * ldo + cc);
        const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[j]));
          const float2 d2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dw[j]));
          acc += a.x * d2.x + a.y * d2.y;
        }
      }
    }
    delta_s[r] = acc;
    lse2_s[r] = r < N ? lse[((long)b * H + h) * N + r] * LOG2E : 0.f;
  }
  __syncthreads();
  const float c = scale * LOG2E;
  const int ntiles = Np / 8, nrb = (N + 15) / 16;
  // One work list: items [0, nrb) are key-owner row blocks (dK, dV), items [nrb, 2 nrb) query-owner row blocks (dQ).  With
  // 257 tokens that is 34 similar items over 12 warps: three nearly full rounds.
  for (int item = warp; item < 2 * nrb; item += nwarps) {
    uint32_t xf[D / 16][4], yf[D / 16][4];
    if (item < nrb) {
      const int row0 = item * 16;
      float dv[D / 8][4], dk[D / 8][4];
#pragma unroll
      for (int i = 0; i < D / 8; ++i) dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        Ks.load_a(xf[ks], row0, ks * 16, lane);
        Vs.load_a(yf[ks], row0, ks * 16, lane);
      }
      bwd_block<D, true>(xf, yf, Qs, Os, lse2_s, delta_s, row0, ntiles, c, lane, dv, dk);
      store_block<D>(dbase + (long)H * D, ldq, row0, N, lane, dk, scale);
      store_block<D>(dbase + 2L * H * D, ldq, row0, N, lane, dv, 1.0f);
      if (dbias != nullptr) {
        colsum_block<D>(cs + D, row0, N, lane, dk, scale);
        colsum_block<D>(cs + 2 * D, row0, N, lane, dv, 1.0f);
      }
    } else {
      const int row0 = (item - nrb) * 16;
      float unused[D / 8][4], dq[D / 8][4];
#pragma unroll
      for (int i = 0; i < D / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        Qs.load_a(xf[ks], row0, ks * 16, lane);
        Os.load_a(yf[ks], row0, ks * 16, lane);
      }
      bwd_block<D, false>(xf, yf, Ks, Vs, lse2_s, delta_s, row0, ntiles, c, lane, unused, dq);
      store_block<D>(dbase, ldq, row0, N, lane, dq, scale);
      if (dbias != nullptr) colsum_block<D>(cs, row0, N, lane, dq, scale);
    }
  }
  if (dbias != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) atomicAdd(dbias + (long)(i / D) * H * D + h * D + (i % D), cs[i]);
  }
}

// Warps per CTA.  Forward: 8 warps, <= 128 registers, two CTAs per SM (16 warps).  Backward: one CTA of 12 warps per SM
// (~168 registers per thread fill the register file) walking the merged key-owner / query-owner work list.
int warps_for(int N, bool bwd) {
  const int rb = (N + 15) / 16;
  if (bwd) return 2 * rb >= 12 ? 12 : 2 * rb;      // one item per warp when there are few
  return rb > 8 ? 8 : rb;
}

template <int D>
int launch_fwd(const void* qkv, void* o, float* lse, int B, int N, int H, int Hk, float scale, cudaStream_t st) {
  const int Np = (N + 15) / 16 * 16;
  const size_t smem = 2 * (size_t)Np * (D + 8) * 2;
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_mma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("vsx_attn_fwd: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
    return VSX_ERR_CUDA;
  }
  cudaFuncSetAttribute(attn_fwd_mma_kernel<D>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  launch_pdl(attn_fwd_mma_kernel<D>, dim3(dim3(H, B)), dim3(warps_for(N, false) * 32), smem, st, (const bf16*)qkv, (bf16*)o, lse, N, H, Hk, scale);
  return check_launch("vsx_attn_fwd");
}
template <int D>
int launch_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int B, int N, int H, int Hk, float scale,
               float* dbias, cudaStream_t st) {
  const int Np = (N + 15) / 16 * 16;
  const size_t smem = 4 * (size_t)Np * (D + 8) * 2 + 2 * (size_t)(Np + STAT_PAD) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_mma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("vsx_attn_bwd: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
    return VSX_ERR_CUDA;
  }
  launch_pdl(attn_bwd_mma_kernel<D>, dim3(dim3(H, B)), dim3(warps_for(N, true) * 32), smem, st, (const bf16*)qkv, (const bf16*)o, (const bf16*)d_o, lse, (bf16*)dqkv, N,
                                                                    H, Hk, scale, dbias);
  return check_launch("vsx_attn_bwd");
}

}  // namespace

bool attn_mma_supported(int N, int D) {
  if (!(D == 32 || D == 48 || D == 64)) return false;
  const int Np = (N + 15) / 16 * 16;
  return 4 * (size_t)Np * (D + 8) * 2 + 2 * (size_t)(Np + STAT_PAD) * sizeof(float) + 3 * 64 * 4 <= 227 * 1024;
}

int attn_fwd_mma(const void* qkv, void* o, float* lse, int B, int N, int H, int D, int Hk, float scale, cudaStream_t st) {
  switch (D) {
    case 32: return launch_fwd<32>(qkv, o, lse, B, N, H, Hk, scale, st);
    case 48: return launch_fwd<48>(qkv, o, lse, B, N, H, Hk, scale, st);
    case 64: return launch_fwd<64>(qkv, o, lse, B, N, H, Hk, scale, st);
  }
  set_error("attn_fwd_mma: unsupported head_dim %d", D);
  return VSX_ERR_ARG;
}

int attn_bwd_mma(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int B, int N, int H, int D, int Hk,
                 float scale, float* dbias, cudaStream_t st) {
  switch (D) {
    case 32: return launch_bwd<32>(qkv, o, d_o, lse, dqkv, B, N, H, Hk, scale, dbias, st);
    case 48: return launch_bwd<48>(qkv, o, d_o, lse, dqkv, B, N, H, Hk, scale, dbias, st);
    case 64: return launch_bwd<64>(qkv, o, d_o, lse, dqkv, B, N, H, Hk, scale, dbias, st);
  }
  set_error("attn_bwd_mma: unsupported head_dim %d", D);
  return VSX_ERR_ARG;
}

}  // namespace vsx
