// tcgen05 / TMEM / TMA attention for the ViT-Res shapes (head_dim 64, 48 or 32; N = 257 / 65 / 17 tokens, N <= 288), bf16 in / fp32 softmax.
// Head dims below 64 (sr_tiny_mh, sr_small, the searched Medium network) run the SAME kernels: the 4-D tensor maps [B][N][heads][D] load
// 64-column boxes whose columns >= D are out of bounds and therefore hardware zero-filled, so every shared-memory tile keeps the
// 128-byte swizzled row of the D = 64 layout; the reductions over the head dim stop after D / 16 k steps and results are written D wide.
//
// Replaces q@k^T*scale -> softmax -> attn@v and their autograd (nets/supernet_blocks.py:103-112).  Nothing of size N x N
// touches HBM (the reference materialises [B,H,N,N] three times); masked heads are never computed.
//
// Both kernels are persistent (one CTA per SM walking a list of (sample, head) pairs) and warp specialised:
//   warp 0      TMA producer: 4-D tensor maps [B][N][heads][D] so that rows >= N and columns >= D are hardware zero-filled
//   warp 1      MMA issuer  : one thread issues every tcgen05.mma; tcgen05.commit publishes results / frees operand buffers
//   warps 2..9  softmax: thread = TMEM lane = one query row; two warps share a lane quarter and split the columns; exp2 on the SFU,
//               P / dS written as bf16 into 128B-swizzled shared-memory atoms that the next MMA reads
//   backward only: warps 10..13 epilogue (row statistics of the next pair, accumulator drains, bias-gradient column sums)
//   odd token (below): forward warps 10..13, backward warps 14..15 = side warps on the CUDA cores
//
// forward, per 128-query tile:  S[128 x N] = Q K^T (TMEM) -> row max / exp2 / row sum -> P (smem) -> O[128 x 64] = P V (TMEM)
//                               -> O / l -> global, lse = max*scale + ln(l)
// backward, key blocks j of 64 keys (outer) x query tiles i of 128 (inner), everything accumulated in tensor memory:
//     S = Q_i K_j^T, dP = dO_i V_j^T            (TMEM columns 0..63, 64..127)
//     P = exp2(S*c - lse), dS = P o (dP - delta) (registers -> smem, bf16, double buffered)
//     [dV_j | dK_j] += [P^T ; dS^T] [Q_i | dO_i] (ONE chain, TMEM columns 128..255: MN-major A and B operands, no transposes)
//     dQ_i += dS K_j                             (TMEM columns 256 + 64 i: all query tiles stay resident)
//   the qkv-bias gradient (column sums of dQ, dK, dV) is reduced with a shuffle butterfly into shared memory and flushed once per CTA
//   (every CTA works on ONE head).
// The odd token: N = 2^k + 1 tokens (257, 65) leave one token over after tiles of 128 queries x 64 keys; it can run on side warps instead of
// costing a query tile and a key block of its own (section "the odd token" below; then [dV|dK] is double buffered at TMEM columns 128..383
// and the two dQ tiles sit at 384..511).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace vsx {

int make_tmap_heads(CUtensorMap* m, const void* base, uint64_t head_dim, uint64_t heads, uint64_t rows, uint64_t batch, uint32_t box_rows);

namespace {

constexpr int KV_BOX = 96;                 // rows per K / V TMA box (= backward key block)
constexpr int ATT_MAX_N = 288;             // 3 boxes of keys, 3 query tiles
constexpr int TILE16K = 128 * 128;         // [128 rows][64 bf16] tile = one swizzle atom column
constexpr int BOX12K = KV_BOX * 128;
constexpr int ATT_THREADS = 320;           // 10 warps (+ 4 side warps with an odd token: ATT_THREADS_ODD)
constexpr int ATT_THREADS_ODD = 448;
constexpr int F_SIDE_WARP0 = 10;
constexpr int SM_WARP0 = 2;                // first softmax warp
constexpr int SM_THREADS = 256;
constexpr float LOG2E_F = 1.4426950408889634f;
constexpr float LN2_F = 0.6931471805599453f;

struct AttnMaps {
  CUtensorMap q;    // qkv as [B][N][3*H heads][D], box 64 x 1 x 128 x 1 (columns >= D zero-filled)
  CUtensorMap kv;   // qkv, box 64 x 1 x 96 (forward) / 64 (backward) x 1
  CUtensorMap d_o;  // d_o as [B][N][H][D], box 64 x 1 x 128 x 1 (backward only)
};

// Several head extents in one launch (multi-architecture batches): consecutive sample ranges with their own number of kept heads.
// count == 0: every sample keeps Hk heads.  With segments, Hk is the LARGEST kept-head count.
struct AttnSegs {
  int count;
  int b_end[VSX_MAX_SEGMENTS];      // exclusive end sample of segment i
  int hk[VSX_MAX_SEGMENTS];         // kept heads of segment i (0: its samples are skipped)
  int w_end[VSX_MAX_SEGMENTS];      // exclusive end of segment i in the flattened (sample, head) work list of the forward kernel
  int cta_end[16];                  // backward: exclusive end of the CTA range that works on head h (CTAs in proportion to the samples keeping h)
};

struct AttnArgs {
  AttnSegs segs;
  int B, N, H, Hk, D;
  int odd;             // the last token of every (sample, head) pair runs on CUDA-core side paths (see "odd token" below):
                       // bit 0 as a query (row t of S), bit 1 as a key (column t of S; backward only)
  float scale;
  const bf16* qkv;     // the side paths read rows of qkv directly
  bf16* o;             // fwd: output; bwd: forward output (for delta)
  const bf16* d_o;     // bwd
  float* lse;          // fwd: written; bwd: read
  bf16* dqkv;          // bwd
  float* dbias;        // bwd, may be null
  long long* dbg;      // optional timeline buffer (tools/attn_timeline.py); null in production
  int dbg_cta;         // the CTA that writes the per-block stamps (VSX_ATTN_DBG_CTA)
};

// one [rows][64] tile of head `head` (index into the 3*H or H head slots), rows row0.. of sample b
__device__ __forceinline__ void tma_load_head(uint32_t dst, const CUtensorMap* m, uint32_t bar, int head, int row0, int b) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(0), "r"(head), "r"(row0), "r"(b)
      : "memory");
}
// forward work item w -> (sample, head)
__device__ __forceinline__ void pair_of(const AttnSegs& sg, int Hk, int w, int& b, int& h) {
  if (sg.count == 0) {
    b = w / Hk, h = w % Hk;
    return;
  }
  int i = 0, w0 = 0, b0 = 0;
  while (i < sg.count - 1 && w >= sg.w_end[i]) w0 = sg.w_end[i], b0 = sg.b_end[i], ++i;
  const int l = w - w0;
  b = b0 + l / sg.hk[i], h = l % sg.hk[i];
}
// backward: the v-th sample (in batch order) that keeps head h, and how many there are
__device__ __forceinline__ int sample_of(const AttnSegs& sg, int h, int v) {
  if (sg.count == 0) return v;
  int acc = 0, b0 = 0;
  for (int i = 0; i < sg.count; ++i) {
    if (sg.hk[i] > h) {
      const int nb = sg.b_end[i] - b0;
      if (v < acc + nb) return b0 + (v - acc);
      acc += nb;
    }
    b0 = sg.b_end[i];
  }
  return b0;      // not reached for v < samples_with(sg, h, B)
}
__device__ __forceinline__ int samples_with(const AttnSegs& sg, int h, int B) {
  if (sg.count == 0) return B;
  int acc = 0, b0 = 0;
  for (int i = 0; i < sg.count; ++i) {
    if (sg.hk[i] > h) acc += sg.b_end[i] - b0;
    b0 = sg.b_end[i];
  }
  return acc;
}

// `ncols` (a multiple of 8, <= 32) of 32 fp32 values -> bf16 -> global
__device__ __forceinline__ void store_bf16_n(bf16* dst, const float (&v)[32], int ncols) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (8 * i < ncols) {
      uint4 r;
      r.x = pack_bf16(v[8 * i], v[8 * i + 1]);
      r.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
      r.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
      r.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
      *reinterpret_cast<uint4*>(dst + 8 * i) = r;
    }
  }
}

// Shared-memory matrix descriptors (128B swizzle, 8-row groups 1024 B apart).
// K-major: rows of 64 contiguous k elements (128 B); a 16-wide k step advances the start address by 32 B.
__device__ __forceinline__ uint64_t desc_k(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major: rows are k, 64 contiguous MN elements per row (128 B); a 16-wide k step advances by 16 rows = 2048 B;
// lbo = distance between consecutive 64-element MN atoms.
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo_bytes) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, M = 128.
__device__ __forceinline__ uint32_t idesc_m128(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a [rows][128 B] 128B-swizzled atom column
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// 32 consecutive columns (col0 % 32 == 0) of row `row` -> bf16 -> staging [atoms of 64 columns][128 rows][128 B]
__device__ __forceinline__ void stage_bf16_32(uint8_t* stg, int row, int col0, const float (&v)[32]) {
  uint8_t* atom = stg + (col0 >> 6) * TILE16K;
  const int ch0 = (col0 & 63) >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 r;
    r.x = pack_bf16(v[8 * i], v[8 * i + 1]);
    r.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    r.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    r.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(atom + swz(row, ch0 + i)) = r;
  }
}

// Column sums over the 32 lanes of a warp: on return lane l holds sum_lanes v[l] (31 shuffles: recursive halving).
__device__ __forceinline__ float butterfly_colsum(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; ++k) {
      const float send = up ? v[k] : v[k + s];
      const float keep = up ? v[k + s] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

__device__ __forceinline__ int round16(int x) { return (x + 15) & ~15; }

// One lane of a converged warp.  The MMA warp runs its loops converged (so descriptor arithmetic stays on the uniform
// datapath instead of per-thread registers + R2UR round trips) and only the tcgen05 instructions are predicated on the leader.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// 16 consecutive columns (col0 % 16 == 0) of row `row` -> bf16 -> staging
__device__ __forceinline__ void stage_bf16_16(uint8_t* stg, int row, int col0, const float (&v)[16]) {
  uint8_t* atom = stg + (col0 >> 6) * TILE16K;
  const int ch0 = (col0 & 63) >> 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 r;
    r.x = pack_bf16(v[8 * i], v[8 * i + 1]);
    r.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    r.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    r.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(atom + swz(row, ch0 + i)) = r;
  }
}

// Producer-side wait: the TMA thread is never on the critical path, so it backs off instead of competing for issue slots
// with the softmax warps of its scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (clock64() - t0 > 3000000000LL) {
      printf("vsx: attention producer wait timed out (block %d)\n", blockIdx.x);
      __trap();
    }
  }
}

// Zero the slices of the masked heads (columns [col0, col0 + ncols) of `nslices` feature groups `slice_stride` apart) for this
// CTA's share of the rows: consumers contract over the full feature width in places, so masked slices are defined as zero.
__device__ __forceinline__ void zero_masked(bf16* base, long ld, long rows, int col0, int ncols, int nslices, long slice_stride, int tid,
                                            int nthreads) {
  if (ncols <= 0) return;
  const int cpr = ncols >> 3, per_row = cpr * nslices;
  const long my_rows = (rows - blockIdx.x + gridDim.x - 1) / gridDim.x;
  for (long idx = tid; idx < my_rows * per_row; idx += nthreads) {
    const long rl = idx / per_row;
    const int i = (int)(idx - rl * per_row), sl = i / cpr, ch = i - sl * cpr;
    const long r = blockIdx.x + rl * gridDim.x;
    *reinterpret_cast<uint4*>(base + r * ld + sl * slice_stride + col0 + ch * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ------------------------------------------------------------------------------------------------ the odd token
// The ViT-Res token counts are 2^k patch tokens + the class token (N = 257 / 65).  On tiles of 128 queries x 64 keys that one token costs a
// whole extra query tile and a whole extra key block per (sample, head) pair -- 15 blocks instead of 8 in the backward at N = 257, and every
// block costs the same latency-bound chain whatever its width.  With `odd` set the tensor-core tiles cover the first N - 1 tokens only and
// the LAST token t = N - 1 runs on CUDA-core side paths of otherwise idle warps, O(N D) work per pair:
//   as a QUERY  (row t of S):     s_tj, p_tj for every key j -> O_t, lse_t (forward); dQ_t = sum_j ds_tj k_j and the rank-1 terms
//                                 dK_j += ds_tj q_t, dV_j += p_tj dO_t, which the accumulator drains add (backward)
//   as a KEY    (column t of S):  forward: still part of the MMA (one more 16-column step); backward: p_it, ds_it for every query i ->
//                                 dK_t = sum_i ds_it q_i, dV_t = sum_i p_it dO_i and the rank-1 term dQ_i += ds_it k_t (added by the dQ drain)
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float dot8(const uint4& w, const float4& x, const float4& y) {
  return (bf_lo(w.x) * x.x + bf_hi(w.x) * x.y + bf_lo(w.y) * x.z + bf_hi(w.y) * x.w) +
         (bf_lo(w.z) * y.x + bf_hi(w.z) * y.y + bf_lo(w.w) * y.z + bf_hi(w.w) * y.w);
}
// v[0..31] += w * vec[0..31] (vec in shared memory, broadcast reads)
__device__ __forceinline__ void axpy32(float (&v)[32], float w, const float* vec) {
  const float4* v4 = reinterpret_cast<const float4*>(vec);
#pragma unroll
  for (int t4 = 0; t4 < 8; ++t4) {
    const float4 x = v4[t4];
    v[4 * t4] = fmaf(w, x.x, v[4 * t4]), v[4 * t4 + 1] = fmaf(w, x.y, v[4 * t4 + 1]);
    v[4 * t4 + 2] = fmaf(w, x.z, v[4 * t4 + 2]), v[4 * t4 + 3] = fmaf(w, x.w, v[4 * t4 + 3]);
  }
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
  return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, s));
  return x;
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int F_K = 0, F_V = 3 * BOX12K, F_Q = 6 * BOX12K, F_P = F_Q + 2 * TILE16K, F_END = F_P + 5 * TILE16K;
constexpr int F_SIDE = 64 + 320 + 16 + 256;   // floats of the odd-token side path: q_t, p_tj, reduction scratch, partial O_t of four warps
constexpr int F_SMEM = F_END + 1024 /*align*/ + 128 /*barriers*/ + 4 * 128 * 4 /*xm, xl*/ + F_SIDE * 4;
constexpr int F_OCOL = 448;   // O accumulator columns 448..511; S occupies 0..287

__global__ void __launch_bounds__(ATT_THREADS_ODD, 1) attn_fwd_tc_kernel(const __grid_constant__ AttnMaps maps, const AttnArgs a) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar0 = base + F_END;
  // barriers: 0 k_full, 1 k_empty, 2 v_full, 3 v_empty, 4-5 q_full, 6-7 q_empty, 8 s_full, 9 p_ready, 10 o_full
  auto bar = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + F_END + 96);
  float* xm = reinterpret_cast<float*>(smem + F_END + 128);   // [2][128] partial row maxima
  float* xl = xm + 256;                                       // [2][128] partial row sums
  float* side = xl + 256;                                     // [F_SIDE] odd-token side path

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, D = a.D, HD = a.H * D, KS = D >> 4;      // KS: 16-wide k steps over the head dim
  const int Nq = N - a.odd;                                   // query rows on the tensor-core tiles (the odd token: side warps)
  const int QT = (Nq + 127) / 128, NKP = round16(N), nkb = (N + KV_BOX - 1) / KV_BOX;
  const int total = a.segs.count ? a.segs.w_end[a.segs.count - 1] : a.B * a.Hk;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.kv);
  }
  if (warp == 1) {
    if (lane == 0) {
      // k_empty / v_empty: the MMA commit + (odd token) one arrive per side warp
      for (int i = 0; i < 11; ++i) mbar_init(bar(i), i == 9 ? 8 : ((i == 1 || i == 3) && a.odd) ? 5 : 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int nw = 0, qit = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++nw) {
        int b, h;
        pair_of(a.segs, a.Hk, w, b, h);
        mbar_wait_relaxed(bar(1), ((uint32_t)nw & 1u) ^ 1u);
        mbar_expect_tx(bar(0), (uint32_t)nkb * BOX12K);
        for (int c = 0; c < nkb; ++c) tma_load_head(base + F_K + c * BOX12K, &maps.kv, bar(0), a.H + h, c * KV_BOX, b);
        for (int i = 0; i < QT; ++i, ++qit) {
          const int s = qit & 1;
          mbar_wait_relaxed(bar(6 + s), (((uint32_t)qit >> 1) & 1u) ^ 1u);
          mbar_expect_tx(bar(4 + s), TILE16K);
          tma_load_head(base + F_Q + s * TILE16K, &maps.q, bar(4 + s), h, i * 128, b);
          if (i == 0) {
            mbar_wait_relaxed(bar(3), ((uint32_t)nw & 1u) ^ 1u);
            mbar_expect_tx(bar(2), (uint32_t)nkb * BOX12K);
            for (int c = 0; c < nkb; ++c) tma_load_head(base + F_V + c * BOX12K, &maps.kv, bar(2), 2 * a.H + h, c * KV_BOX, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    int nw = 0, qit = 0, blk = 0;
    const uint32_t id_o = idesc_m128(64, false, true);
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++nw) {
      mbar_wait(bar(0), (uint32_t)nw & 1u);
      for (int i = 0; i < QT; ++i, ++qit, ++blk) {
        const int s = qit & 1;
        mbar_wait(bar(4 + s), ((uint32_t)qit >> 1) & 1u);
        tc_fence_after();
        const uint32_t qa = base + F_Q + s * TILE16K;
        for (int c = 0; c < nkb; ++c) {
          const int ncols = min(KV_BOX, NKP - c * KV_BOX);
          const uint32_t id_s = idesc_m128(ncols, false, false);
          const uint64_t dq = desc_k(qa), dk = desc_k(base + F_K + c * BOX12K);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (leader && k < KS) umma_bf16(tmem + c * KV_BOX, dq + 2 * k, dk + 2 * k, id_s, k > 0 ? 1u : 0u);
        }
        if (leader) {
          umma_commit(bar(6 + s));
          umma_commit(bar(8));
          if (i == QT - 1) umma_commit(bar(1));
        }
        __syncwarp();
        mbar_wait(bar(9), (uint32_t)blk & 1u);
        tc_fence_after();
        if (i == 0) {
          mbar_wait(bar(2), (uint32_t)nw & 1u);
          tc_fence_after();
        }
        const uint64_t dp0 = desc_k(base + F_P), dv0 = desc_mn(base + F_V, TILE16K);
        for (int kk = 0; kk < NKP / 16; ++kk) {
          // P atom kk / 4 (16 KB apart = 1024 descriptor units), 32 B (2 units) per k step inside it; V advances 16 rows = 2048 B
          const uint64_t dp = dp0 + (uint64_t)((kk >> 2) * (TILE16K >> 4) + (kk & 3) * 2), dv = dv0 + (uint64_t)(kk * 128);
          if (leader) umma_bf16(tmem + F_OCOL, dp, dv, id_o, kk > 0 ? 1u : 0u);
        }
        if (leader) {
          umma_commit(bar(10));
          if (i == QT - 1) umma_commit(bar(3));
        }
        __syncwarp();
      }
    }
  } else if (warp < F_SIDE_WARP0) {
    const int q = warp & 3, hf = (warp - SM_WARP0) >> 2, row = q * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const float c = a.scale * LOG2E_F;
    const int nch = (NKP + 31) / 32;
    uint8_t* Ps = smem + F_P;
    int blk = 0;
    if (a.segs.count == 0) {
      zero_masked(a.o, HD, (long)a.B * N, a.Hk * D, (a.H - a.Hk) * D, 1, 0, threadIdx.x - SM_WARP0 * 32, SM_THREADS);
    } else {
      for (int i = 0, b0 = 0; i < a.segs.count; b0 = a.segs.b_end[i], ++i)
        if (a.segs.hk[i] > 0)
          zero_masked(a.o + (long)b0 * N * HD, HD, (long)(a.segs.b_end[i] - b0) * N, a.segs.hk[i] * D, (a.H - a.segs.hk[i]) * D, 1, 0,
                      threadIdx.x - SM_WARP0 * 32, SM_THREADS);
    }
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      int b, h;
      pair_of(a.segs, a.Hk, w, b, h);
      for (int i = 0; i < QT; ++i, ++blk) {
        const int rows_valid = min(128, Nq - i * 128);
        const bool active = q * 32 < rows_valid;
        mbar_wait(bar(8), (uint32_t)blk & 1u);
        tc_fence_after();
        float m = -INFINITY;
        if (active) {
          for (int cc = hf; cc < nch; cc += 2) {
            float v[32];
            tmem_ld32(tlane + cc * 32, v);
            tmem_ld_wait();
            if (cc * 32 + 32 <= N) {      // only the last chunk needs the key mask
#pragma unroll
              for (int j = 0; j < 32; ++j) m = fmaxf(m, v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) m = fmaxf(m, cc * 32 + j < N ? v[j] : -INFINITY);
            }
          }
        }
        xm[hf * 128 + row] = m;
        named_bar_sync(1, SM_THREADS);
        m = fmaxf(xm[row], xm[128 + row]);
        const float mc = m * c;
        float l = 0.f;
        if (active) {
          for (int cc = hf; cc < nch; cc += 2) {
            float v[32];
            tmem_ld32(tlane + cc * 32, v);
            tmem_ld_wait();
            if (cc * 32 + 32 <= N) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                v[j] = ex2(fmaf(v[j], c, -mc));
                l += v[j];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float p = cc * 32 + j < N ? ex2(fmaf(v[j], c, -mc)) : 0.f;
                l += p;
                v[j] = p;
              }
            }
            stage_bf16_32(Ps, row, cc * 32, v);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(9));
        xl[hf * 128 + row] = l;
        named_bar_sync(2, SM_THREADS);
        l = xl[row] + xl[128 + row];
        mbar_wait(bar(10), (uint32_t)blk & 1u);
        tc_fence_after();
        if (active) {
          float v[32];
          tmem_ld32(tlane + F_OCOL + hf * 32, v);
          tmem_ld_wait();
          if (row < rows_valid) {
            const float inv = 1.0f / l;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= inv;
            const long r = (long)b * N + i * 128 + row;
            store_bf16_n(a.o + r * HD + h * D + hf * 32, v, D - hf * 32);
            if (hf == 0) a.lse[((long)b * a.H + h) * N + i * 128 + row] = (mc + log2f(l)) * LN2_F;
          }
        }
        tc_fence_before();
      }
    }
  } else if (a.odd) {
    // ---------------- side warps: the odd token as a query (row t of S on the CUDA cores; K and V are read from the pair's TMA tiles) ----------------
    const int st = threadIdx.x - F_SIDE_WARP0 * 32, sw = st >> 5, t = N - 1;
    float* qts = side;                 // [64]  q_t
    float* pt = side + 64;             // [320] p_tj
    float* red = pt + 320;             // [4] partial maxima, [4] partial sums
    float* oacc = red + 16;            // [4][64] partial O_t
    const float c = a.scale * LOG2E_F;
    int nw = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++nw) {
      int b, h;
      pair_of(a.segs, a.Hk, w, b, h);
      const long grow = (long)b * N + t;
      if (st < 64) qts[st] = st < D ? __bfloat162float(a.qkv[grow * 3 * HD + h * D + st]) : 0.f;
      named_bar_sync(3, 128);
      mbar_wait(bar(0), (uint32_t)nw & 1u);
      float sj[3], m = -INFINITY;
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        const int j = st + 128 * rr;
        sj[rr] = -INFINITY;
        if (j < N) {
          const uint8_t* kr = smem + F_K + (j / KV_BOX) * BOX12K;
          const int jl = j % KV_BOX;
          const float4* q4 = reinterpret_cast<const float4*>(qts);
          float acc = 0.f;
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) acc += dot8(*reinterpret_cast<const uint4*>(kr + swz(jl, c8)), q4[2 * c8], q4[2 * c8 + 1]);
          sj[rr] = acc;
          m = fmaxf(m, acc);
        }
      }
      m = warp_max(m);
      if (lane == 0) red[sw] = m;
      named_bar_sync(3, 128);
      m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
      const float mc = m * c;
      float l = 0.f;
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        const int j = st + 128 * rr;
        if (j < N) {
          const float pj = ex2(fmaf(sj[rr], c, -mc));
          pt[j] = pj;
          l += pj;
        }
      }
      l = warp_sum(l);
      if (lane == 0) red[4 + sw] = l;
      named_bar_sync(3, 128);
      l = (red[4] + red[5]) + (red[6] + red[7]);
      mbar_wait(bar(2), (uint32_t)nw & 1u);
      // warp sw takes the keys j = sw, sw + 4, ...; lane l the columns 2l, 2l + 1 (one 128-byte row per warp and step: no bank conflicts)
      float ox = 0.f, oy = 0.f;
      for (int j = sw; j < N; j += 4) {
        const uint8_t* vr = smem + F_V + (j / KV_BOX) * BOX12K;
        const uint32_t wv = *reinterpret_cast<const uint32_t*>(vr + swz(j % KV_BOX, lane >> 2) + (lane & 3) * 4);
        const float pj = pt[j];
        ox = fmaf(pj, bf_lo(wv), ox), oy = fmaf(pj, bf_hi(wv), oy);
      }
      oacc[sw * 64 + 2 * lane] = ox, oacc[sw * 64 + 2 * lane + 1] = oy;
      __syncwarp();
      if (lane == 0) {           // this warp is done with the pair's K and V tiles
        mbar_arrive(bar(1));
        mbar_arrive(bar(3));
      }
      named_bar_sync(3, 128);
      if (st < D) a.o[grow * HD + h * D + st] = __float2bfloat16(((oacc[st] + oacc[64 + st]) + (oacc[128 + st] + oacc[192 + st])) / l);
      if (st == 0) a.lse[((long)b * a.H + h) * N + t] = (mc + log2f(l)) * LN2_F;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ backward
// Key blocks of 64 (one swizzle atom of P / dS per block) so that the P / dS staging can be double buffered: the MMA thread
// issues S / dP of block g+1 BEFORE dV / dK / dQ of block g, and the softmax warps work on block g+1 while the tensor pipe
// drains block g.  tcgen05.commit covers every MMA issued before it, which orders all buffer re-use without extra barriers:
//   sdp_full(g+1) fires after S/dP(g+1) and therefore after dV/dK/dQ(g-1), the last readers of staging buffer (g+1) & 1.
// dV and dK come from ONE MMA chain per block: A = [P^T ; dS^T] (M = 128: the staging atoms P | dS are 16 KB apart = the MN-major
// leading-dimension offset) times B = [Q | dO] (N = 128: the two tiles of a slot, again 16 KB apart):
//   D[0..63, 64..127] = P^T dO = dV,   D[64..127, 0..63] = dS^T Q = dK   (the other two quadrants are by-products)
// -- half the shared-memory operand traffic of two M=128 chains whose upper 64 rows would be padding.
// Four extra warps own everything that is not on the per-block critical path: the per-row statistics (lse, delta = dO . O) of the
// NEXT pair (global loads -> shared memory) and the dK / dV / dQ epilogues (TMEM -> bf16 -> global, bias-gradient column sums).
constexpr int BKW = 64;                    // backward key block
constexpr int BOX8K = BKW * 128;
constexpr int BWD_THREADS = 448;           // 14 warps: producer, MMA, 8 softmax, 4 epilogue
constexpr int BWD_THREADS_ODD = 512;       // + 2 side warps (the odd token)
constexpr int EP_WARP0 = 10, SIDE_WARP0 = 14, SIDE_THREADS = 64;
// staging: [P0 | dS0 | P1 | dS1]
constexpr int B_KV = 0, B_QDO = 4 * BOX8K, B_STG = B_QDO + 6 * TILE16K, B_END = B_STG + 4 * TILE16K;
// odd-token side data (floats).  Per buffer (two, by pair parity): ds_it per query row, p_tj / ds_tj per key, q_t, k_t, dO_t, v_t,
// (lse_t, delta_t); shared by the pairs: p_it per query row, the accumulators of dQ_t / dK_t / dV_t
constexpr int SD_DSA = 0, SD_PB = 384, SD_DSB = 704, SD_QT = 1024, SD_KT = 1088, SD_DOT = 1152, SD_VT = 1216, SD_SCAL = 1280, SD_BUF = 1284;   // 3 tiles, 5 blocks
constexpr int SD_PA = 2 * SD_BUF, SD_ACC = SD_PA + 384, SD_FLOATS = SD_ACC + 192;
constexpr int B_MISC0 = 256 /*barriers*/ + 3 * 64 * 4 /*cs*/ + 2 * 3 * 128 * 8 /*stats*/ + 16 /*tmem slot*/;
constexpr int B_MISC = B_MISC0 + SD_FLOATS * 4;
constexpr int B_SMEM = B_END + 1024 + B_MISC;
constexpr int C_S = 0, C_DP = 64, C_DVK = 128, C_DQ = 256;
// N <= 128 (one query tile): the accumulators are double buffered across key blocks / pairs so that the MMA thread never waits for the
// epilogue warps: [dV|dK] at 128 / 256, dQ at 384 / 448.  N > 128: single buffers, dQ tiles at 256, 320, 384.
// With an odd token there are at most two query tiles: [dV|dK] is double buffered across key blocks (128 / 256) and the dQ tiles sit at 384, 448.
__device__ __forceinline__ int dvk_col(int nbkv, int kvit) { return C_DVK + (nbkv == 2 ? 128 * (kvit & 1) : 0); }
__device__ __forceinline__ int dq_col(int nbkv, int i) { return (nbkv == 2 ? 384 : C_DQ) + 64 * i; }

// Q_i / dO_i tiles stay resident for all key blocks of their (sample, head): three 32 KB slots used as a ring over the global
// tile counter t = (pairs done) * QT + i, loaded once per pair (at j == 0) and released after the last key block.
struct BwdCursor {   // position in this CTA's flattened (sample, key block, query tile) sequence
  int b, j, i, n, kvit, t0, pair;   // t0 = global tile counter of query tile 0 of the current pair
  __device__ __forceinline__ void advance(int QT, int KB, int nslots) {
    ++n;
    if (++i == QT) {
      i = 0;
      ++kvit;
      if (++j == KB) {
        j = 0;
        b += nslots;
        t0 += QT;
        ++pair;
      }
    }
  }
  __device__ __forceinline__ int tile() const { return t0 + i; }
};

__global__ void __launch_bounds__(BWD_THREADS_ODD, 1) attn_bwd_tc_kernel(const __grid_constant__ AttnMaps maps, const AttnArgs a) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar0 = base + B_END;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  // barriers: 0-1 kv_full, 2-3 kv_empty, 4-6 q_full, 7-9 q_empty
  constexpr int BAR_SDP = 10, BAR_PDS = 11, BAR_DKV_FULL = 12, BAR_DKV_EMPTY = 14, BAR_DQ_FULL = 16, BAR_DQ_EMPTY = 18;   // two of each
  constexpr int BAR_SIDE = 20;   // four: side data of key block n % 4 complete (one arrive per side warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + B_END + 256 + 768 + 6144);
  // stats are handed over with named barriers (ids 2, 3: one per buffer): epilogue warps arrive, softmax warps sync
  float* cs = reinterpret_cast<float*>(smem + B_END + 256);        // [3][64] column sums of dQ, dK, dV (this CTA's head)
  float2* stats = reinterpret_cast<float2*>(smem + B_END + 256 + 768);   // [2 buffers][3 tiles][128 rows] (lse * log2e, delta)
  float* sd = reinterpret_cast<float*>(smem + B_END + B_MISC0);          // odd-token side data (SD_*)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, D = a.D, HD = a.H * D, KS = D >> 4;
  // the odd token t = N - 1 as a query / as a key on the side warps: the tensor-core tiles cover Nq query rows and Nk keys
  const int oq = a.odd & 1, ok = (a.odd >> 1) & 1, odd = a.odd != 0 ? 1 : 0, Nq = N - oq, Nk = N - ok;
  const int QT = (Nq + 127) / 128, KB = (Nk + BKW - 1) / BKW;
  // accumulator buffers (see dvk_col / dq_col): [dV|dK] double buffered across key blocks when there are at most two query tiles (with the
  // side warps only: double buffering at N <= 128 without them was measured 10 % SLOWER on the N = 17 launches, profiles/: r1l vs r1k.)
  const int nbkv = (odd && QT <= 2) ? 2 : 1;
  const int stat_threads = SM_THREADS + 128 + (odd ? SIDE_THREADS : 0);    // statistics hand-over: epilogue warps arrive, softmax (and side) warps sync
  // this CTA's head and its samples: gridDim.x is a multiple of Hk
  int h, slot, nslots;
  if (a.segs.count == 0) {
    h = blockIdx.x % a.Hk, slot = blockIdx.x / a.Hk, nslots = gridDim.x / a.Hk;
  } else {            // heads kept by fewer samples get fewer CTAs
    h = 0;
    while (h < a.Hk - 1 && (int)blockIdx.x >= a.segs.cta_end[h]) ++h;
    const int c0 = h == 0 ? 0 : a.segs.cta_end[h - 1];
    slot = blockIdx.x - c0, nslots = a.segs.cta_end[h] - c0;
  }
  // samples that keep this CTA's head, in batch order: the loops below count v = slot, slot + nslots, ... < NB and map v to a sample
  const int NB = samples_with(a.segs, h, a.B);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.kv);
    tma_prefetch_desc(&maps.d_o);
  }
  if (warp == 1) {
    if (lane == 0) {
      // kv_empty / q_empty: the MMA commit (+ one arrive per side warp: they read the same tiles)
      for (int i = 0; i < 24; ++i)
        mbar_init(bar(i), i == BAR_PDS ? 8
                          : i >= BAR_SIDE ? 2
                          : ((i >= BAR_DKV_EMPTY && i < BAR_DKV_EMPTY + 2) || i >= BAR_DQ_EMPTY) ? 4
                          : (((i == 2 || i == 3) && odd) || (i >= 7 && i <= 9 && ok)) ? 3 : 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 512);
  }
  if (threadIdx.x < 192) cs[threadIdx.x] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  if (a.dbg != nullptr && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg[1024 + 2 * blockIdx.x] = t;
  }

  if (warp == 0) {
    if (lane == 0) {
      int kvit = 0, qit = 0;
      for (int v = slot; v < NB; v += nslots) {
        const int b = sample_of(a.segs, h, v);
        for (int j = 0; j < KB; ++j, ++kvit) {
          const int ks = kvit & 1;
          mbar_wait_relaxed(bar(2 + ks), (((uint32_t)kvit >> 1) & 1u) ^ 1u);
          mbar_expect_tx(bar(ks), 2 * BOX8K);
          tma_load_head(base + B_KV + ks * 2 * BOX8K, &maps.kv, bar(ks), a.H + h, j * BKW, b);
          tma_load_head(base + B_KV + ks * 2 * BOX8K + BOX8K, &maps.kv, bar(ks), 2 * a.H + h, j * BKW, b);
          if (j == 0) {
            for (int i = 0; i < QT; ++i, ++qit) {
              const int qs = qit % 3;
              mbar_wait_relaxed(bar(7 + qs), (((uint32_t)qit / 3) & 1u) ^ 1u);
              mbar_expect_tx(bar(4 + qs), 2 * TILE16K);
              tma_load_head(base + B_QDO + qs * 2 * TILE16K, &maps.q, bar(4 + qs), h, i * 128, b);
              tma_load_head(base + B_QDO + qs * 2 * TILE16K + TILE16K, &maps.d_o, bar(4 + qs), h, i * 128, b);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t id_vk = idesc_m128(128, true, true);   // [dV | dK] chain: both operands MN-major, N = 128
    const uint32_t id_q = idesc_m128(64, false, true);    // dQ: A = dS K-major, B = K MN-major
    // S = Q_i K_j^T and dP = dO_i V_j^T of the block under cursor c (waits for its operands)
    auto issue_sdp = [&](const BwdCursor& c) {
      const int ks = c.kvit & 1, qs = c.tile() % 3;
      if (c.i == 0) mbar_wait(bar(ks), ((uint32_t)c.kvit >> 1) & 1u);
      if (c.j == 0) mbar_wait(bar(4 + qs), ((uint32_t)c.tile() / 3) & 1u);
      tc_fence_after();
      const int kw = round16(min(BKW, Nk - c.j * BKW));
      const uint32_t id_s = idesc_m128(kw, false, false);
      const uint32_t ka = base + B_KV + ks * 2 * BOX8K, qa = base + B_QDO + qs * 2 * TILE16K;
      const uint64_t dq = desc_k(qa), dk = desc_k(ka), ddo = desc_k(qa + TILE16K), dv = desc_k(ka + BOX8K);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (leader && k < KS) umma_bf16(tmem + C_S, dq + 2 * k, dk + 2 * k, id_s, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (leader && k < KS) umma_bf16(tmem + C_DP, ddo + 2 * k, dv + 2 * k, id_s, k > 0 ? 1u : 0u);
      if (leader) umma_commit(bar(BAR_SDP));
      __syncwarp();
    };
    BwdCursor c = {slot, 0, 0, 0, 0, 0, 0};
    if (c.b < NB) issue_sdp(c);
    BwdCursor nx = c;
    nx.advance(QT, KB, nslots);
    while (c.b < NB) {
      mbar_wait(bar(BAR_PDS), (uint32_t)c.n & 1u);      // P / dS of block c staged; S / dP columns free again
      tc_fence_after();
      if (a.dbg != nullptr && (int)blockIdx.x == (a.dbg_cta & 0xffff) && leader && c.n < 64) a.dbg[c.n * 16 + 0] = clock64();
      if (nx.b < NB) issue_sdp(nx);
      if (a.dbg != nullptr && (int)blockIdx.x == (a.dbg_cta & 0xffff) && leader && c.n < 64) a.dbg[c.n * 16 + 1] = clock64();
      // accumulators about to be overwritten (first block of a key block / of a pair) must have been drained by the epilogue warps
      if (c.i == 0) mbar_wait(bar(BAR_DKV_EMPTY + c.kvit % nbkv), ((uint32_t)(c.kvit / nbkv) & 1u) ^ 1u);
      if (c.i == 0 && c.j == 0) mbar_wait(bar(BAR_DQ_EMPTY), ((uint32_t)c.pair & 1u) ^ 1u);
      tc_fence_after();
      if (a.dbg != nullptr && (int)blockIdx.x == (a.dbg_cta & 0xffff) && leader && c.n < 64) a.dbg[c.n * 16 + 7] = clock64();
      const int ks = c.kvit & 1, qs = c.tile() % 3;
      const uint32_t ka = base + B_KV + ks * 2 * BOX8K, qa = base + B_QDO + qs * 2 * TILE16K;
      const uint32_t ps = base + B_STG + (c.n & 1) * 2 * TILE16K;
      const uint64_t d_pds = desc_mn(ps, TILE16K), d_qdo = desc_mn(qa, TILE16K), d_dsk = desc_k(ps + TILE16K), d_k = desc_mn(ka, TILE16K);
      const int kw = round16(min(BKW, Nk - c.j * BKW));
      const int kq = (min(128, Nq - c.i * 128) + 15) / 16;    // 16-row k steps over the valid queries of this tile
      const uint32_t acc_i = c.i != 0 ? 1u : 0u, acc_j = c.j != 0 ? 1u : 0u;
#pragma unroll
      for (int k = 0; k < 8; ++k)       // a 16-row k step = 2048 B = 128 descriptor units
        if (leader && k < kq) umma_bf16(tmem + dvk_col(nbkv, c.kvit), d_pds + 128 * k, d_qdo + 128 * k, id_vk, k > 0 ? 1u : acc_i);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        if (leader && kk * 16 < kw) umma_bf16(tmem + dq_col(nbkv, c.i), d_dsk + 2 * kk, d_k + 128 * kk, id_q, kk > 0 ? 1u : acc_j);
      if (leader) {
        if (c.j == KB - 1) umma_commit(bar(7 + qs));     // last key block: this query tile's slot may be reloaded
        if (c.i == QT - 1) {
          umma_commit(bar(2 + ks));
          umma_commit(bar(BAR_DKV_FULL + c.kvit % nbkv));
          if (c.j == KB - 1) umma_commit(bar(BAR_DQ_FULL));
        }
      }
      __syncwarp();
      if (a.dbg != nullptr && (int)blockIdx.x == (a.dbg_cta & 0xffff) && leader && c.n < 64) a.dbg[c.n * 16 + 2] = clock64();
      c = nx;
      nx.advance(QT, KB, nslots);
    }
  } else if (warp < EP_WARP0) {
    // ---------------- softmax warps: P = exp2(S*c - lse), dS = P o (dP - delta) -> bf16 staging ----------------
    const int q = warp & 3, hf = (warp - SM_WARP0) >> 2, row = q * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const float c_exp = a.scale * LOG2E_F;
    BwdCursor c = {slot, 0, 0, 0, 0, 0, 0};
    float lse2[3] = {INFINITY, INFINITY, INFINITY}, delta[3] = {0.f, 0.f, 0.f};
    while (c.b < NB) {
      if (c.j == 0 && c.i == 0) {
        // statistics of this pair, produced by the epilogue warps one pair ahead
        named_bar_sync(2 + (c.pair & 1), stat_threads);
        const float2* st = stats + (c.pair & 1) * 3 * 128;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float2 v = st[i * 128 + row];
          lse2[i] = v.x, delta[i] = v.y;
        }
      }
      const int keys_valid = min(BKW, Nk - c.j * BKW);
      const int kw = round16(keys_valid);
      const int rows_valid = min(128, Nq - c.i * 128);
      const bool active = q * 32 < round16(rows_valid);
      const float my_lse = c.i == 0 ? lse2[0] : (c.i == 1 ? lse2[1] : lse2[2]);
      const float my_delta = c.i == 0 ? delta[0] : (c.i == 1 ? delta[1] : delta[2]);
      uint8_t* Ps = smem + B_STG + (c.n & 1) * 2 * TILE16K;
      uint8_t* dSs = Ps + TILE16K;
      const bool stamp = a.dbg != nullptr && (int)blockIdx.x == (a.dbg_cta & 0xffff) && warp == SM_WARP0 + 2 && lane == 0 && c.n < 64;   // warp 4: lane quarter 0
      if (stamp) a.dbg[c.n * 16 + 3] = clock64();
      mbar_wait(bar(BAR_SDP), (uint32_t)c.n & 1u);
      tc_fence_after();
      if (stamp) a.dbg[c.n * 16 + 4] = clock64();
      if (active) {
        if (kw == BKW) {
          // full block: each warp of the lane quarter takes 32 columns; both TMEM loads are in flight before the first use
          const int col = hf * 32;
          float s[32], dp[32];
          tmem_ld32(tlane + C_S + col, s);
          tmem_ld32(tlane + C_DP + col, dp);
          tmem_ld_wait();
          if (col + 32 <= keys_valid) {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              s[t] = ex2(fmaf(s[t], c_exp, -my_lse));
              dp[t] = s[t] * (dp[t] - my_delta);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const bool ok = col + t < keys_valid;
              const float p = ok ? ex2(fmaf(s[t], c_exp, -my_lse)) : 0.f;
              s[t] = p;
              dp[t] = ok ? p * (dp[t] - my_delta) : 0.f;
            }
          }
          stage_bf16_32(Ps, row, col, s);
          stage_bf16_32(dSs, row, col, dp);
        } else {
          // partial block: the two warps split the kw columns in 16-column units: [0, kh) and [kh, kw)
          const int kh = round16(kw >> 1);
          const int c0 = hf == 0 ? 0 : kh, c1 = hf == 0 ? kh : kw;
          for (int col = c0; col < c1; col += 16) {
            float s[16], dp[16];
            tmem_ld16(tlane + C_S + col, s);
            tmem_ld16(tlane + C_DP + col, dp);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              const bool ok = col + t < keys_valid;
              const float p = ok ? ex2(fmaf(s[t], c_exp, -my_lse)) : 0.f;
              s[t] = p;
              dp[t] = ok ? p * (dp[t] - my_delta) : 0.f;
            }
            stage_bf16_16(Ps, row, col, s);
            stage_bf16_16(dSs, row, col, dp);
          }
        }
      }
      if (stamp) a.dbg[c.n * 16 + 5] = clock64();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_PDS));
      if (stamp) a.dbg[c.n * 16 + 6] = clock64();
      c.advance(QT, KB, nslots);
    }
  } else if (warp < SIDE_WARP0) {
    // ---------------- epilogue warps: statistics of the next pair, dK / dV / dQ -> global, bias-gradient column sums ----------------
    const int q = warp & 3, row = q * 32 + lane, et = threadIdx.x - EP_WARP0 * 32;   // et = 0..127
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const long ldq = 3L * HD;
    if (a.segs.count == 0) {
      zero_masked(a.dqkv, ldq, (long)a.B * N, a.Hk * D, (a.H - a.Hk) * D, 3, HD, et, 128);
    } else {
      for (int i = 0, b0 = 0; i < a.segs.count; b0 = a.segs.b_end[i], ++i)
        if (a.segs.hk[i] > 0)
          zero_masked(a.dqkv + (long)b0 * N * ldq, ldq, (long)(a.segs.b_end[i] - b0) * N, a.segs.hk[i] * D, (a.H - a.segs.hk[i]) * D, 3, HD, et, 128);
    }
    // lse (log2 units) and delta = dO . O of sample b's query rows -> stats buffer; rows >= N get lse = +inf, so P = exp2(S*c - inf) = 0
    // and dS = 0 there (S and dP are exact zeros on those rows: TMA zero fill)
    auto make_stats = [&](int b, int buf) {
      float2* st = stats + buf * 3 * 128;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float2 v = make_float2(INFINITY, 0.f);
        const int r = i * 128 + et;
        if (i < QT && r < Nq) {
          v.x = a.lse[((long)b * a.H + h) * N + r] * LOG2E_F;
          const bf16* op = a.o + ((long)b * N + r) * HD + h * D;
          const bf16* dp = a.d_o + ((long)b * N + r) * HD + h * D;
          // all (up to) 16 row loads are issued before the first use: a loop over the runtime head dim would serialise their latencies
          uint4 ov[8], dv[8];
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            ov[c8] = make_uint4(0u, 0u, 0u, 0u), dv[c8] = make_uint4(0u, 0u, 0u, 0u);
            if (c8 * 8 < D) {
              ov[c8] = *reinterpret_cast<const uint4*>(op + c8 * 8);
              dv[c8] = *reinterpret_cast<const uint4*>(dp + c8 * 8);
            }
          }
          float acc = 0.f;
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const uint32_t ow[4] = {ov[c8].x, ov[c8].y, ov[c8].z, ov[c8].w}, dw[4] = {dv[c8].x, dv[c8].y, dv[c8].z, dv[c8].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[t]));
              const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dw[t]));
              acc += x.x * y.x + x.y * y.y;
            }
          }
          v.y = acc;
        }
        st[i * 128 + et] = v;
      }
      // hand-over: the 8 softmax warps bar.sync on the same id; buffer `buf` was last read two pairs ago (program order of the
      // softmax warps guarantees they passed that read long before they can reach this barrier again)
      asm volatile("bar.arrive %0, %1;" ::"r"(2 + buf), "r"(stat_threads) : "memory");
    };
    int nkv = 0, pair = 0;
    const bool pf_l2 = (a.dbg_cta >> 16) == 0;
    const bool is_v = q < 2;          // lanes 0..63 hold dV (columns 64..127), lanes 64..127 hold dK (columns 0..63)
    const int krow = row & 63;
    const float kv_mul = is_v ? 1.0f : a.scale;
    if (slot < NB) make_stats(sample_of(a.segs, h, slot), 0);
    for (int v = slot; v < NB; v += nslots, ++pair) {
      const int b = sample_of(a.segs, h, v);
      // statistics (and side paths) of the NEXT pair while the MMAs of this one run (moving the statistics behind the first accumulator
      // drain was measured: 4-15 % slower per launch -- with one key block per pair it delays the dQ drain the next pair's MMAs wait for).
      if (v + nslots < NB) make_stats(sample_of(a.segs, h, v + nslots), (pair + 1) & 1);
      if (v + 2 * nslots < NB && pf_l2) {
        // the O / dO rows of the pair after the next one -> L2 (their use above is bound by the latency of three dependent row loads)
        const int b2 = sample_of(a.segs, h, v + 2 * nslots);
        for (int r = et; r < Nq; r += 128) {
          const long off = ((long)b2 * N + r) * HD + h * D;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.o + off));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.d_o + off));
        }
      }
      const float* sbc = sd + (pair & 1) * SD_BUF;      // this pair's odd-token side data (written by the side warps)
      float csum[2][32];              // this thread's rows, summed over the key blocks of the pair: the butterfly runs once per pair
#pragma unroll
      for (int t = 0; t < 32; ++t) csum[0][t] = 0.f, csum[1][t] = 0.f;
      for (int j = 0; j < KB; ++j, ++nkv) {
        const bool estamp = a.dbg != nullptr && (int)blockIdx.x == (a.dbg_cta & 0xffff) && et == 0 && nkv < 64;
        if (estamp) a.dbg[nkv * 16 + 8] = clock64();
        mbar_wait(bar(BAR_DKV_FULL + nkv % nbkv), (uint32_t)(nkv / nbkv) & 1u);
        if (estamp) a.dbg[nkv * 16 + 9] = clock64();
        if (odd) mbar_wait(bar(BAR_SIDE + (nkv & 3)), (uint32_t)(nkv >> 2) & 1u);     // p_tj, ds_tj of this block (and, from block 0 on, ds_it)
        tc_fence_after();
        if (estamp) a.dbg[nkv * 16 + 10] = clock64();
        const int keys_valid = min(BKW, Nk - j * BKW);
        if ((q & 1) * 32 < keys_valid) {
          const bool kvalid = krow < keys_valid;
          bf16* dst = a.dqkv + ((long)b * N + j * BKW + krow) * ldq + (is_v ? 2 * HD : HD) + h * D;
          const float w_t = oq ? sbc[(is_v ? SD_PB : SD_DSB) + j * BKW + krow] : 0.f;      // p_tj (dV) / ds_tj (dK) of the odd query
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float v[32];
            tmem_ld32(tlane + dvk_col(nbkv, nkv) + (is_v ? 64 : 0) + half * 32, v);
            tmem_ld_wait();
            if (kvalid) {
              if (oq) axpy32(v, w_t, sbc + (is_v ? SD_DOT : SD_QT) + half * 32);
#pragma unroll
              for (int t = 0; t < 32; ++t) {
                v[t] *= kv_mul;
                csum[half][t] += v[t];
              }
              store_bf16_n(dst + half * 32, v, D - half * 32);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_DKV_EMPTY + nkv % nbkv));
        if (estamp) a.dbg[nkv * 16 + 11] = clock64();
      }
      if (a.dbias != nullptr) {
#pragma unroll
        for (int half = 0; half < 2; ++half) atomicAdd(cs + (is_v ? 128 : 64) + half * 32 + lane, butterfly_colsum(csum[half], lane));
      }
      // dQ: all query tiles are complete after the last key block
      mbar_wait(bar(BAR_DQ_FULL), (uint32_t)pair & 1u);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 32; ++t) csum[0][t] = 0.f, csum[1][t] = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (i < QT) {
          const int rows_valid = min(128, Nq - i * 128);
          if (q * 32 < rows_valid) {
            const float w_t = ok ? sbc[SD_DSA + i * 128 + row] : 0.f;        // ds_it of the odd key
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float v[32];
              tmem_ld32(tlane + dq_col(nbkv, i) + half * 32, v);
              tmem_ld_wait();
              if (row < rows_valid) {
                if (ok) axpy32(v, w_t, sbc + SD_KT + half * 32);
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                  v[t] *= a.scale;
                  csum[half][t] += v[t];
                }
                store_bf16_n(a.dqkv + ((long)b * N + i * 128 + row) * ldq + h * D + half * 32, v, D - half * 32);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_DQ_EMPTY));
      if (a.dbias != nullptr) {
#pragma unroll
        for (int half = 0; half < 2; ++half) atomicAdd(cs + half * 32 + lane, butterfly_colsum(csum[half], lane));
      }
    }
    if (a.dbias != nullptr) {
      named_bar_sync(1, 128 + (odd ? SIDE_THREADS : 0));       // the side warps' contributions to cs are complete when they arrive
      for (int t = et; t < 192; t += 128)
        if ((t & 63) < D) atomicAdd(a.dbias + (long)(t >> 6) * HD + h * D + (t & 63), cs[t]);
    }
  } else if (odd) {
    // ---------------- side warps: the odd token t = N - 1 on the CUDA cores, from the pair's own TMA tiles in shared memory ----------------
    // Per pair:  key block j (as soon as its K / V tile has landed): p_tj, ds_tj of its 64 keys (thread = key) -> side data for the
    // [dV|dK] drain, and this block's share of dQ_t = sum_j ds_tj k_j (lane = two columns, the warps split the keys);  query tile i:
    // p_it, ds_it of its rows (thread = row) -> side data for the dQ drain, dK_t = sum_i ds_it q_i, dV_t = sum_i p_it dO_i.
    const int sid = threadIdx.x - SIDE_WARP0 * 32, sw = sid >> 5, c2 = 2 * lane;
    const uint32_t coff = (uint32_t)(((lane >> 2) << 4) | ((lane & 3) << 2));   // byte offset of columns 2l, 2l + 1 in an unswizzled row
    const float c_exp = a.scale * LOG2E_F;
    const long ldq = 3L * HD;
    float* s_pa = sd + SD_PA;
    float* s_acc = sd + SD_ACC;        // [3][64] dQ_t, dK_t, dV_t before scaling
    int kvit = 0, pair = 0;
    for (int v = slot; v < NB; v += nslots, ++pair) {
      const int b = sample_of(a.segs, h, v);
      float* sb = sd + (pair & 1) * SD_BUF;
      // The statistics of this pair: their producers (the epilogue warps) wrote them after the drains of the pair before the previous
      // one, the last readers of this side-data buffer.
      named_bar_sync(2 + (pair & 1), stat_threads);
      const float2* st = stats + (pair & 1) * 3 * 128;
      const long trow = (long)b * N + (N - 1);
      for (int t = sid; t < 192; t += SIDE_THREADS) s_acc[t] = 0.f;
      if (sw == 0) {
        float2 qv = make_float2(0.f, 0.f), kv = qv, vv = qv, ov = qv, dv = qv;
        if (c2 < D) {
          const bf16* qp = a.qkv + trow * ldq + h * D + c2;
          qv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qp));
          kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qp + HD));
          vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qp + 2 * HD));
          ov = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a.o + trow * HD + h * D + c2));
          dv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a.d_o + trow * HD + h * D + c2));
        }
        sb[SD_QT + c2] = qv.x, sb[SD_QT + c2 + 1] = qv.y;
        sb[SD_KT + c2] = kv.x, sb[SD_KT + c2 + 1] = kv.y;
        sb[SD_DOT + c2] = dv.x, sb[SD_DOT + c2 + 1] = dv.y;
        sb[SD_VT + c2] = vv.x, sb[SD_VT + c2 + 1] = vv.y;
        const float dl = warp_sum(ov.x * dv.x + ov.y * dv.y);
        if (lane == 0) {
          sb[SD_SCAL] = a.lse[((long)b * a.H + h) * N + (N - 1)] * LOG2E_F;
          sb[SD_SCAL + 1] = dl;
        }
      }
      named_bar_sync(4, SIDE_THREADS);
      const float lse_t = sb[SD_SCAL], delta_t = sb[SD_SCAL + 1];
      float2 aq = make_float2(0.f, 0.f), ak = aq, av = aq;      // columns 2l, 2l + 1 of dQ_t, dK_t, dV_t: this warp's share of the rows
      for (int j = 0; j < KB; ++j, ++kvit) {
        const bool sstamp = a.dbg != nullptr && (int)blockIdx.x == (a.dbg_cta & 0xffff) && sid == 0 && kvit < 64;
        if (sstamp) a.dbg[kvit * 16 + 12] = clock64();
        // The K / V ring paces these warps in every mode (they stay within three key blocks of the MMAs: the four BAR_SIDE phases cannot alias).
        const int ks = kvit & 1, keys = min(BKW, Nk - j * BKW);
        mbar_wait(bar(ks), ((uint32_t)kvit >> 1) & 1u);
        if (sstamp) a.dbg[kvit * 16 + 13] = clock64();
        if (oq) {
          // the odd QUERY against the keys of block j
          const uint8_t* kt = smem + B_KV + ks * 2 * BOX8K;      // K rows of the block; its V rows are BOX8K further
          if (sid < keys) {
            const float4* q4 = reinterpret_cast<const float4*>(sb + SD_QT);
            const float4* d4 = reinterpret_cast<const float4*>(sb + SD_DOT);
            float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;
#pragma unroll
            for (int c8 = 0; c8 < 8; c8 += 2) {
              s0 += dot8(*reinterpret_cast<const uint4*>(kt + swz(sid, c8)), q4[2 * c8], q4[2 * c8 + 1]);
              s1 += dot8(*reinterpret_cast<const uint4*>(kt + swz(sid, c8 + 1)), q4[2 * c8 + 2], q4[2 * c8 + 3]);
              p0 += dot8(*reinterpret_cast<const uint4*>(kt + BOX8K + swz(sid, c8)), d4[2 * c8], d4[2 * c8 + 1]);
              p1 += dot8(*reinterpret_cast<const uint4*>(kt + BOX8K + swz(sid, c8 + 1)), d4[2 * c8 + 2], d4[2 * c8 + 3]);
            }
            const float p_b = ex2(fmaf(s0 + s1, c_exp, -lse_t));
            sb[SD_PB + j * BKW + sid] = p_b, sb[SD_DSB + j * BKW + sid] = p_b * ((p0 + p1) - delta_t);
          }
          named_bar_sync(4, SIDE_THREADS);
          const int k1 = min(keys, sw * 32 + 32);
#pragma unroll 8
          for (int key = sw * 32; key < k1; ++key) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(kt + key * 128 + (coff ^ ((uint32_t)(key & 7) << 4)));
            const float ds = sb[SD_DSB + j * BKW + key];
            aq.x = fmaf(ds, bf_lo(w), aq.x), aq.y = fmaf(ds, bf_hi(w), aq.y);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(2 + ks));                 // this warp is done with the K / V slot
        if (ok && j == 0) {
          // the odd KEY against the query rows (the Q / dO tiles of a pair follow its first K / V block)
          for (int i = 0; i < QT; ++i) {
            const int tl = pair * QT + i, qs = tl % 3;
            mbar_wait(bar(4 + qs), ((uint32_t)tl / 3) & 1u);
            const uint8_t* qt = smem + B_QDO + qs * 2 * TILE16K;   // Q tile; the dO tile is TILE16K further
            const int rows = min(128, Nq - i * 128);
            const float4* k4 = reinterpret_cast<const float4*>(sb + SD_KT);
            const float4* v4 = reinterpret_cast<const float4*>(sb + SD_VT);
            for (int r = sid; r < rows; r += SIDE_THREADS) {
              float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;
#pragma unroll
              for (int c8 = 0; c8 < 8; c8 += 2) {
                s0 += dot8(*reinterpret_cast<const uint4*>(qt + swz(r, c8)), k4[2 * c8], k4[2 * c8 + 1]);
                s1 += dot8(*reinterpret_cast<const uint4*>(qt + swz(r, c8 + 1)), k4[2 * c8 + 2], k4[2 * c8 + 3]);
                p0 += dot8(*reinterpret_cast<const uint4*>(qt + TILE16K + swz(r, c8)), v4[2 * c8], v4[2 * c8 + 1]);
                p1 += dot8(*reinterpret_cast<const uint4*>(qt + TILE16K + swz(r, c8 + 1)), v4[2 * c8 + 2], v4[2 * c8 + 3]);
              }
              const float2 sv = st[i * 128 + r];
              const float p_a = ex2(fmaf(s0 + s1, c_exp, -sv.x));
              s_pa[i * 128 + r] = p_a, sb[SD_DSA + i * 128 + r] = p_a * ((p0 + p1) - sv.y);
            }
            named_bar_sync(4, SIDE_THREADS);
            for (int r0 = sw * 32; r0 < rows; r0 += 64) {
              const int r1 = min(rows, r0 + 32);
#pragma unroll 8
              for (int r = r0; r < r1; ++r) {
                const uint32_t o = (uint32_t)(r * 128) + (coff ^ ((uint32_t)(r & 7) << 4));
                const uint32_t wq = *reinterpret_cast<const uint32_t*>(qt + o), wd = *reinterpret_cast<const uint32_t*>(qt + TILE16K + o);
                const float ds = sb[SD_DSA + i * 128 + r], pa = s_pa[i * 128 + r];
                ak.x = fmaf(ds, bf_lo(wq), ak.x), ak.y = fmaf(ds, bf_hi(wq), ak.y);
                av.x = fmaf(pa, bf_lo(wd), av.x), av.y = fmaf(pa, bf_hi(wd), av.y);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(7 + qs));             // this warp is done with the Q / dO slot
          }
        }
        // side data of key block j complete (block 0: also ds_it of every query row, q_t, k_t, dO_t): the drains may use it
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_SIDE + (kvit & 3)));
        if (sstamp) a.dbg[kvit * 16 + 14] = clock64();
      }
      if (oq && ok && sw == 0) {     // the (t, t) element: on no tile when both sides run here
        const float qx = sb[SD_QT + c2], qy = sb[SD_QT + c2 + 1], kx = sb[SD_KT + c2], ky = sb[SD_KT + c2 + 1];
        const float dx = sb[SD_DOT + c2], dy = sb[SD_DOT + c2 + 1], vx = sb[SD_VT + c2], vy = sb[SD_VT + c2 + 1];
        const float s_tt = warp_sum(qx * kx + qy * ky), dp_tt = warp_sum(dx * vx + dy * vy);
        const float p_tt = ex2(fmaf(s_tt, c_exp, -lse_t)), ds_tt = p_tt * (dp_tt - delta_t);
        aq.x = fmaf(ds_tt, kx, aq.x), aq.y = fmaf(ds_tt, ky, aq.y);
        ak.x = fmaf(ds_tt, qx, ak.x), ak.y = fmaf(ds_tt, qy, ak.y);
        av.x = fmaf(p_tt, dx, av.x), av.y = fmaf(p_tt, dy, av.y);
      }
      atomicAdd(s_acc + c2, aq.x), atomicAdd(s_acc + c2 + 1, aq.y);
      atomicAdd(s_acc + 64 + c2, ak.x), atomicAdd(s_acc + 64 + c2 + 1, ak.y);
      atomicAdd(s_acc + 128 + c2, av.x), atomicAdd(s_acc + 128 + c2 + 1, av.y);
      named_bar_sync(4, SIDE_THREADS);
      // the gradient rows of token t that no tile produces: dQ_t (odd query), dK_t and dV_t (odd key)
      for (int t = sid; t < 192; t += SIDE_THREADS) {
        const int col = t & 63, which = t >> 6;       // 0 dQ_t, 1 dK_t, 2 dV_t
        if (col < D && (which == 0 ? oq : ok)) {
          const float val = s_acc[t] * (which == 2 ? 1.0f : a.scale);
          a.dqkv[trow * ldq + which * HD + h * D + col] = __float2bfloat16(val);
          if (a.dbias != nullptr) atomicAdd(cs + t, val);
        }
      }
    }
    if (a.dbias != nullptr) asm volatile("bar.arrive %0, %1;" ::"r"(1), "r"(128 + SIDE_THREADS) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (a.dbg != nullptr && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg[1024 + 2 * blockIdx.x + 1] = t;
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    a.dbg[1024 + 2 * 160 + blockIdx.x] = smid;
  }
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int set_smem(const void* fn, int bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%d) failed: %s", what, bytes, cudaGetErrorString(e));
    return VSX_ERR_CUDA;
  }
  return VSX_OK;
}

// masked heads: their slices are defined to be zero (consumers contract over the full feature width in places)
int zero_cols(void* base, long ld_elems, long rows, long col0, long ncols, cudaStream_t st, const char* what) {
  if (ncols <= 0 || rows <= 0) return VSX_OK;
  cudaError_t e = cudaMemset2DAsync(static_cast<uint8_t*>(base) + col0 * 2, (size_t)ld_elems * 2, 0, (size_t)ncols * 2, (size_t)rows, st);
  if (e != cudaSuccess) {
    set_error("%s: cudaMemset2DAsync failed: %s", what, cudaGetErrorString(e));
    return VSX_ERR_CUDA;
  }
  return VSX_OK;
}

}  // namespace

// kernel-side copy of a segment list (plus the forward work-list prefix sums); sg == nullptr: uniform
static void fill_segs(AttnSegs& t, const vsx_sample_segments* sg) {
  memset(&t, 0, sizeof(t));
  if (sg == nullptr) return;
  t.count = sg->count;
  int w = 0, b0 = 0;
  for (int i = 0; i < sg->count; ++i) {
    t.b_end[i] = sg->sample_end[i], t.hk[i] = sg->heads_keep[i];
    w += (sg->sample_end[i] - b0) * sg->heads_keep[i];
    t.w_end[i] = w;
    b0 = sg->sample_end[i];
  }
}

static long long* g_attn_dbg = nullptr;
void attn_set_debug(long long* p) { g_attn_dbg = p; }
void attn_set_odd_modes(int f, int b, int s);

// Which launches run the last token on the side paths.  Forward: as a query when that saves a query tile (N = 128 k + 1 > 128: 97 -> 75 us
// at N = 257, B = 256, 4 heads).  Backward, N = 64 k + 1 > 128: both sides (8 instead of 15 blocks per pair: 189 -> 162 us; the query side
// alone, 10 blocks: 176 us).  Backward at N = 65: off (measured: 105 us without, 108 .. 131 us with either side -- one block per pair leaves
// the side warps no time to hide in).  VSX_ATTN_ODD = "f,b,s" overrides (development / A-B measurements): f = forward 0 / 1; b, s =
// backward mode for N > 128 / N <= 128 (bit 0 query side, bit 1 key side).  Numbers: tools/attn_bench.py, profiles/r2_attention.md.
static int g_odd_force[3] = {-1, -1, -1};      // vsx_attn_odd_token_modes
static int odd_mode(int which, int dflt) {
  static int m[3] = {-1, -1, -1};
  if (m[0] == -1) {
    m[0] = -2, m[1] = -2, m[2] = -2;
    const char* e = getenv("VSX_ATTN_ODD");
    if (e != nullptr) sscanf(e, "%d,%d,%d", &m[0], &m[1], &m[2]);
  }
  if (g_odd_force[which] >= 0) return g_odd_force[which] & 3;
  return m[which] >= 0 ? m[which] & 3 : dflt;
}

void attn_set_odd_modes(int f, int b, int s) { g_odd_force[0] = f, g_odd_force[1] = b, g_odd_force[2] = s; }

bool attn_tc_supported(int N, int D) { return (D == 64 || D == 48 || D == 32) && N >= 1 && N <= ATT_MAX_N; }

int attn_fwd_tc(const void* qkv, void* o, float* lse, int B, int N, int H, int D, int Hk, float scale, cudaStream_t st, const vsx_sample_segments* sg) {
  const long HD = (long)H * D;
  int rc = VSX_OK;
  if (Hk == 0) return zero_cols(o, HD, (long)B * N, 0, HD, st, "vsx_attn_fwd");
  AttnMaps maps;
  memset(&maps, 0, sizeof(maps));
  if ((rc = make_tmap_heads(&maps.q, qkv, D, 3 * H, N, B, 128))) return rc;
  if ((rc = make_tmap_heads(&maps.kv, qkv, D, 3 * H, N, B, KV_BOX))) return rc;
  static bool configured = false;
  if (!configured) {
    if ((rc = set_smem((const void*)attn_fwd_tc_kernel, F_SMEM, "vsx_attn_fwd"))) return rc;
    configured = true;
  }
  AttnArgs a;
  a.B = B, a.N = N, a.H = H, a.Hk = Hk, a.D = D, a.scale = scale, a.o = (bf16*)o, a.d_o = nullptr, a.lse = lse, a.dqkv = nullptr, a.dbias = nullptr, a.dbg = nullptr;
  a.dbg_cta = 0;
  a.qkv = (const bf16*)qkv;
  a.odd = (N > 128 && N % 128 == 1) ? (odd_mode(0, 1) & 1) : 0;
  fill_segs(a.segs, sg);
  const int total = sg != nullptr ? a.segs.w_end[a.segs.count - 1] : B * Hk;
  if (total == 0) return VSX_OK;
  const int grid = total < num_sms() ? total : num_sms();
  launch_pdl(attn_fwd_tc_kernel, dim3(grid), dim3(a.odd ? ATT_THREADS_ODD : ATT_THREADS), F_SMEM, st, maps, a);
  return check_launch("vsx_attn_fwd");
}

int attn_bwd_tc(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int B, int N, int H, int D, int Hk, float scale,
                float* dbias, cudaStream_t st, const vsx_sample_segments* sg) {
  const long HD = (long)H * D;
  int rc = VSX_OK;
  if (Hk == 0) return zero_cols(dqkv, 3 * HD, (long)B * N, 0, 3 * HD, st, "vsx_attn_bwd");
  AttnMaps maps;
  if ((rc = make_tmap_heads(&maps.q, qkv, D, 3 * H, N, B, 128))) return rc;
  if ((rc = make_tmap_heads(&maps.kv, qkv, D, 3 * H, N, B, BKW))) return rc;
  if ((rc = make_tmap_heads(&maps.d_o, d_o, D, H, N, B, 128))) return rc;
  static bool configured = false;
  if (!configured) {
    if ((rc = set_smem((const void*)attn_bwd_tc_kernel, B_SMEM, "vsx_attn_bwd"))) return rc;
    configured = true;
  }
  AttnArgs a;
  a.B = B, a.N = N, a.H = H, a.Hk = Hk, a.D = D, a.scale = scale, a.o = (bf16*)const_cast<void*>(o), a.d_o = (const bf16*)d_o,
  a.lse = const_cast<float*>(lse), a.dqkv = (bf16*)dqkv, a.dbias = dbias, a.dbg = g_attn_dbg;
  a.dbg_cta = (g_attn_dbg != nullptr && getenv("VSX_ATTN_DBG_CTA") != nullptr) ? atoi(getenv("VSX_ATTN_DBG_CTA")) : 0;
  static const int no_pf = getenv("VSX_ATTN_NO_PF") != nullptr ? 1 : 0;      // development: A-B of the L2 prefetch
  a.dbg_cta |= no_pf << 16;
  a.qkv = (const bf16*)qkv;
  a.odd = (N > 64 && N % 64 == 1) ? (N > 128 ? odd_mode(1, 3) : odd_mode(2, 0)) : 0;
  fill_segs(a.segs, sg);
  int per_head = num_sms() / Hk;
  if (per_head < 1) per_head = 1;
  if (per_head > B) per_head = B;
  int grid = per_head * Hk;
  if (sg != nullptr) {
    // CTAs per head in proportion to the number of samples that keep it (at least one, at most one per sample)
    VSX_REQUIRE(Hk <= 16, "vsx_attn_bwd_segs: at most 16 heads (got %d)", Hk);
    int cnt[16], total = 0;
    for (int hh = 0; hh < Hk; ++hh) {
      cnt[hh] = 0;
      for (int i = 0, b0 = 0; i < sg->count; b0 = sg->sample_end[i], ++i)
        if (sg->heads_keep[i] > hh) cnt[hh] += sg->sample_end[i] - b0;
      total += cnt[hh];
    }
    const int sms = num_sms() > Hk ? num_sms() : Hk;
    int end = 0;
    for (int hh = 0; hh < Hk; ++hh) {
      int c = total > 0 ? (int)((long)sms * cnt[hh] / total) : 1;
      c = c < 1 ? 1 : (c > cnt[hh] && cnt[hh] > 0 ? cnt[hh] : c);
      end += c;
      a.segs.cta_end[hh] = end;
    }
    grid = end;
  }
  launch_pdl(attn_bwd_tc_kernel, dim3(grid), dim3(a.odd ? BWD_THREADS_ODD : BWD_THREADS), B_SMEM, st, maps, a);
  return check_launch("vsx_attn_bwd");
}

}  // namespace vsx
