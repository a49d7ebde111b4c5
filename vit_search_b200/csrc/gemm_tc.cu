// tcgen05 / TMEM / TMA GEMM with fused epilogues for the ViT-Res super-network hot path (sm_100a).
//
//   D[M,N] = epilogue( sum_t A_t[M,K] * B_t[N,K]^T ),  bf16 operands, fp32 accumulation in tensor memory.
//
// One CTA computes one 128x128 output tile.  Warp roles (192 threads):
//   warp 0      TMA producer : cp.async.bulk.tensor loads of 128x64 A / B boxes (128B swizzle) into a
//                              STAGES-deep shared-memory ring, completion on "full" mbarriers
//   warp 1      MMA issuer   : one elected thread issues tcgen05.mma (128x128x16, cta_group::1); tcgen05.commit
//                              releases ring slots ("empty" mbarriers) and finally signals the epilogue
//   warps 2..5  epilogue     : tcgen05.ld of the fp32 accumulator (one TMEM lane = one output row per thread), fused
//                              bias / GELU / drop-path scale / prefix mask / residual / GELU'; the tile is staged in the
//                              idle operand ring (128B-swizzled) and leaves by TMA store, or TMA reduce-add for the
//                              split-K weight gradients; residual / pre-activation tiles arrive by TMA load
// Two CTAs fit per SM (96 KB shared memory, 128 TMEM columns each), so one tile's epilogue overlaps the other's
// main loop.  Both operand layouts are supported through the shared-memory descriptors (K-major for the forward
// pass, MN-major for dgrad / wgrad), so no transposed copies of activations or weights are ever made.
//
// Replaces: nn.Linear forward/backward at nets/supernet_blocks.py:38,50,102,118 and the elementwise tails at
// :39 (GELU), :238-253 (mask, residual), nets/drop.py:11-26 (drop-path scale); see include/vsx.h.
#include <stdlib.h>

#include "common.cuh"

namespace vsx {

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int UMMA_K = 16;
constexpr int SUB_A = BM * BK * 2;      // one 128-row A sub-tile of a k block
constexpr int STAGE_B = BN * BK * 2;
// MT = 128-row sub-tiles per CTA tile.  MT = 2 computes a 256 x 128 tile from ONE B box per k block (two accumulators share it):
// 3/4 of the L2 -> shared-memory bytes per flop of two independent 128 x 128 tiles.  These GEMMs have K <= 1536 and run at the
// L2 -> SM fabric limit (~45 B/clk/SM) long before the tensor pipe saturates, so bytes per flop is what sets their speed.
constexpr bool RESID_NBUF2 = false;   // measured: helps K <= 384 (48 -> 37 us), hurts K >= 768 (61 -> 68 us): the ring then is too shallow
// CG = CTAs per tile (cta_group).  CG = 2: a CTA PAIR (one cluster of two SMs) computes a 256 x 256 tile with tcgen05.mma.cta_group::2
// (M = 256, N = 256): each CTA stages its own 128 rows of A and its own 128 columns of B -- 32 KB per k block for 128 x 256 outputs per
// CTA, i.e. 2/3 of the shared-memory fill bytes per flop of the 256 x 128 single-CTA tile and half of its MMA operand reads -- and owns
// the 128 x 256 fp32 accumulator of its rows in its tensor memory (two sets = all 512 columns).
template <int MT, int CG> struct Shape {
  static_assert(CG == 1 || (CG == 2 && MT == 1), "CTA pairs use one 128-row A sub-tile per CTA");
  static constexpr int NT = CG;                      // 128-column accumulator units per CTA
  static constexpr int UNITS = MT * NT;              // 128 x 128 epilogue units per CTA and tile
  static constexpr int STAGE_A = MT * SUB_A;
  static constexpr int STAGE_BYTES = STAGE_A + STAGE_B;
  static constexpr int TILE_M = CG * MT * BM, TILE_N = NT * BN;
  static constexpr int ACC_COLS = UNITS * BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;     // two accumulator sets (double buffered across tiles)
};
constexpr int EPI_WARPS = 8;            // two warps per TMEM lane quarter, each owning half of the tile's columns
constexpr int EPI0 = 3;                 // warp 0 TMA producer, warp 1 MMA issuer, warp 2 store/aux warp, warps 3.. epilogue
constexpr int GEMM_THREADS = (EPI0 + EPI_WARPS) * 32;
constexpr int MAX_STAGES = 5;
constexpr int SMEM_LIMIT = 227 * 1024;
// No alignment slack: the kernel has no static shared memory, so its dynamic shared memory starts right behind the 1 KB the system reserves
// per block, i.e. 1024-byte aligned (declared __align__(1024) and checked at run time).  The 1 KB this frees is what a third operand stage of
// the GELU pair launches and a fifth of the RESIDUAL pair launches were short of.
constexpr int SMEM_MISC_BASE = 256 /*barriers, tmem slot*/;      // + two bias tiles (Plan::SMEM_MISC)

constexpr int COLACC_MAX = 3072;      // widest output whose bias-gradient column sums are accumulated in shared memory
constexpr int MAX_TERMS = 6;
template <int TERMS> struct TmapPackT {
  CUtensorMap a[TERMS];
  CUtensorMap b[TERMS];
  CUtensorMap out, out2, aux;   // epilogue tiles (128B-swizzled boxes of 128 rows x 128 bytes)
};
// Launches of up to 4 problems carry six operand-map pairs per problem (the split-bf16 parity mode contracts six terms); launches of up to
// 16 problems (all segments of a multi-architecture half block at once) carry one pair: single-term bf16 problems only.
template <int G> struct GroupPack { using type = TmapPackT<(G > 4 ? 1 : MAX_TERMS)>; };

static long long* g_gemm_dbg = nullptr;   // development aid: clock stamps of CTA 0 (tools/gemm_timeline.py)
static int g_force_mt = 0;   // 0 = heuristic, 1 / 2 = force the 128- / 256-row CTA tile (tests)
static int g_force_cg = 0;   // 0 = heuristic, 1 = single-CTA tiles only, 2 = CTA pairs (cta_group::2) for every launch (tests, A/B runs)

struct GemmArgs {
  int M, N, K, num_kb, terms, a_mn, b_mn, n_out, split_k;
  const float* bias;
  const float* row_scale;
  float* colsum;
  int rows_per_sample, n_keep;
  int kseg_kb, kseg_stride;   // k-blocks per reduction window and distance between windows (0: one contiguous reduction)
  int m_fastest;              // tile order: 0 = consecutive tiles walk the N tiles of one row block (they share the A tile), 1 = they walk the row blocks of one
                              // column block (they share the B tile).  Development switch VSX_GEMM_TILE_ORDER; default 0.
  long long* dbg;
};

// A launch computes up to G independent problems that share epilogue, output type and tile shape (same kernel instantiation):
// the q / k / v row blocks of a head-masked qkv projection, or all weight gradients of a half block.  Tiles of all problems form
// one list that the persistent CTAs walk, so small problems fill the machine together instead of paying one launch each.
template <int G> struct Group {
  typename GroupPack<G>::type maps[G];
  GemmArgs args[G];
  int tile_end[G];     // exclusive prefix sums of the problems' tile counts
  int count, total;
  int colacc_n;        // > 0: every problem accumulates its fused column sums into ONE target of this width (per-CTA shared-memory partials)
};

// Shared-memory matrix descriptor (tcgen05), 128B swizzle.  K-major: rows of 128 B, 8-row groups 1024 B apart (SBO).
// MN-major: 64-element (128 B) MN atoms, k-rows 128 B apart, 8-row groups SBO = 1024 B, MN atoms LBO = BK*128 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, bool mn_major) {
  const uint64_t lbo = mn_major ? (uint64_t)((BK * 128) >> 4) : 0ull;
  const uint64_t sbo = 1024 >> 4;
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}

// Instruction descriptor, kind::f16: D=f32, A=B=bf16, M=128, N=128.
__device__ __forceinline__ uint32_t make_idesc(bool a_mn, bool b_mn, int m = BM, int n = BN) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------- epilogue staging (shared memory, TMA layout)
// A staged tile is a row of 16 KB boxes: [128 rows][128 bytes], 16-byte chunks XOR-swizzled with (row & 7) -- the
// SWIZZLE_128B layout the output / aux tensor maps use.  bf16: 64 columns per box, fp32: 32 columns per box.
// Thread = tile row; a 512-byte warp access to one chunk column touches each bank group 4 times = the 4-wavefront minimum.
constexpr int BOX_BYTES = 128 * 128;

__device__ __forceinline__ uint32_t box_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// 32 consecutive columns starting at tile column c (multiple of 32) of row `row`
template <typename OutT>
__device__ __forceinline__ void stage_write32(uint8_t* tile, int row, int c, const float (&v)[32]);
template <>
__device__ __forceinline__ void stage_write32<float>(uint8_t* tile, int row, int c, const float (&v)[32]) {
  uint8_t* box = tile + (c / 32) * BOX_BYTES;
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(box + box_off(row, i)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <>
__device__ __forceinline__ void stage_write32<bf16>(uint8_t* tile, int row, int c, const float (&v)[32]) {
  uint8_t* box = tile + (c / 64) * BOX_BYTES;
  const int ch0 = (c % 64) / 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 r;
    r.x = pack_bf16(v[8 * i], v[8 * i + 1]);
    r.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    r.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    r.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(box + box_off(row, ch0 + i)) = r;
  }
}
template <typename OutT>
__device__ __forceinline__ void stage_read32(const uint8_t* tile, int row, int c, float (&v)[32]);
template <>
__device__ __forceinline__ void stage_read32<float>(const uint8_t* tile, int row, int c, float (&v)[32]) {
  const uint8_t* box = tile + (c / 32) * BOX_BYTES;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(box + box_off(row, i));
    v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
  }
}
template <>
__device__ __forceinline__ void stage_read32<bf16>(const uint8_t* tile, int row, int c, float (&v)[32]) {
  const uint8_t* box = tile + (c / 64) * BOX_BYTES;
  const int ch0 = (c % 64) / 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 r = *reinterpret_cast<const uint4*>(box + box_off(row, ch0 + i));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
      v[8 * i + 2 * j] = f.x, v[8 * i + 2 * j + 1] = f.y;
    }
  }
}

// Column sums over the 32 lanes (= 32 tile rows) of a warp: on return lane l holds sum_lanes v[l] (31 shuffles, recursive halving).
__device__ __forceinline__ float warp_colsum32(const float (&in)[32], int lane) {
  float v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = in[k];
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

// Per-instantiation shared-memory plan: NBUF staging tiles (epilogue output / aux input), the rest is the operand ring.
template <int EPI, typename OutT, int MT, int CG, bool CA> struct Plan {
  static constexpr int STAGE_BYTES = Shape<MT, CG>::STAGE_BYTES;
  static constexpr int SMEM_MISC = SMEM_MISC_BASE + 2 * Shape<MT, CG>::TILE_N * 4;
  static constexpr int TILE_BYTES = BM * BN * (int)sizeof(OutT) * (EPI == VSX_EPI_GELU ? 2 : 1);
  // per-CTA accumulator of the fused bias-gradient column sums (flushed once at the end of the persistent loop: thousands of
  // per-tile global atomics on a few hundred addresses serialise in L2 and used to dominate the GELU' dgrad GEMMs)
  // (CA: only the instantiations launched WITH fused column sums reserve it -- 12 KB are the difference between four and five operand
  // stages for the bf16 STORE pair launches, which are bound by operand latency at K <= 256)
  static constexpr int COLACC = (CA && (EPI == VSX_EPI_GELUGRAD || EPI == VSX_EPI_STORE)) ? COLACC_MAX * 4 : 0;
  // two staging tiles when at least three operand stages still fit (the ring has to cover the L2 latency: ~100 KB in flight)
  // GELU (two output tiles per unit) and RESIDUAL (fp32 tile in, fp32 tile out): double buffering the 64 KB staging tile matters
  // more than ring depth (measured: GELU 88 -> 76 us at 65792 x 768 x 224 with two stages + two staging tiles)
  static constexpr int MIN_STAGES2 = (EPI == VSX_EPI_GELU || (EPI == VSX_EPI_RESIDUAL && RESID_NBUF2)) ? 2 : 3;
  static constexpr int NBUF = (2 * TILE_BYTES + MIN_STAGES2 * STAGE_BYTES + SMEM_MISC + COLACC <= SMEM_LIMIT) ? 2 : 1;
  static constexpr int ROOM = (SMEM_LIMIT - SMEM_MISC - COLACC - NBUF * TILE_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = ROOM > MAX_STAGES ? MAX_STAGES : ROOM;
  static constexpr int SMEM = STAGES * STAGE_BYTES + NBUF * TILE_BYTES + SMEM_MISC + COLACC;
  static_assert(STAGES >= 2, "operand ring too small");
};

// Persistent, warp-specialised GEMM.  Each CTA (one per SM) walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... with
// n fastest (CTAs running at the same time share A rows in L2).  Four roles run as independent pipelines over that list:
//   warp 0   producer : TMA operand loads into the ring, running ahead across tile boundaries
//   warp 1   MMA      : tcgen05.mma into TMEM accumulator (tile & 1); tcgen05.commit frees ring slots / publishes the accumulator
//   warp 2   store    : per tile, makes a staging buffer available (TMA-loading the aux tile into it when the epilogue needs one),
//                       later TMA-stores / reduce-adds the staged result; staging is double buffered
//   warps 3+ epilogue : TMEM -> registers -> fused math -> swizzled staging tile
// so the loads of tile i+1, the MMAs of tile i+1 and the stores of tile i-1 overlap the epilogue of tile i.
template <int EPI, typename OutT, int MT, int G, int CG, bool CA>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ Group<G> grp) {
  using P = Plan<EPI, OutT, MT, CG, CA>;
  using S = Shape<MT, CG>;
  constexpr int STAGES = P::STAGES, NBUF = P::NBUF;
  constexpr int STAGE_BYTES = P::STAGE_BYTES, STAGE_A = S::STAGE_A, TMEM_COLS = S::TMEM_COLS;
  constexpr int NT = S::NT, UNITS = S::UNITS, TILE_M = S::TILE_M, TILE_N = S::TILE_N;
  constexpr bool AUX = (EPI == VSX_EPI_RESIDUAL || EPI == VSX_EPI_GELUGRAD);
  constexpr int BOXC = 128 / (int)sizeof(OutT);     // columns per 128-byte box row
  constexpr int NBOX = BN / BOXC;                   // boxes per output tile
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);         // 128B-swizzle atoms need 1024-byte alignment
  if ((base & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("vsx: gemm_tc_kernel: dynamic shared memory is not 1024-byte aligned (0x%x)\n", base);
    __trap();
  }
  uint8_t* smem = smem_raw;
  const uint32_t ring = base, stg = base + STAGES * STAGE_BYTES;
  uint8_t* stg_g = smem + STAGES * STAGE_BYTES;
  const uint32_t bar0 = stg + NBUF * P::TILE_BYTES;
  uint8_t* misc = stg_g + NBUF * P::TILE_BYTES;
  // barriers (8 bytes each): full[S], empty[S], acc_full[2], acc_empty[2], ready[2] (staging writable / aux landed), staged[2]
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 2 + b); };
  auto ready_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 4 + b); };
  auto staged_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 6 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * (2 * MAX_STAGES + 8));
  float* bias_s = reinterpret_cast<float*>(misc + 256);   // [2][TILE_N]
  float* colacc = reinterpret_cast<float*>(misc + P::SMEM_MISC);          // [COLACC_MAX] when P::COLACC != 0
  const bool use_colacc = P::COLACC != 0 && grp.colacc_n > 0;
  long long* const dbg = grp.args[0].dbg;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = grp.total;
  // CTA pairs: both CTAs of a cluster walk the SAME tile list; rank 0 (the leader) issues the MMAs for both
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int tile_first = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_stride = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  // tile t -> (problem, m0, n0, k-block range); identical arithmetic in every role
  auto tile_info = [&](int t, int& pi, int& m0, int& n0, int& kb0, int& nkb) {
    pi = 0;
    int t0 = 0;
    if (G > 1) {
#pragma unroll
      for (int q = 0; q < G - 1; ++q)
        if (q < grp.count - 1 && t >= grp.tile_end[q]) pi = q + 1, t0 = grp.tile_end[q];
    }
    const GemmArgs& g = grp.args[pi];
    const int tl = t - t0;
    const int tiles_m = (g.M + TILE_M - 1) / TILE_M, tiles_n = (g.n_out + TILE_N - 1) / TILE_N;
    const int splits = (EPI == VSX_EPI_ATOMIC) ? g.split_k : 1;
    const int kb_per = (g.num_kb + splits - 1) / splits;
    const int z = tl / (tiles_n * tiles_m);
    const int ni = g.m_fastest ? (tl / tiles_m) % tiles_n : tl % tiles_n, mi = g.m_fastest ? tl % tiles_m : (tl / tiles_n) % tiles_m;
    m0 = mi * TILE_M + cta_rank * (MT * BM), n0 = ni * TILE_N;      // this CTA's rows; the tile's first column
    kb0 = z * kb_per;
    const int kb1 = min(g.num_kb, kb0 + kb_per);
    nkb = (n0 < g.N && kb1 > kb0) ? kb1 - kb0 : 0;
  };

  if (warp == 0 && lane == 0) {
    for (int q = 0; q < grp.count; ++q) {
      for (int t = 0; t < grp.args[q].terms; ++t) {
        tma_prefetch_desc(&grp.maps[q].a[t]);
        tma_prefetch_desc(&grp.maps[q].b[t]);
      }
      tma_prefetch_desc(&grp.maps[q].out);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(acc_full(b), 1);
        mbar_init(acc_empty(b), CG * EPI_WARPS);        // pairs: the epilogue warps of BOTH CTAs release the leader's accumulator barrier
        mbar_init(ready_bar(b), 1);
        mbar_init(staged_bar(b), EPI_WARPS);
      }
      fence_mbar_init();
    }
    __syncwarp();
    if (CG == 2) tmem_alloc_pair(smem_u32(tmem_slot), TMEM_COLS);
    else tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();      // barrier inits of both CTAs are visible cluster-wide before any remote arrive / multicast commit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // barriers, tensor memory and descriptor prefetches are set up: from here on the previous kernel's results are read

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      // pairs: each CTA loads its own A rows and its own 128 B columns into its own ring; all bytes are counted on the LEADER's full
      // barrier (the leader's MMA thread is the only consumer), slots are released to both producers by a multicast commit
      auto load = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
        if (CG == 2) tma_load_2d_pair(dst, m, bar, c0, c1);
        else tma_load_2d(dst, m, bar, c0, c1);
      };
      int it = 0;
      for (int t = tile_first; t < total; t += tile_stride) {
        int pi, m0, n0, kb0, nkb;
        tile_info(t, pi, m0, n0, kb0, nkb);
        const GemmArgs& g = grp.args[pi];
        const auto& maps = grp.maps[pi];
        const int nb0 = n0 + cta_rank * BN;          // this CTA's B columns
        for (int i = 0; i < nkb * g.terms; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const int term = i / nkb, kbi = kb0 + i % nkb;
          const int k0 = g.kseg_kb > 0 ? (kbi / g.kseg_kb) * g.kseg_stride + (kbi % g.kseg_kb) * BK : kbi * BK;
          const uint32_t dA = ring + s * STAGE_BYTES, dB = dA + STAGE_A;
          const uint32_t fb = CG == 2 ? mapa_cluster(full_bar(s), 0) : full_bar(s);
          if (cta_rank == 0) mbar_expect_tx(full_bar(s), CG * STAGE_BYTES);
#pragma unroll
          for (int sub = 0; sub < MT; ++sub) {
            const uint32_t dS = dA + sub * SUB_A;
            const int ms = m0 + sub * BM;
            if (!g.a_mn) {
              load(dS, &maps.a[term], fb, k0, ms);
            } else {
              load(dS, &maps.a[term], fb, ms, k0);
              load(dS + BK * 128, &maps.a[term], fb, ms + 64, k0);
            }
          }
          if (!g.b_mn) {
            load(dB, &maps.b[term], fb, k0, nb0);
          } else {
            load(dB, &maps.b[term], fb, nb0, k0);
            load(dB + BK * 128, &maps.b[term], fb, nb0 + 64, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {
      // ---------------- MMA issuer (pairs: the leader CTA only) ----------------
      auto commit = [&](uint32_t bar) {
        if (CG == 2) umma_commit_pair(bar);
        else umma_commit(bar);
      };
      int it = 0, uses[2] = {0, 0}, j = 0;
      for (int t = tile_first; t < total; t += tile_stride, ++j) {
        int pi, m0, n0, kb0, nkb;
        tile_info(t, pi, m0, n0, kb0, nkb);
        if (nkb == 0) continue;                       // epilogue-only tile: the accumulator is not involved
        const GemmArgs& g = grp.args[pi];
        const uint32_t idesc = CG == 2 ? make_idesc(g.a_mn != 0, g.b_mn != 0, 2 * BM, 2 * BN) : make_idesc(g.a_mn != 0, g.b_mn != 0);
        const uint32_t a_step = g.a_mn ? (UMMA_K * 128) : (UMMA_K * 2);   // bytes per 16-wide k step
        const uint32_t b_step = g.b_mn ? (UMMA_K * 128) : (UMMA_K * 2);
        const int ab = j & 1;
        mbar_wait(acc_empty(ab), ((uint32_t)uses[ab] & 1u) ^ 1u);   // epilogue has drained this accumulator
        ++uses[ab];
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)ab * S::ACC_COLS;
        for (int i = 0; i < nkb * g.terms; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t aA = ring + s * STAGE_BYTES, aB = aA + STAGE_A;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t db = make_smem_desc(aB + k * b_step, g.b_mn != 0);
            if (CG == 2) {
              umma_bf16_pair(acc, make_smem_desc(aA + k * a_step, g.a_mn != 0), db, idesc, (i | k) != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int sub = 0; sub < MT; ++sub)
                umma_bf16(acc + sub * BN, make_smem_desc(aA + sub * SUB_A + k * a_step, g.a_mn != 0), db, idesc, (i | k) != 0 ? 1u : 0u);
            }
          }
          commit(empty_bar(s));        // slot reusable (in both CTAs of a pair) once these MMAs have read it
        }
        commit(acc_full(ab));          // accumulator complete (signalled to the epilogue warps of both CTAs)
        if (dbg != nullptr && blockIdx.x == 0 && j * UNITS < 64) dbg[j * UNITS * 8 + 7] = clock64();
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ---------------- store / aux warp ----------------
      // announce(j): staging buffer j % NBUF is free (its previous store has been read out) -> TMA-load the aux tile into it
      // (completing `ready`), or just arrive on `ready` when the epilogue needs no aux tile.
      // A CTA tile is MT units of 128 rows; units are staged / stored one at a time through the NBUF staging buffers.
      // unit `sub` of a tile: rows m0 + (sub / NT) * 128, columns n0 + (sub % NT) * 128
      auto announce = [&](int t, int sub, int sb) {
        int pi, m0, n0, kb0, nkb;
        tile_info(t, pi, m0, n0, kb0, nkb);
        const GemmArgs& g = grp.args[pi];
        const auto& maps = grp.maps[pi];
        const int nu = n0 + (sub % NT) * BN;
        const int ncols = max(0, min(BN, g.n_out - nu));
        const int nbox = (ncols + BOXC - 1) / BOXC;
        if (AUX) {
          mbar_expect_tx(ready_bar(sb), (uint32_t)nbox * BOX_BYTES);
          for (int bx = 0; bx < nbox; ++bx)
            tma_load_2d(stg + sb * P::TILE_BYTES + bx * BOX_BYTES, &maps.aux, ready_bar(sb), nu + bx * BOXC, m0 + (sub / NT) * BM);
        } else {
          mbar_arrive(ready_bar(sb));
        }
      };
      int un = 0, t = tile_first, sub = 0;
      if (t < total) announce(t, 0, 0);
      while (t < total) {
        const int sb = un % NBUF;
        int tn = t, subn = sub + 1;
        if (subn == UNITS) subn = 0, tn = t + tile_stride;
        if (NBUF == 2 && tn < total) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // store of unit un-1 (same buffer as un+1) has been read
          announce(tn, subn, (un + 1) % NBUF);
        }
        int pi, m0, n0, kb0, nkb;
        tile_info(t, pi, m0, n0, kb0, nkb);
        const GemmArgs& g = grp.args[pi];
        const auto& maps = grp.maps[pi];
        const int ms = m0 + (sub / NT) * BM, nu = n0 + (sub % NT) * BN;
        const int ncols = max(0, min(BN, g.n_out - nu));
        const int nbox = (ncols + BOXC - 1) / BOXC;
        if (dbg != nullptr && blockIdx.x == 0 && un < 64) dbg[un * 8 + 5] = clock64();   // store warp starts waiting for unit un
        mbar_wait(staged_bar(sb), (uint32_t)(un / NBUF) & 1u);               // epilogue has staged unit un
        const uint32_t src = stg + sb * P::TILE_BYTES;
        for (int bx = 0; bx < nbox; ++bx) {
          if (EPI == VSX_EPI_ATOMIC) {
            tma_reduce_add_2d(&maps.out, src + bx * BOX_BYTES, nu + bx * BOXC, ms);
          } else {
            tma_store_2d(&maps.out, src + bx * BOX_BYTES, nu + bx * BOXC, ms);
            if (EPI == VSX_EPI_GELU) tma_store_2d(&maps.out2, src + (NBOX + bx) * BOX_BYTES, nu + bx * BOXC, ms);
          }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (dbg != nullptr && blockIdx.x == 0 && un < 64) dbg[un * 8 + 6] = clock64();   // store of unit un issued
        if (NBUF == 1 && tn < total) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          announce(tn, subn, 0);
        }
        t = tn, sub = subn, ++un;
      }
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");        // shared memory must outlive the last store
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> fused math -> swizzled staging tile ----------------
    const int ew = warp - EPI0;                       // 0..7
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                    // tile row == TMEM lane
    const int et = ew * 32 + lane;                    // 0..255 among the epilogue threads
    // the two warps that share a lane quarter split the columns
    int chalf;
    {
      // warps (EPI0 + i) and (EPI0 + i + 4) have the same (warp & 3): the first takes columns [0,64), the second [64,128)
      chalf = ew >> 2;
    }
    if (use_colacc) {
      for (int i = et; i < ((grp.colacc_n + 3) & ~3); i += EPI_WARPS * 32) colacc[i] = 0.f;
      named_bar_sync(1, EPI_WARPS * 32);
    }
    const uint32_t acc_empty_leader = CG == 2 ? mapa_cluster(acc_empty(0), 0) : 0u;      // + 8 * ab
    float va[32];                                     // one 32-column chunk of this thread's accumulator row
    int uses[2] = {0, 0}, j = 0;
    for (int t = tile_first; t < total; t += tile_stride, ++j) {
      int pi, m0, n0t, kb0, nkb;
      tile_info(t, pi, m0, n0t, kb0, nkb);
      const GemmArgs& g = grp.args[pi];
      const int lim = g.n_keep < g.N ? g.n_keep : g.N;
      const bool has_mma = nkb > 0;
      const int ab = j & 1;
      float* bst = bias_s + ab * TILE_N;
      if (et < TILE_N) bst[et] = (g.bias != nullptr && n0t + et < g.N) ? __ldg(g.bias + n0t + et) : 0.f;
      named_bar_sync(1, EPI_WARPS * 32);                           // bias tile visible
      if (has_mma) {
        mbar_wait(acc_full(ab), (uint32_t)uses[ab] & 1u);
        ++uses[ab];
        tc_fence_after();
      }
      if (dbg != nullptr && blockIdx.x == 0 && warp == EPI0 && lane == 0 && j * UNITS < 64) dbg[j * UNITS * 8 + 4] = clock64();
#pragma unroll 1
      for (int sub = 0; sub < UNITS; ++sub) {
        const int un = j * UNITS + sub, sb = un % NBUF;
        const int ms = m0 + (sub / NT) * BM, m = ms + row;
        const int n0 = n0t + (sub % NT) * BN;
        const int ncols = max(0, min(BN, g.n_out - n0));
        const float* bs = bst + (sub % NT) * BN;
        const bool stamp = dbg != nullptr && blockIdx.x == 0 && warp == EPI0 && lane == 0 && un < 64;
        if (stamp) dbg[un * 8 + 0] = clock64();
        mbar_wait(ready_bar(sb), (uint32_t)(un / NBUF) & 1u);      // staging buffer writable (and aux tile landed)
        if (stamp) dbg[un * 8 + 1] = clock64();
        uint8_t* tile = stg_g + sb * P::TILE_BYTES;
        uint8_t* tile2 = tile + NBOX * BOX_BYTES;
        float scale = 1.0f;
        if (EPI == VSX_EPI_RESIDUAL && g.row_scale != nullptr && m < g.M) scale = __ldg(g.row_scale + m / g.rows_per_sample);
        const uint32_t acc = tmem_base + (uint32_t)(ab * UNITS + sub) * BN + ((uint32_t)(q * 32) << 16);
        // fused math of one 32-column chunk (columns c .. c+31 of the unit) held in v
        auto process = [&](float (&v)[32], const int c) {
          const int n = n0 + c;
          // `full`: all 32 columns of the chunk are real output columns -- the common case runs without per-element predicates
          // (a predicated expensive expression compiles to one branch per element, which serialises the whole chunk)
          const bool full = n + 32 <= g.N;
          if (EPI == VSX_EPI_ATOMIC) {
            stage_write32<float>(tile, row, c, v);
          } else if (EPI == VSX_EPI_STORE) {
            if (full) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) v[jj] += bs[c + jj];
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) v[jj] = (n + jj < g.N) ? v[jj] + bs[c + jj] : 0.f;
            }
            stage_write32<OutT>(tile, row, c, v);
            if (g.colsum != nullptr) {     // bias gradient of the producing Linear: column sums over the valid rows, from registers
              if (m >= g.M) {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) v[jj] = 0.f;
              }
              const float cs = warp_colsum32(v, lane);
              if (n + lane < g.N) atomicAdd((use_colacc ? colacc : g.colsum) + n + lane, cs);
            }
          } else if (EPI == VSX_EPI_GELU) {
            if (full) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) v[jj] += bs[c + jj];
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) v[jj] = (n + jj < g.N) ? v[jj] + bs[c + jj] : 0.f;
            }
            // out = gelu'(pre-activation), out2 = gelu(pre-activation): both from the fp32 accumulator, so the backward's data gradient
            // only multiplies by the stored derivative (no transcendental math, no bf16-rounded pre-activation in the backward)
            float d[32];
            if (sizeof(OutT) == 2) {
              gelu_and_grad32(v, d);                                          // bf16 training path: packed fp32x2 polynomials
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) d[jj] = gelu_grad_sel<OutT>(v[jj]), v[jj] = gelu_sel<OutT>(v[jj]);
            }
            if (!full) {                                                      // gelu'(0) = 0.5: the zero fill beyond N is explicit
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) d[jj] = (n + jj < g.N) ? d[jj] : 0.f;
            }
            stage_write32<OutT>(tile, row, c, d);
            stage_write32<OutT>(tile2, row, c, v);                            // gelu(0) = 0 keeps the zero fill
          } else if (EPI == VSX_EPI_RESIDUAL) {
            float r[32];
            stage_read32<float>(tile, row, c, r);
            if (n + 32 <= lim) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) r[jj] = fmaf(scale, v[jj] + bs[c + jj], r[jj]);
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) r[jj] += (n + jj < lim) ? scale * (v[jj] + bs[c + jj]) : 0.f;
            }
            stage_write32<float>(tile, row, c, r);
          } else if (EPI == VSX_EPI_GELUGRAD) {
            float u[32];                                                       // aux = gelu'(pre-activation), stored by the forward
            stage_read32<OutT>(tile, row, c, u);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] *= u[jj];
            if (!full) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) v[jj] = (n + jj < g.N) ? v[jj] : 0.f;
            }
            stage_write32<OutT>(tile, row, c, v);
            if (g.colsum != nullptr) {     // fc1 bias gradient fused into the dgrad epilogue
              if (m >= g.M) {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) v[jj] = 0.f;
              }
              const float cs = warp_colsum32(v, lane);
              if (n + lane < g.N) atomicAdd((use_colacc ? colacc : g.colsum) + n + lane, cs);
            }
          }
        };
        // (Software pipelining of the tensor-memory reads -- the next chunk's tcgen05.ld in flight while this one is processed -- was
        // measured: no gain for STORE / GELU, 6 % slower for GELU' (84 extra live registers); the epilogue warps are latency bound by
        // their dependent FMA chains, not by the TMEM port.  tools/gemm_timeline.py, profiles/r2_gemm_epilogue.md.)
        for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 32) {
          if (c >= ncols) break;
          if (has_mma) {
            tmem_ld32(acc + (uint32_t)c, va);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) va[jj] = 0.f;
          }
          process(va, c);
        }
        if (stamp) dbg[un * 8 + 2] = clock64();
        fence_proxy_async();                                       // generic-proxy smem writes -> visible to the TMA (async proxy)
        if (has_mma && sub == UNITS - 1) tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(staged_bar(sb));                             // 8 warps -> the store warp may ship the unit
          if (has_mma && sub == UNITS - 1) {                       // this warp's TMEM reads of the accumulator set are done
            if (CG == 2) mbar_arrive_cluster(acc_empty_leader + 8u * ab);
            else mbar_arrive(acc_empty(ab));
          }
        }
        if (stamp) dbg[un * 8 + 3] = clock64();
      }
    }
    if (use_colacc) {
      named_bar_sync(1, EPI_WARPS * 32);
      for (int i = et * 4; i < grp.colacc_n; i += EPI_WARPS * 32 * 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(colacc + i);
        red_add4(grp.args[0].colsum + i, t4, i, grp.colacc_n);
      }
    }
  }
  tc_fence_before();
  if (CG == 2) {
    cluster_sync_all();          // neither CTA may exit (or free tensor memory) while its peer can still signal its barriers / write its TMEM
    if (warp == 1) tmem_dealloc_pair(tmem_base, TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

constexpr int MAX_GROUP = 4, MAX_GROUP_WIDE = 16;

template <int EPI, typename OutT, int MT, int G, int CG, bool CA>
int launch_ca(const Group<G>& grp, cudaStream_t st) {
  using P = Plan<EPI, OutT, MT, CG, CA>;
  auto kern = gemm_tc_kernel<EPI, OutT, MT, G, CG, CA>;
  static bool configured = false;   // benign race: attribute set is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM);
    if (e != cudaSuccess) {
      set_error("vsx_gemm: cudaFuncSetAttribute(%d) failed: %s", P::SMEM, cudaGetErrorString(e));
      return VSX_ERR_CUDA;
    }
    configured = true;
  }
  if (CG == 1) {
    const int grid = grp.total < num_sms() ? grp.total : num_sms();
    launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), P::SMEM, st, grp);
    return check_launch("vsx_gemm");
  }
  // CTA pairs: clusters of two CTAs (the two SMs of a TPC), one cluster per tile at a time
  const int pairs = num_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (grp.total < pairs ? grp.total : pairs));
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = P::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, grp);
  if (e != cudaSuccess) {
    set_error("vsx_gemm: cluster launch failed: %s", cudaGetErrorString(e));
    return VSX_ERR_CUDA;
  }
  return check_launch("vsx_gemm");
}

// with / without the shared-memory accumulator of the fused column sums (only STORE and GELUGRAD launches can have one)
template <int EPI, typename OutT, int MT, int G, int CG>
int launch_g(const Group<G>& grp, cudaStream_t st) {
  if ((EPI == VSX_EPI_GELUGRAD || EPI == VSX_EPI_STORE) && grp.colacc_n > 0) return launch_ca<EPI, OutT, MT, G, CG, true>(grp, st);
  return launch_ca<EPI, OutT, MT, G, CG, false>(grp, st);
}

// Tile shape and reduction splits for a list of problems that run in one launch.
//  * CTA pairs (256 x 256 tiles, cta_group::2) when those tiles fill the machine's 74 SM pairs at least once: the fewest shared-memory
//    fill bytes and MMA operand reads per flop;
//  * else 256 x 128 single-CTA tiles when they fill the 148 SMs at least once; 128 x 128 tiles for small problems;
//  * weight gradients (ATOMIC): few output tiles, long reductions: the reductions are split so that about two work items per SM
//    (pair) exist over ALL problems (split_k > 1 only permits splitting).
template <int EPI, typename OutT, int G>
int launch(Group<G>& grp, cudaStream_t st) {
  int t128 = 0, t256 = 0, tpair = 0, max_kb = 0;
  long m_sum = 0;
  for (int q = 0; q < grp.count; ++q) {
    const int tn = ceil_div(grp.args[q].n_out, BN);
    t128 += ceil_div(grp.args[q].M, BM) * tn, t256 += ceil_div(grp.args[q].M, 2 * BM) * tn;
    tpair += ceil_div(grp.args[q].M, 2 * BM) * ceil_div(grp.args[q].n_out, 2 * BN);
    max_kb = grp.args[q].num_kb * grp.args[q].terms > max_kb ? grp.args[q].num_kb * grp.args[q].terms : max_kb;
    m_sum += grp.args[q].M;
  }
  static const int env_cg = getenv("VSX_GEMM_CTA_GROUP") ? atoi(getenv("VSX_GEMM_CTA_GROUP")) : 0;
  const int force_cg = g_force_cg != 0 ? g_force_cg : env_cg;
  const int pairs = num_sms() / 2;
  // Measured per shape on the bench step (tools/gemm_table.py with VSX_GEMM_CTA_GROUP=1 / 2, profiles/r2_gemm_cta_pairs.md): pairs win
  // when their 256-wide column tiles waste no more padding than 128-wide ones (0.70-0.95 x the single-CTA time; most at long
  // reductions, where the operand fill dominates), and lose 5-20 % when a third or a fifth column tile of 128 becomes a half-empty
  // pair tile (n_out = 384, 640) -- unless the reduction is long enough (K >= 1536) for the fill savings to outweigh the padding.
  // area of the padded output in 128 x 128 units: 2 * t256 (256 x 128 tiles) vs 4 * tpair (256 x 256 tiles)
  const bool no_extra_padding = 4 * tpair <= 2 * t256;
  bool mt2, cg2;
  if (EPI == VSX_EPI_ATOMIC) {
    // weight gradients: pairs only for small row counts without extra padding (0.78-0.96 x); larger ones measured 4-8 % slower
    cg2 = force_cg == 2 || (force_cg == 0 && g_force_mt == 0 && no_extra_padding && m_sum <= 1024 && 4 * tpair * 5 <= t128 * 6);
    mt2 = !cg2 && (g_force_mt == 2 || (g_force_mt == 0 && 2 * t256 * 5 <= t128 * 6));
    static const int per_sm = getenv("VSX_WGRAD_ITEMS_PER_SM") ? atoi(getenv("VSX_WGRAD_ITEMS_PER_SM")) : 2;
    const int tiles_mn = cg2 ? tpair : (mt2 ? t256 : t128);
    const int units = cg2 ? pairs : num_sms();
    for (int q = 0; q < grp.count; ++q) {
      GemmArgs& g = grp.args[q];
      if (cg2 || mt2 || grp.count > 1) {        // single 128-row problems keep the caller's split (tuned for that shape)
        if (g.split_k > 1) {
          const int want = (per_sm * units) / (tiles_mn > 0 ? tiles_mn : 1);
          g.split_k = want < 1 ? 1 : (want > g.num_kb ? g.num_kb : want);
        }
      }
    }
  } else {
    const bool fills = 2 * tpair >= pairs;      // at least half of the SM pairs get a tile (a 256 x 128 tiling of such a problem does not fill 148 SMs either)
    cg2 = force_cg == 2 || (force_cg == 0 && g_force_mt == 0 && fills && (no_extra_padding || (max_kb >= 24 && 4 * tpair * 4 <= 2 * t256 * 5)));
    mt2 = !cg2 && g_force_mt != 1 && (g_force_mt == 2 || t256 >= num_sms());
    // RESIDUAL with a very short reduction at stage 1 (proj forward with few kept heads: K <= 192) is bound by the latency of its fp32
    // residual tiles: 128 x 128 single-CTA tiles (6.95 instead of 3.47 -> 4 waves) measured 30.3 / 29.8 / 30.8 us against 31.0 / 31.8 / 33.2 us
    // on pairs at K = 64 / 128 / 192; from K = 256 on (and at the smaller stages) the pairs win since they double buffer their staging tile
    // (34.9 vs 35.5 us at K = 256, 37.9 vs 42.2 us at K = 384: tools/gemm_tiles_resid.py)
    if (EPI == VSX_EPI_RESIDUAL && force_cg == 0 && g_force_mt == 0 && max_kb <= 3 && m_sum >= 32768) cg2 = false, mt2 = false;
  }
  int total = 0;
  for (int q = 0; q < grp.count; ++q) {
    const GemmArgs& g = grp.args[q];
    const int tm = ceil_div(g.M, (cg2 || mt2) ? 2 * BM : BM), tn = ceil_div(g.n_out, cg2 ? 2 * BN : BN);
    total += tm * tn * (EPI == VSX_EPI_ATOMIC ? g.split_k : 1);
    grp.tile_end[q] = total;
  }
  grp.total = total;
  if (total == 0) return VSX_OK;
  if (cg2) return launch_g<EPI, OutT, 1, G, 2>(grp, st);
  return mt2 ? launch_g<EPI, OutT, 2, G, 1>(grp, st) : launch_g<EPI, OutT, 1, G, 1>(grp, st);
}

// One problem descriptor -> kernel arguments + tensor maps.  Returns VSX_OK, an error, or 1 when there is nothing to do.
template <int TERMS>
int build_problem(const vsx_gemm_desc* d, TmapPackT<TERMS>& maps, GemmArgs& g) {
  VSX_REQUIRE(d->terms >= 1 && d->terms <= TERMS, "vsx_gemm: this launch takes 1..%d product terms per problem (got %d)", TERMS, d->terms);
  VSX_REQUIRE(d->M > 0 && d->N >= 0 && d->K >= 0, "vsx_gemm: bad extents M=%d N=%d K=%d", d->M, d->N, d->K);
  VSX_REQUIRE(d->lda % 8 == 0 && d->ldb % 8 == 0, "vsx_gemm: operand pitches must be multiples of 8 elements (lda=%ld ldb=%ld)", d->lda, d->ldb);
  VSX_REQUIRE(d->out != nullptr && d->n_out >= d->N && d->n_out <= d->ldo, "vsx_gemm: need N <= n_out <= ldo (N=%d n_out=%d ldo=%ld)", d->N, d->n_out, d->ldo);
  const bool f32 = d->out_dtype == VSX_F32;
  VSX_REQUIRE(d->out_dtype == VSX_F32 || d->out_dtype == VSX_BF16, "vsx_gemm: bad out_dtype %d", d->out_dtype);
  VSX_REQUIRE(d->ldo % (f32 ? 4 : 8) == 0, "vsx_gemm: ldo must be a multiple of %d elements (16 bytes) for the TMA store (ldo=%ld)", f32 ? 4 : 8, d->ldo);
  if (d->n_out == 0) return 1;
  g.M = d->M, g.N = d->N, g.K = d->K, g.num_kb = ceil_div(d->K, BK), g.terms = d->terms;
  g.kseg_kb = 0, g.kseg_stride = 0;
  static const int tile_order = getenv("VSX_GEMM_TILE_ORDER") ? atoi(getenv("VSX_GEMM_TILE_ORDER")) : 0;
  g.m_fastest = tile_order == 1 ? 1 : 0;
  if (d->k_segments > 1) {
    VSX_REQUIRE(d->k_seg_len > 0 && d->k_seg_stride >= d->k_seg_len && (long)(d->k_segments - 1) * d->k_seg_stride + d->k_seg_len <= d->K,
                "vsx_gemm: reduction windows must lie inside [0, K) (segments=%d len=%d stride=%d K=%d)", d->k_segments, d->k_seg_len, d->k_seg_stride, d->K);
    g.kseg_kb = ceil_div(d->k_seg_len, BK), g.kseg_stride = d->k_seg_stride, g.num_kb = d->k_segments * g.kseg_kb;
  }
  g.a_mn = d->a_layout == VSX_MNMAJOR, g.b_mn = d->b_layout == VSX_MNMAJOR;
  g.n_out = d->n_out, g.split_k = d->split_k < 1 ? 1 : d->split_k;
  g.bias = d->bias;
  g.colsum = d->colsum;
  g.row_scale = d->row_scale, g.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1, g.n_keep = d->n_keep;
  g.dbg = g_gemm_dbg;
  if (d->N > 0 && d->K > 0) {
    for (int t = 0; t < d->terms; ++t) {
      VSX_REQUIRE(d->a[t] != nullptr && d->b[t] != nullptr, "vsx_gemm: null operand for term %d", t);
      int rc;
      if (!g.a_mn) rc = make_tmap_2d(&maps.a[t], d->a[t], (uint64_t)d->K, (uint64_t)d->M, (uint64_t)d->lda, BK, BM);
      else         rc = make_tmap_2d(&maps.a[t], d->a[t], (uint64_t)d->M, (uint64_t)d->K, (uint64_t)d->lda, 64, BK);
      if (rc) return rc;
      if (!g.b_mn) rc = make_tmap_2d(&maps.b[t], d->b[t], (uint64_t)d->K, (uint64_t)d->N, (uint64_t)d->ldb, BK, BN);
      else         rc = make_tmap_2d(&maps.b[t], d->b[t], (uint64_t)d->N, (uint64_t)d->K, (uint64_t)d->ldb, 64, BK);
      if (rc) return rc;
    }
  } else {
    g.N = 0;   // nothing to contract: epilogue-only tiles
  }
  // epilogue tensor maps: [M rows, n_out columns], boxes of 128 rows x 128 bytes
  const int odt = f32 ? VSX_F32 : VSX_BF16;
  const uint32_t boxc = f32 ? 32 : 64;
  const uint64_t ocols = d->epilogue == VSX_EPI_ATOMIC ? (uint64_t)d->N : (uint64_t)d->n_out;
  if (ocols == 0) return 1;
  int rc = make_tmap_2d(&maps.out, d->out, ocols, (uint64_t)d->M, (uint64_t)d->ldo, boxc, BM, odt);
  if (rc) return rc;
  maps.out2 = maps.out;
  maps.aux = maps.out;
  switch (d->epilogue) {
    case VSX_EPI_STORE:
      VSX_REQUIRE(d->colsum == nullptr || d->bias == nullptr, "vsx_gemm: STORE with colsum must not add a bias");
      break;
    case VSX_EPI_GELU:
      VSX_REQUIRE(d->out2 != nullptr && d->ldo2 >= d->n_out && d->ldo2 % (f32 ? 4 : 8) == 0, "vsx_gemm: GELU epilogue needs out2 with a 16-byte-multiple pitch");
      if ((rc = make_tmap_2d(&maps.out2, d->out2, ocols, (uint64_t)d->M, (uint64_t)d->ldo2, boxc, BM, odt))) return rc;
      break;
    case VSX_EPI_RESIDUAL:
      VSX_REQUIRE(f32 && d->aux != nullptr, "vsx_gemm: RESIDUAL epilogue is fp32 and needs aux");
    /* fall through */
    case VSX_EPI_GELUGRAD:
      VSX_REQUIRE(d->aux != nullptr && d->ld_aux % (f32 ? 4 : 8) == 0, "vsx_gemm: this epilogue needs aux with a 16-byte-multiple pitch");
      if ((rc = make_tmap_2d(&maps.aux, d->aux, ocols, (uint64_t)d->M, (uint64_t)d->ld_aux, boxc, BM, odt))) return rc;
      break;
    case VSX_EPI_ATOMIC:
      VSX_REQUIRE(f32, "vsx_gemm: ATOMIC epilogue accumulates into fp32");
      if (g.N == 0) return 1;
      g.split_k = (g.split_k > g.num_kb ? (g.num_kb > 0 ? g.num_kb : 1) : g.split_k);
      g.n_out = d->N;
      break;
    default:
      set_error("vsx_gemm: unknown epilogue %d", d->epilogue);
      return VSX_ERR_ARG;
  }
  return VSX_OK;
}

template <int G>
int run_group(const vsx_gemm_desc* descs, int count, cudaStream_t st) {
  static thread_local Group<G>* tl = nullptr;     // host staging of the kernel parameter block (forward and backward threads each own one)
  if (tl == nullptr) tl = new Group<G>();
  Group<G>& gr = *tl;
  gr.count = 0;
  for (int i = 0; i < count; ++i) {
    VSX_REQUIRE(descs[i].epilogue == descs[0].epilogue && descs[i].out_dtype == descs[0].out_dtype,
                "vsx_gemm_grouped: all problems of a launch must share the epilogue and the output dtype");
    const int rc = build_problem(&descs[i], gr.maps[gr.count], gr.args[gr.count]);
    if (rc < 0) return rc;
    if (rc == 0) ++gr.count;
  }
  if (gr.count == 0) return VSX_OK;
  // fused column sums (bias gradients): per-CTA shared-memory partials when every problem of the launch adds into the same vector
  gr.colacc_n = 0;
  if (gr.args[0].colsum != nullptr) {
    int nmax = 0;
    bool same = true;
    for (int q = 0; q < gr.count; ++q) same = same && gr.args[q].colsum == gr.args[0].colsum, nmax = gr.args[q].N > nmax ? gr.args[q].N : nmax;
    if (same && nmax <= COLACC_MAX) gr.colacc_n = nmax;
  }
  const bool f32 = descs[0].out_dtype == VSX_F32;
  switch (descs[0].epilogue) {
    case VSX_EPI_STORE: return f32 ? launch<VSX_EPI_STORE, float, G>(gr, st) : launch<VSX_EPI_STORE, bf16, G>(gr, st);
    case VSX_EPI_GELU: return f32 ? launch<VSX_EPI_GELU, float, G>(gr, st) : launch<VSX_EPI_GELU, bf16, G>(gr, st);
    case VSX_EPI_RESIDUAL: return launch<VSX_EPI_RESIDUAL, float, G>(gr, st);
    case VSX_EPI_GELUGRAD: return f32 ? launch<VSX_EPI_GELUGRAD, float, G>(gr, st) : launch<VSX_EPI_GELUGRAD, bf16, G>(gr, st);
    case VSX_EPI_ATOMIC: return launch<VSX_EPI_ATOMIC, float, G>(gr, st);
  }
  set_error("vsx_gemm: unknown epilogue %d", descs[0].epilogue);
  return VSX_ERR_ARG;
}

}  // namespace
}  // namespace vsx

using namespace vsx;

extern "C" int vsx_gemm_debug_buffer(void* p) {
  g_gemm_dbg = static_cast<long long*>(p);
  return VSX_OK;
}

extern "C" int vsx_gemm_force_tile_rows(int rows) {
  VSX_REQUIRE(rows == 0 || rows == 128 || rows == 256, "vsx_gemm_force_tile_rows: 0 (heuristic), 128 or 256");
  g_force_mt = rows / 128;
  return VSX_OK;
}

extern "C" int vsx_gemm_force_cta_group(int cta_group) {
  VSX_REQUIRE(cta_group == 0 || cta_group == 1 || cta_group == 2, "vsx_gemm_force_cta_group: 0 (heuristic), 1 or 2");
  g_force_cg = cta_group;
  return VSX_OK;
}

extern "C" int vsx_gemm(const vsx_gemm_desc* d, void* stream) {
  VSX_REQUIRE(d != nullptr, "vsx_gemm: null descriptor");
  return run_group<1>(d, 1, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int vsx_gemm_grouped(const vsx_gemm_desc* descs, int count, void* stream) {
  VSX_REQUIRE(descs != nullptr && count >= 1 && count <= MAX_GROUP_WIDE, "vsx_gemm_grouped: 1..%d problems per launch (got %d)", MAX_GROUP_WIDE, count);
  if (count == 1) return run_group<1>(descs, 1, reinterpret_cast<cudaStream_t>(stream));
  if (count <= MAX_GROUP) return run_group<MAX_GROUP>(descs, count, reinterpret_cast<cudaStream_t>(stream));
  return run_group<MAX_GROUP_WIDE>(descs, count, reinterpret_cast<cudaStream_t>(stream));
}
