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
#include "common.cuh"

namespace vsx {

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 2;              // x 3 co-resident CTAs per SM = 6 operand stages in flight per SM
constexpr int STAGE_A = BM * BK * 2;
constexpr int STAGE_B = BN * BK * 2;
constexpr int TMEM_COLS = 128;
constexpr int EPI_WARPS = 8;            // two warps per TMEM lane quarter, each owning half of the tile's columns
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int SMEM_BYTES = STAGES * (STAGE_A + STAGE_B) + 1024 /*align*/ + 128 /*barriers, tmem slot*/ + 512 /*bias tile*/;

constexpr int MAX_TERMS = 6;
struct TmapPack {
  CUtensorMap a[MAX_TERMS];
  CUtensorMap b[MAX_TERMS];
  CUtensorMap out, out2, aux;   // epilogue tiles (128B-swizzled boxes of 128 rows x 128 bytes)
};

struct GemmArgs {
  int M, N, K, num_kb, terms, a_mn, b_mn, n_out, split_k;
  const float* bias;
  const float* row_scale;
  float* colsum;
  int rows_per_sample, n_keep;
};

// Shared-memory matrix descriptor (tcgen05), 128B swizzle.  K-major: rows of 128 B, 8-row groups 1024 B apart (SBO).
// MN-major: 64-element (128 B) MN atoms, k-rows 128 B apart, 8-row groups SBO = 1024 B, MN atoms LBO = BK*128 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, bool mn_major) {
  const uint64_t lbo = mn_major ? (uint64_t)((BK * 128) >> 4) : 0ull;
  const uint64_t sbo = 1024 >> 4;
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}

// Instruction descriptor, kind::f16: D=f32, A=B=bf16, M=128, N=128.
__device__ __forceinline__ uint32_t make_idesc(bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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

template <int EPI, typename OutT>
__global__ void __launch_bounds__(GEMM_THREADS, 3) gemm_tc_kernel(const __grid_constant__ TmapPack maps, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;   // 128B-swizzle atoms need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + STAGES * STAGE_A;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (STAGE_A + STAGE_B));
  const uint32_t bar0 = base + STAGES * (STAGE_A + STAGE_B);
  // bars: [0,S) full, [S,2S) empty, [2S] accumulator ready, [2S+1] aux tile landed; then the TMEM base slot and the bias tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);
  float* bias_s = reinterpret_cast<float*>(bars + 2 * STAGES + 4);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t acc_bar = bar0 + 8u * (2 * STAGES);
  const uint32_t aux_bar = bar0 + 8u * (2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  int kb_begin = 0, kb_end = g.num_kb;
  if (EPI == VSX_EPI_ATOMIC) {
    const int per = (g.num_kb + g.split_k - 1) / g.split_k;
    kb_begin = blockIdx.z * per;
    kb_end = min(g.num_kb, kb_begin + per);
    if (kb_begin >= kb_end) return;   // uniform for the whole CTA
  }
  const int nkb = kb_end - kb_begin;
  const int iters = nkb * g.terms;
  const bool has_mma = (n0 < g.N) && iters > 0;

  if (warp == 0 && lane == 0) {
    for (int t = 0; t < g.terms; ++t) {
      tma_prefetch_desc(&maps.a[t]);
      tma_prefetch_desc(&maps.b[t]);
    }
    tma_prefetch_desc(&maps.out);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      mbar_init(acc_bar, 1);
      mbar_init(aux_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (has_mma && warp == 0 && lane == 0) {
    // ---------------- TMA producer ----------------
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(empty_bar(s), ph ^ 1u);
      const int term = it / nkb, k0 = (kb_begin + it % nkb) * BK;
      const uint32_t dA = sA + s * STAGE_A, dB = sB + s * STAGE_B, fb = full_bar(s);
      mbar_expect_tx(fb, STAGE_A + STAGE_B);
      if (!g.a_mn) {
        tma_load_2d(dA, &maps.a[term], fb, k0, m0);
      } else {
        tma_load_2d(dA, &maps.a[term], fb, m0, k0);
        tma_load_2d(dA + BK * 128, &maps.a[term], fb, m0 + 64, k0);
      }
      if (!g.b_mn) {
        tma_load_2d(dB, &maps.b[term], fb, k0, n0);
      } else {
        tma_load_2d(dB, &maps.b[term], fb, n0, k0);
        tma_load_2d(dB + BK * 128, &maps.b[term], fb, n0 + 64, k0);
      }
    }
  } else if (has_mma && warp == 1 && lane == 0) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = make_idesc(g.a_mn != 0, g.b_mn != 0);
    const uint32_t a_step = g.a_mn ? (UMMA_K * 128) : (UMMA_K * 2);   // bytes per 16-wide k step
    const uint32_t b_step = g.b_mn ? (UMMA_K * 128) : (UMMA_K * 2);
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t aA = sA + s * STAGE_A, aB = sB + s * STAGE_B;
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k) {
        umma_bf16(tmem_base, make_smem_desc(aA + k * a_step, g.a_mn != 0), make_smem_desc(aB + k * b_step, g.b_mn != 0),
                  idesc, (it | k) != 0 ? 1u : 0u);
      }
      umma_commit(empty_bar(s));   // slot reusable once these MMAs have read it
    }
    umma_commit(acc_bar);          // accumulator complete
  } else if (warp >= 2) {
    // ---------------- epilogue: TMEM -> registers -> fused math -> swizzled smem tile -> TMA store / reduce-add ----------------
    constexpr int BOXC = 128 / (int)sizeof(OutT);     // columns per 128-byte box row
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                    // tile row == TMEM lane
    const int et = (int)threadIdx.x - 64;             // 0..255 among the epilogue threads
    const int chalf = (warp - 2) >> 2;                // which 64-column half of the tile this warp handles
    const int m = m0 + row;
    const int ncols = min(BN, g.n_out - n0);          // columns of this tile that exist in the output
    if (et < BN) bias_s[et] = (g.bias != nullptr && n0 + et < g.N) ? __ldg(g.bias + n0 + et) : 0.f;
    if (has_mma) {
      mbar_wait(acc_bar, 0);                          // all MMAs done => the operand ring is free: reuse it as the staging tile
      tc_fence_after();
    }
    uint8_t* tile = smem;                             // generic view of the (now idle) stage memory
    // GELU writes two tiles (pre-activation and activation): side by side for bf16 (2 x 32 KB); for fp32 (2 x 64 KB > ring)
    // the activation tile is produced in a second pass over the accumulator after the first tile has left.
    constexpr bool TWO_PASS = (EPI == VSX_EPI_GELU) && sizeof(OutT) == 4;
    uint8_t* tile2 = TWO_PASS ? smem : smem + (BN / BOXC) * BOX_BYTES;
    const int nbox = (ncols + BOXC - 1) / BOXC;
    if (EPI == VSX_EPI_RESIDUAL || EPI == VSX_EPI_GELUGRAD) {
      if (et == 0) {                                  // aux tile (residual stream / pre-activation) by TMA, OOB zero-filled
        mbar_expect_tx(aux_bar, (uint32_t)nbox * BOX_BYTES);
        for (int bx = 0; bx < nbox; ++bx) tma_load_2d(base + bx * BOX_BYTES, &maps.aux, aux_bar, n0 + bx * BOXC, m0);
      }
    }
    named_bar_sync(1, EPI_WARPS * 32);                           // bias tile visible
    if (EPI == VSX_EPI_RESIDUAL || EPI == VSX_EPI_GELUGRAD) mbar_wait(aux_bar, 0);
    float scale = 1.0f;
    if (EPI == VSX_EPI_RESIDUAL && g.row_scale != nullptr && m < g.M) scale = __ldg(g.row_scale + m / g.rows_per_sample);
    const int lim = g.n_keep < g.N ? g.n_keep : g.N;
    for (int pass = 0; pass < (TWO_PASS ? 2 : 1); ++pass) {
      for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 32) {
        if (c >= ncols) break;
        float v[32];
        if (has_mma) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        const int n = n0 + c;
        if (EPI == VSX_EPI_ATOMIC) {
          stage_write32<float>(tile, row, c, v);
        } else if (EPI == VSX_EPI_STORE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (n + j < g.N) ? v[j] + bias_s[c + j] : 0.f;
          stage_write32<OutT>(tile, row, c, v);
        } else if (EPI == VSX_EPI_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (n + j < g.N) ? v[j] + bias_s[c + j] : 0.f;
          if (!TWO_PASS || pass == 0) stage_write32<OutT>(tile, row, c, v);
          if (!TWO_PASS || pass == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);   // gelu(0) = 0 keeps the zero fill
            stage_write32<OutT>(tile2, row, c, v);
          }
        } else if (EPI == VSX_EPI_RESIDUAL) {
          float r[32];
          stage_read32<float>(tile, row, c, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] += (n + j < lim) ? scale * (v[j] + bias_s[c + j]) : 0.f;
          stage_write32<float>(tile, row, c, r);
        } else if (EPI == VSX_EPI_GELUGRAD) {
          float u[32];
          stage_read32<OutT>(tile, row, c, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (n + j < g.N) ? v[j] * gelu_grad_f(u[j]) : 0.f;
          stage_write32<OutT>(tile, row, c, v);
        }
      }
      fence_proxy_async();                            // generic-proxy smem writes -> visible to the TMA (async proxy)
      named_bar_sync(1, EPI_WARPS * 32);
      if ((EPI == VSX_EPI_GELUGRAD || EPI == VSX_EPI_STORE) && g.colsum != nullptr && et >= 128 && n0 + (et - 128) < g.N) {
        // bias gradient fused into the dgrad epilogue: column sums of the staged (already rounded) tile over its valid rows
        const int cc = et - 128, rmax = min(BM, g.M - m0);
        const uint8_t* colp = tile + (cc / BOXC) * BOX_BYTES + (cc % (16 / (int)sizeof(OutT))) * (int)sizeof(OutT);
        const int chunk = (cc % BOXC) / (16 / (int)sizeof(OutT));
        float acc = 0.f;
        for (int r = 0; r < rmax; ++r) acc += Store<OutT>::ld(reinterpret_cast<const OutT*>(colp + box_off(r, chunk)));
        atomicAdd(g.colsum + n0 + cc, acc);
      }
      if (et == 0) {
        for (int bx = 0; bx < nbox; ++bx) {
          if (EPI == VSX_EPI_ATOMIC) {
            tma_reduce_add_2d(&maps.out, base + bx * BOX_BYTES, n0 + bx * BOXC, m0);
          } else if (EPI == VSX_EPI_GELU) {
            if (!TWO_PASS || pass == 0) tma_store_2d(&maps.out, base + bx * BOX_BYTES, n0 + bx * BOXC, m0);
            if (!TWO_PASS) tma_store_2d(&maps.out2, base + (BN / BOXC + bx) * BOX_BYTES, n0 + bx * BOXC, m0);
            if (TWO_PASS && pass == 1) tma_store_2d(&maps.out2, base + bx * BOX_BYTES, n0 + bx * BOXC, m0);
          } else {
            tma_store_2d(&maps.out, base + bx * BOX_BYTES, n0 + bx * BOXC, m0);
          }
        }
        tma_store_commit_wait();                      // shared memory must stay valid until the TMA has read it
      }
      if (TWO_PASS) named_bar_sync(1, EPI_WARPS * 32);           // the staging tile may be overwritten by the second pass
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int EPI, typename OutT>
int launch(const TmapPack& maps, const GemmArgs& g, dim3 grid, cudaStream_t st) {
  static bool configured = false;   // benign race: attribute set is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<EPI, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("vsx_gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return VSX_ERR_CUDA;
    }
    configured = true;
  }
  gemm_tc_kernel<EPI, OutT><<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(maps, g);
  return check_launch("vsx_gemm");
}

}  // namespace
}  // namespace vsx

using namespace vsx;

extern "C" int vsx_gemm(const vsx_gemm_desc* d, void* stream) {
  VSX_REQUIRE(d != nullptr, "vsx_gemm: null descriptor");
  VSX_REQUIRE(d->terms >= 1 && d->terms <= MAX_TERMS, "vsx_gemm: terms must be 1..6 (got %d)", d->terms);
  VSX_REQUIRE(d->M > 0 && d->N >= 0 && d->K >= 0, "vsx_gemm: bad extents M=%d N=%d K=%d", d->M, d->N, d->K);
  VSX_REQUIRE(d->lda % 8 == 0 && d->ldb % 8 == 0, "vsx_gemm: operand pitches must be multiples of 8 elements (lda=%ld ldb=%ld)", d->lda, d->ldb);
  VSX_REQUIRE(d->out != nullptr && d->n_out >= d->N && d->n_out <= d->ldo, "vsx_gemm: need N <= n_out <= ldo (N=%d n_out=%d ldo=%ld)", d->N, d->n_out, d->ldo);
  const bool f32 = d->out_dtype == VSX_F32;
  VSX_REQUIRE(d->out_dtype == VSX_F32 || d->out_dtype == VSX_BF16, "vsx_gemm: bad out_dtype %d", d->out_dtype);
  VSX_REQUIRE(d->ldo % (f32 ? 4 : 8) == 0, "vsx_gemm: ldo must be a multiple of %d elements (16 bytes) for the TMA store (ldo=%ld)", f32 ? 4 : 8, d->ldo);
  if (d->n_out == 0) return VSX_OK;

  GemmArgs g;
  g.M = d->M, g.N = d->N, g.K = d->K, g.num_kb = ceil_div(d->K, BK), g.terms = d->terms;
  g.a_mn = d->a_layout == VSX_MNMAJOR, g.b_mn = d->b_layout == VSX_MNMAJOR;
  g.n_out = d->n_out, g.split_k = d->split_k < 1 ? 1 : d->split_k;
  g.bias = d->bias;
  g.colsum = d->colsum;
  g.row_scale = d->row_scale, g.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1, g.n_keep = d->n_keep;

  TmapPack maps;
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
    memset(&maps, 0, sizeof(maps));
  }
  dim3 grid(ceil_div(d->M, BM), ceil_div(d->n_out, BN), 1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {  // epilogue tensor maps: [M rows, n_out columns], boxes of 128 rows x 128 bytes
    const int odt = f32 ? VSX_F32 : VSX_BF16;
    const uint32_t boxc = f32 ? 32 : 64;
    const uint64_t ocols = d->epilogue == VSX_EPI_ATOMIC ? (uint64_t)d->N : (uint64_t)d->n_out;
    if (ocols == 0) return VSX_OK;
    int rc = make_tmap_2d(&maps.out, d->out, ocols, (uint64_t)d->M, (uint64_t)d->ldo, boxc, BM, odt);
    if (rc) return rc;
    maps.out2 = maps.out;
    maps.aux = maps.out;
    if (d->epilogue == VSX_EPI_GELU) {
      VSX_REQUIRE(d->out2 != nullptr && d->ldo2 >= d->n_out && d->ldo2 % (f32 ? 4 : 8) == 0, "vsx_gemm: GELU epilogue needs out2 with a 16-byte-multiple pitch");
      rc = make_tmap_2d(&maps.out2, d->out2, ocols, (uint64_t)d->M, (uint64_t)d->ldo2, boxc, BM, odt);
      if (rc) return rc;
    }
    if (d->epilogue == VSX_EPI_RESIDUAL || d->epilogue == VSX_EPI_GELUGRAD) {
      VSX_REQUIRE(d->aux != nullptr && d->ld_aux % (f32 ? 4 : 8) == 0, "vsx_gemm: this epilogue needs aux with a 16-byte-multiple pitch");
      rc = make_tmap_2d(&maps.aux, d->aux, ocols, (uint64_t)d->M, (uint64_t)d->ld_aux, boxc, BM, odt);
      if (rc) return rc;
    }
  }
  switch (d->epilogue) {
    case VSX_EPI_STORE:
      VSX_REQUIRE(d->colsum == nullptr || d->bias == nullptr, "vsx_gemm: STORE with colsum must not add a bias");
      return f32 ? launch<VSX_EPI_STORE, float>(maps, g, grid, st) : launch<VSX_EPI_STORE, bf16>(maps, g, grid, st);
    case VSX_EPI_GELU:
      return f32 ? launch<VSX_EPI_GELU, float>(maps, g, grid, st) : launch<VSX_EPI_GELU, bf16>(maps, g, grid, st);
    case VSX_EPI_RESIDUAL:
      VSX_REQUIRE(f32 && d->aux != nullptr, "vsx_gemm: RESIDUAL epilogue is fp32 and needs aux");
      return launch<VSX_EPI_RESIDUAL, float>(maps, g, grid, st);
    case VSX_EPI_GELUGRAD:
      VSX_REQUIRE(d->aux != nullptr, "vsx_gemm: GELUGRAD epilogue needs aux (pre-activation)");
      return f32 ? launch<VSX_EPI_GELUGRAD, float>(maps, g, grid, st) : launch<VSX_EPI_GELUGRAD, bf16>(maps, g, grid, st);
    case VSX_EPI_ATOMIC:
      VSX_REQUIRE(f32, "vsx_gemm: ATOMIC epilogue accumulates into fp32");
      if (g.N == 0) return VSX_OK;
      grid.y = ceil_div(d->N, BN);
      grid.z = g.split_k = (g.split_k > g.num_kb ? (g.num_kb > 0 ? g.num_kb : 1) : g.split_k);
      g.n_out = d->N;
      return launch<VSX_EPI_ATOMIC, float>(maps, g, grid, st);
    default:
      set_error("vsx_gemm: unknown epilogue %d", d->epilogue);
      return VSX_ERR_ARG;
  }
}
