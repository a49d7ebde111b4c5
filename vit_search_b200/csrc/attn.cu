// C-ABI entry points of the attention core (dispatch between the tensor-core and the fp32 kernels).
#include "common.cuh"

namespace vsx {
template <typename T>
int attn_fwd_ref(const void* qkv, void* o, float* lse, int B, int N, int H, int D, int Hk, float scale, cudaStream_t st);
template <typename T>
int attn_bwd_ref(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int B, int N, int H, int D, int Hk,
                 float scale, cudaStream_t st);
int attn_fwd_mma(const void* qkv, void* o, float* lse, int B, int N, int H, int D, int Hk, float scale, cudaStream_t st);
int attn_bwd_mma(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int B, int N, int H, int D, int Hk,
                 float scale, float* dbias, cudaStream_t st);
bool attn_mma_supported(int N, int D);
int attn_fwd_tc(const void* qkv, void* o, float* lse, int B, int N, int H, int D, int Hk, float scale, cudaStream_t st,
                const vsx_sample_segments* sg = nullptr);
int attn_bwd_tc(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int B, int N, int H, int D, int Hk, float scale,
                float* dbias, cudaStream_t st, const vsx_sample_segments* sg = nullptr);
bool attn_tc_supported(int N, int D);
void attn_set_debug(long long* p);
void attn_set_odd_modes(int f, int b, int s);
}  // namespace vsx

using namespace vsx;

// Which bf16 kernel serves a shape.  The tcgen05 kernels work on 128-query tiles: at N = 17 (stage 3) a (sample, head) pair fills 13 % of
// a tile and every pair still pays the full TMA -> MMA -> softmax -> MMA -> drain chain, so the register-resident mma.sync kernel is
// faster there (B = 256, H = 12, D = 64, tools/attn_bench.py: forward 13.4 vs 32.8 us, backward 61 vs 104 us); from N = 65 up the
// tensor-memory kernels win (backward 94 vs 201 us at N = 65, 184 vs 389 us at N = 257).
static bool prefer_tc(int impl, int tokens, int head_dim) {
  if (!attn_tc_supported(tokens, head_dim)) return false;
  if (impl == VSX_ATTN_IMPL_TCGEN05) return true;
  return impl == VSX_ATTN_IMPL_AUTO && !(tokens <= 32 && attn_mma_supported(tokens, head_dim));
}

static int check_shape(const char* what, int B, int N, int H, int D, int Hk) {
  VSX_REQUIRE(B >= 0 && N > 0 && H > 0 && Hk >= 0 && Hk <= H, "%s: bad shape batch=%d tokens=%d heads=%d heads_keep=%d", what, B, N, H, Hk);
  VSX_REQUIRE(D % 4 == 0 && D > 0 && D <= 64, "%s: head_dim must be a multiple of 4 and <= 64 (got %d)", what, D);
  return VSX_OK;
}

extern "C" int vsx_attn_fwd(const void* qkv, void* o, float* lse, int dtype, int batch, int tokens, int heads, int head_dim,
                            int heads_keep, float scale, int impl, void* stream) {
  int rc = check_shape("vsx_attn_fwd", batch, tokens, heads, head_dim, heads_keep);
  if (rc) return rc;
  if (batch == 0) return VSX_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == VSX_F32) return attn_fwd_ref<float>(qkv, o, lse, batch, tokens, heads, head_dim, heads_keep, scale, st);
  if (dtype == VSX_BF16) {
    if (prefer_tc(impl, tokens, head_dim)) return attn_fwd_tc(qkv, o, lse, batch, tokens, heads, head_dim, heads_keep, scale, st);
    VSX_REQUIRE(impl != VSX_ATTN_IMPL_TCGEN05, "vsx_attn_fwd: the tcgen05 kernel needs head_dim 32 / 48 / 64 and tokens <= 288 (got %d, %d)", head_dim, tokens);
    if (impl != VSX_ATTN_IMPL_FP32 && attn_mma_supported(tokens, head_dim))
      return attn_fwd_mma(qkv, o, lse, batch, tokens, heads, head_dim, heads_keep, scale, st);
    return attn_fwd_ref<bf16>(qkv, o, lse, batch, tokens, heads, head_dim, heads_keep, scale, st);
  }
  set_error("vsx_attn_fwd: bad dtype %d", dtype);
  return VSX_ERR_ARG;
}

extern "C" int vsx_colsum(const void* x, int dtype, long ldx, int rows, int cols, float* out, void* stream);

extern "C" int vsx_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int dtype, int batch,
                            int tokens, int heads, int head_dim, int heads_keep, float scale, int impl, float* dbias, void* stream) {
  int rc = check_shape("vsx_attn_bwd", batch, tokens, heads, head_dim, heads_keep);
  if (rc) return rc;
  if (batch == 0) return VSX_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long ld = 3L * heads * head_dim;
  if (dtype == VSX_F32) {
    rc = attn_bwd_ref<float>(qkv, o, d_o, lse, dqkv, batch, tokens, heads, head_dim, heads_keep, scale, st);
    if (rc == VSX_OK && dbias != nullptr) rc = vsx_colsum(dqkv, dtype, ld, batch * tokens, (int)ld, dbias, stream);   // fused only in the tensor-core kernel
    return rc;
  }
  if (dtype == VSX_BF16) {
    if (prefer_tc(impl, tokens, head_dim)) return attn_bwd_tc(qkv, o, d_o, lse, dqkv, batch, tokens, heads, head_dim, heads_keep, scale, dbias, st);
    VSX_REQUIRE(impl != VSX_ATTN_IMPL_TCGEN05, "vsx_attn_bwd: the tcgen05 kernel needs head_dim 32 / 48 / 64 and tokens <= 288 (got %d, %d)", head_dim, tokens);
    if (impl != VSX_ATTN_IMPL_FP32 && attn_mma_supported(tokens, head_dim))
      return attn_bwd_mma(qkv, o, d_o, lse, dqkv, batch, tokens, heads, head_dim, heads_keep, scale, dbias, st);
    rc = attn_bwd_ref<bf16>(qkv, o, d_o, lse, dqkv, batch, tokens, heads, head_dim, heads_keep, scale, st);
    if (rc == VSX_OK && dbias != nullptr) rc = vsx_colsum(dqkv, dtype, ld, batch * tokens, (int)ld, dbias, stream);
    return rc;
  }
  set_error("vsx_attn_bwd: bad dtype %d", dtype);
  return VSX_ERR_ARG;
}

// ---------------------------------------------------------------- several head extents in one launch (multi-architecture batches)
static int check_sample_segs(const vsx_sample_segments* sg, int batch, int heads, int* hmax, const char* what) {
  VSX_REQUIRE(sg != nullptr && sg->count >= 1 && sg->count <= VSX_MAX_SEGMENTS, "%s: 1..%d segments", what, VSX_MAX_SEGMENTS);
  int prev = 0;
  *hmax = 0;
  for (int i = 0; i < sg->count; ++i) {
    VSX_REQUIRE(sg->sample_end[i] >= prev && sg->heads_keep[i] >= 0 && sg->heads_keep[i] <= heads, "%s: segment %d is not ordered or keeps more than %d heads", what, i, heads);
    prev = sg->sample_end[i];
    *hmax = sg->heads_keep[i] > *hmax ? sg->heads_keep[i] : *hmax;
  }
  VSX_REQUIRE(prev == batch, "%s: the segments must cover the %d samples (last sample_end = %d)", what, batch, prev);
  return VSX_OK;
}

extern "C" int vsx_attn_fwd_segs(const void* qkv, void* o, float* lse, int dtype, int batch, int tokens, int heads, int head_dim,
                                 const vsx_sample_segments* segs, float scale, int impl, void* stream) {
  int hmax = 0;
  int rc = check_sample_segs(segs, batch, heads, &hmax, "vsx_attn_fwd_segs");
  if (rc) return rc;
  if ((rc = check_shape("vsx_attn_fwd_segs", batch, tokens, heads, head_dim, hmax))) return rc;
  if (batch == 0 || hmax == 0) return VSX_OK;
  if (dtype == VSX_BF16 && prefer_tc(impl, tokens, head_dim))
    return attn_fwd_tc(qkv, o, lse, batch, tokens, heads, head_dim, hmax, scale, reinterpret_cast<cudaStream_t>(stream), segs);
  const size_t es = dtype == VSX_F32 ? 4 : 2;       // other implementations: one launch per segment
  const long HD = (long)heads * head_dim;
  for (int i = 0, b0 = 0; i < segs->count; b0 = segs->sample_end[i], ++i) {
    const int nb = segs->sample_end[i] - b0;
    if (nb == 0 || segs->heads_keep[i] == 0) continue;
    rc = vsx_attn_fwd(static_cast<const uint8_t*>(qkv) + (size_t)b0 * tokens * 3 * HD * es, static_cast<uint8_t*>(o) + (size_t)b0 * tokens * HD * es,
                      lse + (long)b0 * heads * tokens, dtype, nb, tokens, heads, head_dim, segs->heads_keep[i], scale, impl, stream);
    if (rc) return rc;
  }
  return VSX_OK;
}

extern "C" int vsx_attn_bwd_segs(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int dtype, int batch, int tokens,
                                 int heads, int head_dim, const vsx_sample_segments* segs, float scale, int impl, float* dbias, void* stream) {
  int hmax = 0;
  int rc = check_sample_segs(segs, batch, heads, &hmax, "vsx_attn_bwd_segs");
  if (rc) return rc;
  if ((rc = check_shape("vsx_attn_bwd_segs", batch, tokens, heads, head_dim, hmax))) return rc;
  if (batch == 0 || hmax == 0) return VSX_OK;
  if (dtype == VSX_BF16 && prefer_tc(impl, tokens, head_dim))
    return attn_bwd_tc(qkv, o, d_o, lse, dqkv, batch, tokens, heads, head_dim, hmax, scale, dbias, reinterpret_cast<cudaStream_t>(stream), segs);
  const size_t es = dtype == VSX_F32 ? 4 : 2;
  const long HD = (long)heads * head_dim;
  for (int i = 0, b0 = 0; i < segs->count; b0 = segs->sample_end[i], ++i) {
    const int nb = segs->sample_end[i] - b0;
    if (nb == 0 || segs->heads_keep[i] == 0) continue;
    const size_t ro = (size_t)b0 * tokens;
    rc = vsx_attn_bwd(static_cast<const uint8_t*>(qkv) + ro * 3 * HD * es, static_cast<const uint8_t*>(o) + ro * HD * es,
                      static_cast<const uint8_t*>(d_o) + ro * HD * es, lse + (long)b0 * heads * tokens, static_cast<uint8_t*>(dqkv) + ro * 3 * HD * es, dtype, nb,
                      tokens, heads, head_dim, segs->heads_keep[i], scale, impl, dbias, stream);
    if (rc) return rc;
  }
  return VSX_OK;
}

/* Development aid (tools/attn_timeline.py): device buffer of 64 x 8 clock64 stamps written by CTA 0 of the tcgen05 backward kernel. */
extern "C" int vsx_attn_odd_token_modes(int forward, int backward_large, int backward_small) {
  attn_set_odd_modes(forward, backward_large, backward_small);
  return VSX_OK;
}

extern "C" int vsx_attn_debug_buffer(void* p) {
  attn_set_debug(static_cast<long long*>(p));
  return VSX_OK;
}
