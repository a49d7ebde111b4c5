// One C-ABI call per half of a transformer Block (bf16 training path): the host-side orchestration of
//     x_out = x + mask * drop_path( branch( LN(x) ) )          (nets/supernet_blocks.py:213-253, Attention :100-120, Mlp :37-52)
// and of its backward, i.e. the per-segment sequence of LayerNorm, GEMM, attention and gradient-cast launches that
// vit_search_b200/core.py otherwise issues from Python (~10 ctypes calls per half block and direction).  Nothing is computed here:
// every step is one of the kernels behind include/vsx.h.  A "segment" is a run of consecutive samples that share one sub-architecture
// in this layer (embed / inner / output keep counts); masked heads, hidden channels, embedding channels and dropped layers are
// never computed.
#include <string.h>

#include "common.cuh"

using namespace vsx;

namespace {

inline int up8(int n) { return (n + 7) / 8 * 8; }
inline int up64(int n) { return (n + 63) / 64 * 64; }

inline int split_k_for(int m_rows, int n_cols, int red_rows) {
  const int tiles = ceil_div(m_rows, 128) * ceil_div(n_cols, 128);
  int kb = ceil_div(red_rows, 64);
  if (kb < 1) kb = 1;
  int s = (2 * 148) / (tiles > 0 ? tiles : 1);
  s = s < kb ? s : kb;
  return s < 1 ? 1 : s;
}

struct Gemm {
  vsx_gemm_desc d;
  Gemm(const void* a, long lda, int a_layout, const void* b, long ldb, int b_layout, int M, int N, int K, int epi, int out_dtype, void* out,
       long ldo) {
    memset(&d, 0, sizeof(d));
    d.a[0] = a, d.b[0] = b, d.terms = 1, d.lda = lda, d.ldb = ldb, d.a_layout = a_layout, d.b_layout = b_layout;
    d.M = M, d.N = N, d.K = K, d.epilogue = epi, d.out_dtype = out_dtype, d.out = out, d.ldo = ldo, d.n_out = N;
    d.rows_per_sample = 1, d.split_k = 1;
  }
};

#define HB_CHECK(expr)      \
  do {                      \
    const int rc_ = (expr); \
    if (rc_ != VSX_OK) return rc_; \
  } while (0)

inline const bf16* B16(const void* p) { return static_cast<const bf16*>(p); }
inline bf16* B16(void* p) { return static_cast<bf16*>(p); }

int pre_norm_or_cast(const vsx_half_block* h, const vsx_segment& s, long r0, int rows, void* stream) {
  const int C = h->width;
  if (h->pre_norm)
    return vsx_masked_ln_fwd(h->x + r0 * C, C, h->ln_w, h->ln_b, B16(h->xn) + r0 * C, nullptr, VSX_BF16, C, h->mean + r0, h->rstd + r0, rows, C,
                             s.embed_keep, h->eps, 0, 0, stream);
  return vsx_scale_mask_cast(h->x + r0 * C, C, nullptr, 1, s.embed_keep, B16(h->xn) + r0 * C, VSX_BF16, C, rows, C, nullptr, stream);
}

int passthrough(const float* src, float* dst, long elems, bool copy, void* stream) {
  cudaError_t e = copy ? cudaMemcpyAsync(dst, src, elems * sizeof(float), cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream))
                       : cudaMemsetAsync(dst, 0, elems * sizeof(float), reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    set_error("vsx_half_block: pass-through of a dropped layer failed: %s", cudaGetErrorString(e));
    return VSX_ERR_CUDA;
  }
  return VSX_OK;
}

int check_desc(const vsx_half_block* h, const char* what) {
  VSX_REQUIRE(h != nullptr && (h->kind == VSX_HALF_ATTN || h->kind == VSX_HALF_MLP), "%s: bad descriptor", what);
  VSX_REQUIRE(h->batch > 0 && h->tokens > 0 && h->width > 0 && h->width % 8 == 0, "%s: bad shape batch=%d tokens=%d width=%d", what, h->batch, h->tokens, h->width);
  VSX_REQUIRE(h->kind == VSX_HALF_MLP || (h->heads > 0 && h->head_dim > 0), "%s: attention half needs heads / head_dim", what);
  VSX_REQUIRE(h->kind == VSX_HALF_ATTN || (h->hidden > 0 && h->hidden % 8 == 0), "%s: MLP half needs hidden %% 8 == 0", what);
  VSX_REQUIRE(h->num_segments >= 0 && (h->num_segments == 0 || h->segments != nullptr), "%s: bad segment list", what);
  return VSX_OK;
}

}  // namespace

extern "C" int vsx_half_block_fwd(const vsx_half_block* h, void* stream) {
  HB_CHECK(check_desc(h, "vsx_half_block_fwd"));
  const int N = h->tokens, C = h->width, H = h->heads, D = h->head_dim, HD = H * D, F = h->hidden;
  const float scale = h->kind == VSX_HALF_ATTN ? 1.0f / sqrtf((float)D) : 0.f;
  for (int si = 0; si < h->num_segments; ++si) {
    const vsx_segment& s = h->segments[si];
    const long r0 = (long)s.b0 * N;
    const int nb = s.b1 - s.b0, rows = nb * N;
    if (rows <= 0) continue;
    if (!s.active) {
      HB_CHECK(passthrough(h->x + r0 * C, h->out + r0 * C, (long)rows * C, h->residual != 0, stream));
      continue;
    }
    HB_CHECK(pre_norm_or_cast(h, s, r0, rows, stream));
    const bf16* xn = B16(h->xn) + r0 * C;
    if (h->kind == VSX_HALF_ATTN) {
      const int hk = s.inner_keep / D, hkd = hk * D;
      bf16* qkv = B16(h->act1) + r0 * 3 * HD;
      bf16* o = B16(h->act2) + r0 * HD;
      if (hk == H) {
        Gemm g(xn, C, VSX_KMAJOR, h->w1, C, VSX_KMAJOR, rows, 3 * HD, s.embed_keep, VSX_EPI_STORE, VSX_BF16, qkv, 3 * HD);
        g.d.bias = h->b1;
        HB_CHECK(vsx_gemm(&g.d, stream));
      } else {
        // q / k / v row blocks of the kept heads only (features ordered (3,H,D)): three problems, one launch
        vsx_gemm_desc ds[3];
        for (int j = 0; j < 3; ++j) {
          Gemm g(xn, C, VSX_KMAJOR, B16(h->w1) + (long)j * HD * C, C, VSX_KMAJOR, rows, hkd, s.embed_keep, VSX_EPI_STORE, VSX_BF16, qkv + j * HD, 3 * HD);
          g.d.bias = h->b1 != nullptr ? h->b1 + j * HD : nullptr;
          ds[j] = g.d;
        }
        HB_CHECK(vsx_gemm_grouped(ds, 3, stream));
      }
      HB_CHECK(vsx_attn_fwd(qkv, o, h->lse + (long)s.b0 * H * N, VSX_BF16, nb, N, H, D, hk, scale, VSX_ATTN_IMPL_AUTO, stream));
      if (h->residual) {
        Gemm g(o, HD, VSX_KMAJOR, h->w2, HD, VSX_KMAJOR, rows, s.out_keep, hkd, VSX_EPI_RESIDUAL, VSX_F32, h->out + r0 * C, C);
        g.d.n_out = C, g.d.bias = h->b2, g.d.aux = h->x + r0 * C, g.d.ld_aux = C;
        g.d.row_scale = h->row_scale != nullptr ? h->row_scale + h->scale_off + s.b0 : nullptr, g.d.rows_per_sample = N, g.d.n_keep = s.out_keep;
        HB_CHECK(vsx_gemm(&g.d, stream));
      } else {
        Gemm g(o, HD, VSX_KMAJOR, h->w2, HD, VSX_KMAJOR, rows, C, hkd, VSX_EPI_STORE, VSX_F32, h->out + r0 * C, C);
        g.d.bias = h->b2;
        HB_CHECK(vsx_gemm(&g.d, stream));
      }
    } else {
      bf16* u = B16(h->act1) + r0 * F;
      bf16* hh = B16(h->act2) + r0 * F;
      {
        Gemm g(xn, C, VSX_KMAJOR, h->w1, C, VSX_KMAJOR, rows, s.inner_keep, s.embed_keep, VSX_EPI_GELU, VSX_BF16, u, F);
        g.d.n_out = up8(s.inner_keep), g.d.out2 = hh, g.d.ldo2 = F, g.d.bias = h->b1;
        HB_CHECK(vsx_gemm(&g.d, stream));
      }
      if (h->residual) {
        Gemm g(hh, F, VSX_KMAJOR, h->w2, F, VSX_KMAJOR, rows, s.out_keep, s.inner_keep, VSX_EPI_RESIDUAL, VSX_F32, h->out + r0 * C, C);
        g.d.n_out = C, g.d.bias = h->b2, g.d.aux = h->x + r0 * C, g.d.ld_aux = C;
        g.d.row_scale = h->row_scale != nullptr ? h->row_scale + h->scale_off + s.b0 : nullptr, g.d.rows_per_sample = N, g.d.n_keep = s.out_keep;
        HB_CHECK(vsx_gemm(&g.d, stream));
      } else {
        Gemm g(hh, F, VSX_KMAJOR, h->w2, F, VSX_KMAJOR, rows, C, s.inner_keep, VSX_EPI_STORE, VSX_F32, h->out + r0 * C, C);
        g.d.bias = h->b2;
        HB_CHECK(vsx_gemm(&g.d, stream));
      }
    }
  }
  return VSX_OK;
}

extern "C" int vsx_half_block_bwd(const vsx_half_block_grad* b, void* stream) {
  VSX_REQUIRE(b != nullptr, "vsx_half_block_bwd: null descriptor");
  const vsx_half_block* h = &b->fwd;
  HB_CHECK(check_desc(h, "vsx_half_block_bwd"));
  const int N = h->tokens, C = h->width, H = h->heads, D = h->head_dim, HD = H * D, F = h->hidden;
  const float scale = h->kind == VSX_HALF_ATTN ? 1.0f / sqrtf((float)D) : 0.f;
  for (int si = 0; si < h->num_segments; ++si) {
    const vsx_segment& s = h->segments[si];
    const long r0 = (long)s.b0 * N;
    const int nb = s.b1 - s.b0, rows = nb * N;
    if (rows <= 0) continue;
    if (!s.active) {
      HB_CHECK(passthrough(b->g_out + r0 * C, b->g_in + r0 * C, (long)rows * C, h->residual != 0, stream));
      continue;
    }
    const int ck = h->residual ? s.out_keep : C;
    bf16* df = B16(b->df) + r0 * C;
    const bf16* xn = B16(h->xn) + r0 * C;
    bf16* dxn = B16(b->dxn) + r0 * C;
    // gradient of the branch output: drop-path scale, output mask, cast; its column sums are the bias gradient of proj / fc2
    HB_CHECK(vsx_scale_mask_cast(b->g_out + r0 * C, C, (h->residual && h->row_scale != nullptr) ? h->row_scale + h->scale_off + s.b0 : nullptr, N, ck,
                                 df, VSX_BF16, C, rows, C, b->d_b2, stream));
    vsx_gemm_desc wg[4];
    int nwg = 0;
    if (h->kind == VSX_HALF_ATTN) {
      const int hk = s.inner_keep / D, hkd = hk * D;
      const bf16* qkv = B16(h->act1) + r0 * 3 * HD;
      const bf16* o = B16(h->act2) + r0 * HD;
      bf16* d_o = B16(b->d_act2) + r0 * HD;
      bf16* dqkv = B16(b->d_act1) + r0 * 3 * HD;
      {  // dWproj[ck, hkd] += df^T o
        Gemm g(df, C, VSX_MNMAJOR, o, HD, VSX_MNMAJOR, ck, hkd, rows, VSX_EPI_ATOMIC, VSX_F32, b->d_w2, HD);
        g.d.split_k = split_k_for(ck, hkd, rows);
        wg[nwg++] = g.d;
      }
      {  // d_o[rows, hkd] = df[rows, ck] Wproj[ck, hkd]
        Gemm g(df, C, VSX_KMAJOR, h->w2, HD, VSX_MNMAJOR, rows, hkd, ck, VSX_EPI_STORE, VSX_BF16, d_o, HD);
        HB_CHECK(vsx_gemm(&g.d, stream));
      }
      HB_CHECK(vsx_attn_bwd(qkv, o, d_o, h->lse + (long)s.b0 * H * N, dqkv, VSX_BF16, nb, N, H, D, hk, scale, VSX_ATTN_IMPL_AUTO, b->d_b1, stream));
      const int parts = hk < H ? 3 : 1, nrow = hk < H ? hkd : 3 * HD;
      for (int j = 0; j < parts; ++j) {   // dWqkv[j] += dqkv_j^T xn
        Gemm g(dqkv + j * HD, 3 * HD, VSX_MNMAJOR, xn, C, VSX_MNMAJOR, nrow, s.embed_keep, rows, VSX_EPI_ATOMIC, VSX_F32, b->d_w1 + (long)j * HD * C, C);
        g.d.split_k = split_k_for(nrow, s.embed_keep, rows);
        wg[nwg++] = g.d;
      }
      HB_CHECK(vsx_gemm_grouped(wg, nwg, stream));
      // dxn[rows, ek] = dqkv[rows, 3HD] Wqkv[3HD, ek]   (masked heads are zero columns of dqkv)
      if (h->pre_norm) {
        Gemm g(dqkv, 3 * HD, VSX_KMAJOR, h->w1, C, VSX_MNMAJOR, rows, s.embed_keep, 3 * HD, VSX_EPI_STORE, VSX_BF16, dxn, C);
        g.d.n_out = up8(s.embed_keep);
        // reduce over the kept heads only (the 64-wide k steps may overrun a window only into masked, i.e. zero, columns)
        if (hk < H && up64(hkd) <= HD) g.d.k_segments = 3, g.d.k_seg_len = hkd, g.d.k_seg_stride = HD;
        HB_CHECK(vsx_gemm(&g.d, stream));
      } else {
        Gemm g(dqkv, 3 * HD, VSX_KMAJOR, h->w1, C, VSX_MNMAJOR, rows, s.embed_keep, 3 * HD, VSX_EPI_STORE, VSX_F32, b->g_in + r0 * C, C);
        g.d.n_out = C;
        if (hk < H && up64(hkd) <= HD) g.d.k_segments = 3, g.d.k_seg_len = hkd, g.d.k_seg_stride = HD;
        HB_CHECK(vsx_gemm(&g.d, stream));
      }
    } else {
      const bf16* u = B16(h->act1) + r0 * F;
      const bf16* hh = B16(h->act2) + r0 * F;
      bf16* du = B16(b->d_act1) + r0 * F;
      {  // dW2[ck, ik] += df^T h
        Gemm g(df, C, VSX_MNMAJOR, hh, F, VSX_MNMAJOR, ck, s.inner_keep, rows, VSX_EPI_ATOMIC, VSX_F32, b->d_w2, F);
        g.d.split_k = split_k_for(ck, s.inner_keep, rows);
        wg[nwg++] = g.d;
      }
      {  // du[rows, ik] = (df[rows, ck] W2[ck, ik]) * gelu'(u); its column sums are the fc1 bias gradient
        Gemm g(df, C, VSX_KMAJOR, h->w2, F, VSX_MNMAJOR, rows, s.inner_keep, ck, VSX_EPI_GELUGRAD, VSX_BF16, du, F);
        g.d.n_out = up8(s.inner_keep), g.d.aux = u, g.d.ld_aux = F, g.d.colsum = b->d_b1;
        HB_CHECK(vsx_gemm(&g.d, stream));
      }
      {  // dW1[ik, ek] += du^T xn
        Gemm g(du, F, VSX_MNMAJOR, xn, C, VSX_MNMAJOR, s.inner_keep, s.embed_keep, rows, VSX_EPI_ATOMIC, VSX_F32, b->d_w1, C);
        g.d.split_k = split_k_for(s.inner_keep, s.embed_keep, rows);
        wg[nwg++] = g.d;
      }
      HB_CHECK(vsx_gemm_grouped(wg, nwg, stream));
      // dxn[rows, ek] = du[rows, ik] W1[ik, ek]
      if (h->pre_norm) {
        Gemm g(du, F, VSX_KMAJOR, h->w1, C, VSX_MNMAJOR, rows, s.embed_keep, s.inner_keep, VSX_EPI_STORE, VSX_BF16, dxn, C);
        g.d.n_out = up8(s.embed_keep);
        HB_CHECK(vsx_gemm(&g.d, stream));
      } else {
        Gemm g(du, F, VSX_KMAJOR, h->w1, C, VSX_MNMAJOR, rows, s.embed_keep, s.inner_keep, VSX_EPI_STORE, VSX_F32, b->g_in + r0 * C, C);
        g.d.n_out = C;
        HB_CHECK(vsx_gemm(&g.d, stream));
      }
    }
    if (h->pre_norm)
      HB_CHECK(vsx_masked_ln_bwd(dxn, nullptr, VSX_BF16, C, h->x + r0 * C, C, h->mean + r0, h->rstd + r0, h->ln_w, h->residual ? b->g_out + r0 * C : nullptr,
                                 b->g_in + r0 * C, C, b->d_ln_w, b->d_ln_b, rows, C, s.embed_keep, 0, 0, stream));
  }
  return VSX_OK;
}
