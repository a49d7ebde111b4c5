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

// Row-segment table of a half block for the single-launch row kernels: `which` selects the extent (0 embed, 2 out), dropped segments get 0.
// Returns false when the batch has too many segments for one table (the callers then launch per segment).
bool row_segments(const vsx_half_block* h, int which, vsx_row_segments* t) {
  if (h->num_segments < 2 || h->num_segments > VSX_MAX_SEGMENTS) return false;
  memset(t, 0, sizeof(*t));
  t->count = h->num_segments;
  int active = 0;
  for (int i = 0; i < h->num_segments; ++i) {
    const vsx_segment& s = h->segments[i];
    t->row_end[i] = s.b1 * h->tokens;
    const int ext = which == 0 ? s.embed_keep : (h->residual ? s.out_keep : h->width);
    t->keep[i] = (s.active && s.b1 > s.b0) ? ext : 0;
    active += t->keep[i] > 0;
  }
  return active >= 2 && h->segments[0].b0 == 0 && h->segments[h->num_segments - 1].b1 == h->batch;
}

// Sample-segment table (kept heads per segment) for the single-launch attention core; false: launch per segment.
bool sample_segments(const vsx_half_block* h, vsx_sample_segments* t) {
  if (h->num_segments < 2 || h->num_segments > VSX_MAX_SEGMENTS) return false;
  memset(t, 0, sizeof(*t));
  t->count = h->num_segments;
  int active = 0;
  for (int i = 0; i < h->num_segments; ++i) {
    const vsx_segment& s = h->segments[i];
    t->sample_end[i] = s.b1;
    t->heads_keep[i] = (s.active && s.b1 > s.b0) ? s.inner_keep / h->head_dim : 0;
    active += t->heads_keep[i] > 0;
  }
  return active >= 2 && h->segments[0].b0 == 0 && h->segments[h->num_segments - 1].b1 == h->batch;
}

int check_desc(const vsx_half_block* h, const char* what) {
  VSX_REQUIRE(h != nullptr && (h->kind == VSX_HALF_ATTN || h->kind == VSX_HALF_MLP), "%s: bad descriptor", what);
  VSX_REQUIRE(h->batch > 0 && h->tokens > 0 && h->width > 0 && h->width % 8 == 0, "%s: bad shape batch=%d tokens=%d width=%d", what, h->batch, h->tokens, h->width);
  VSX_REQUIRE(h->kind == VSX_HALF_MLP || (h->heads > 0 && h->head_dim > 0), "%s: attention half needs heads / head_dim", what);
  VSX_REQUIRE(h->kind == VSX_HALF_ATTN || (h->hidden > 0 && h->hidden % 8 == 0), "%s: MLP half needs hidden %% 8 == 0", what);
  VSX_REQUIRE(h->num_segments >= 0 && (h->num_segments == 0 || h->segments != nullptr), "%s: bad segment list", what);
  return VSX_OK;
}

// Problems of the same epilogue collected over the segments of a half block and launched together (up to 16 single-term problems per
// launch: four segments x {q, k, v} row blocks, or all weight gradients of a half block): a step with several sub-architectures
// (multi / hybrid sampling) then costs the same number of GEMM launches as a single-architecture step, and the persistent CTAs walk
// ONE tile list.
struct GemmBatch {
  static constexpr int CAP = 16;
  vsx_gemm_desc d[CAP];
  int n = 0;
  void* stream;
  explicit GemmBatch(void* st) : stream(st) {}
  int flush() {
    if (n == 0) return VSX_OK;
    const int rc = n == 1 ? vsx_gemm(&d[0], stream) : vsx_gemm_grouped(d, n, stream);
    n = 0;
    return rc;
  }
  int add(const vsx_gemm_desc& g) {
    d[n++] = g;
    return n == CAP ? flush() : VSX_OK;
  }
};

struct SegView {
  const vsx_segment* s;
  long r0;
  int nb, rows;
};

}  // namespace

extern "C" int vsx_half_block_fwd(const vsx_half_block* h, void* stream) {
  HB_CHECK(check_desc(h, "vsx_half_block_fwd"));
  const int N = h->tokens, C = h->width, H = h->heads, D = h->head_dim, HD = H * D, F = h->hidden;
  const float scale = h->kind == VSX_HALF_ATTN ? 1.0f / sqrtf((float)D) : 0.f;
  // phase 1: dropped layers pass through; LayerNorm (or the bare input mask) of every active segment -- ONE launch over the whole batch
  // with per-row extents when the batch carries several sub-architectures
  vsx_row_segments tab;
  const bool one_launch = row_segments(h, 0, &tab);
  if (one_launch) {
    const int rows_all = h->batch * N;
    if (h->pre_norm)
      HB_CHECK(vsx_masked_ln_fwd_segs(h->x, C, h->ln_w, h->ln_b, h->xn, VSX_BF16, C, h->mean, h->rstd, rows_all, C, &tab, h->eps, stream));
    else
      HB_CHECK(vsx_scale_mask_cast_segs(h->x, C, nullptr, 1, h->xn, VSX_BF16, C, rows_all, C, &tab, nullptr, stream));
  }
  for (int si = 0; si < h->num_segments; ++si) {
    const vsx_segment& s = h->segments[si];
    const long r0 = (long)s.b0 * N;
    const int rows = (s.b1 - s.b0) * N;
    if (rows <= 0) continue;
    if (!s.active) {
      HB_CHECK(passthrough(h->x + r0 * C, h->out + r0 * C, (long)rows * C, h->residual != 0, stream));
      continue;
    }
    if (!one_launch) HB_CHECK(pre_norm_or_cast(h, s, r0, rows, stream));
  }
  auto for_active = [&](auto&& fn) -> int {
    for (int si = 0; si < h->num_segments; ++si) {
      const vsx_segment& s = h->segments[si];
      SegView v{&s, (long)s.b0 * N, s.b1 - s.b0, (s.b1 - s.b0) * N};
      if (v.rows <= 0 || !s.active) continue;
      HB_CHECK(fn(v));
    }
    return VSX_OK;
  };
  GemmBatch batch(stream);
  auto out_gemm = [&](const SegView& v, const bf16* a, long lda, int kdim) {   // proj / fc2 with the residual tail (or a plain store)
    const vsx_segment& s = *v.s;
    if (h->residual) {
      Gemm g(a, lda, VSX_KMAJOR, h->w2, lda, VSX_KMAJOR, v.rows, s.out_keep, kdim, VSX_EPI_RESIDUAL, VSX_F32, h->out + v.r0 * C, C);
      g.d.n_out = C, g.d.bias = h->b2, g.d.aux = h->x + v.r0 * C, g.d.ld_aux = C;
      g.d.row_scale = h->row_scale != nullptr ? h->row_scale + h->scale_off + s.b0 : nullptr, g.d.rows_per_sample = N, g.d.n_keep = s.out_keep;
      return batch.add(g.d);
    }
    Gemm g(a, lda, VSX_KMAJOR, h->w2, lda, VSX_KMAJOR, v.rows, C, kdim, VSX_EPI_STORE, VSX_F32, h->out + v.r0 * C, C);
    g.d.bias = h->b2;
    return batch.add(g.d);
  };
  if (h->kind == VSX_HALF_ATTN) {
    // phase 2: qkv projections -- one problem per segment, or the q / k / v row blocks of the kept heads (features ordered (3,H,D))
    HB_CHECK(for_active([&](const SegView& v) -> int {
      const vsx_segment& s = *v.s;
      const int hk = s.inner_keep / D, hkd = hk * D;
      const bf16* xn = B16(h->xn) + v.r0 * C;
      bf16* qkv = B16(h->act1) + v.r0 * 3 * HD;
      const int parts = hk == H ? 1 : 3;
      for (int j = 0; j < parts; ++j) {
        Gemm g(xn, C, VSX_KMAJOR, B16(h->w1) + (long)j * HD * C, C, VSX_KMAJOR, v.rows, hk == H ? 3 * HD : hkd, s.embed_keep, VSX_EPI_STORE, VSX_BF16,
               qkv + j * HD, 3 * HD);
        g.d.bias = h->b1 != nullptr ? h->b1 + j * HD : nullptr;
        HB_CHECK(batch.add(g.d));
      }
      return VSX_OK;
    }));
    HB_CHECK(batch.flush());
    // phase 3: attention core: one launch over the batch with per-segment kept heads, or one per segment
    vsx_sample_segments stab;
    if (sample_segments(h, &stab)) {
      HB_CHECK(vsx_attn_fwd_segs(h->act1, h->act2, h->lse, VSX_BF16, h->batch, N, H, D, &stab, scale, VSX_ATTN_IMPL_AUTO, stream));
    } else
    HB_CHECK(for_active([&](const SegView& v) -> int {
      return vsx_attn_fwd(B16(h->act1) + v.r0 * 3 * HD, B16(h->act2) + v.r0 * HD, h->lse + (long)v.s->b0 * H * N, VSX_BF16, v.nb, N, H, D,
                          v.s->inner_keep / D, scale, VSX_ATTN_IMPL_AUTO, stream);
    }));
    // phase 4: output projections
    HB_CHECK(for_active([&](const SegView& v) -> int { return out_gemm(v, B16(h->act2) + v.r0 * HD, HD, (v.s->inner_keep / D) * D); }));
    HB_CHECK(batch.flush());
  } else {
    HB_CHECK(for_active([&](const SegView& v) -> int {
      const vsx_segment& s = *v.s;
      Gemm g(B16(h->xn) + v.r0 * C, C, VSX_KMAJOR, h->w1, C, VSX_KMAJOR, v.rows, s.inner_keep, s.embed_keep, VSX_EPI_GELU, VSX_BF16,
             B16(h->act1) + v.r0 * F, F);
      g.d.n_out = up8(s.inner_keep), g.d.out2 = B16(h->act2) + v.r0 * F, g.d.ldo2 = F, g.d.bias = h->b1;
      return batch.add(g.d);
    }));
    HB_CHECK(batch.flush());
    HB_CHECK(for_active([&](const SegView& v) -> int { return out_gemm(v, B16(h->act2) + v.r0 * F, F, v.s->inner_keep); }));
    HB_CHECK(batch.flush());
  }
  return VSX_OK;
}

extern "C" int vsx_half_block_bwd(const vsx_half_block_grad* b, void* stream) {
  VSX_REQUIRE(b != nullptr, "vsx_half_block_bwd: null descriptor");
  const vsx_half_block* h = &b->fwd;
  HB_CHECK(check_desc(h, "vsx_half_block_bwd"));
  const int N = h->tokens, C = h->width, H = h->heads, D = h->head_dim, HD = H * D, F = h->hidden;
  const float scale = h->kind == VSX_HALF_ATTN ? 1.0f / sqrtf((float)D) : 0.f;
  if (b->df_ready || b->next_df != nullptr) {
    bool ok = h->pre_norm && h->residual && h->num_segments >= 1 && h->num_segments <= VSX_MAX_SEGMENTS && h->segments[0].b0 == 0 &&
              h->segments[h->num_segments - 1].b1 == h->batch && (h->num_segments == 1 || b->next_df == nullptr || b->next_segments != nullptr);
    for (int i = 0; ok && i < h->num_segments; ++i) ok = h->segments[i].b1 > h->segments[i].b0;
    VSX_REQUIRE(ok, "vsx_half_block_bwd: df_ready / next_df need pre_norm, residual and non-empty segments covering the batch (next_segments for several)");
  }
  // extent of the CONSUMING half block's cast for this call's segment si (0: the consumer drops its layer for these samples)
  auto next_keep_of = [&](int si) -> int {
    if (b->next_segments == nullptr) return b->next_keep;
    return b->next_segments[si].active ? b->next_segments[si].out_keep : 0;
  };
  // phase 1: dropped layers pass the gradient through; gradient of the branch output of every active segment: drop-path scale,
  // output mask, cast -- its column sums are the bias gradient of proj / fc2
  vsx_row_segments tab_out, tab_in;
  const bool one_cast = !b->df_ready && row_segments(h, 2, &tab_out);
  const bool one_ln = h->pre_norm && row_segments(h, 0, &tab_in);
  if (one_cast)
    HB_CHECK(vsx_scale_mask_cast_segs(b->g_out, C, (h->residual && h->row_scale != nullptr) ? h->row_scale + h->scale_off : nullptr, N, b->df, VSX_BF16, C,
                                      h->batch * N, C, &tab_out, b->d_b2, stream));
  for (int si = 0; si < h->num_segments; ++si) {
    const vsx_segment& s = h->segments[si];
    const long r0 = (long)s.b0 * N;
    const int rows = (s.b1 - s.b0) * N;
    if (rows <= 0) continue;
    if (!s.active) {
      HB_CHECK(passthrough(b->g_out + r0 * C, b->g_in + r0 * C, (long)rows * C, h->residual != 0, stream));
      // no LayerNorm backward runs on these rows: the cast the consuming half block starts from is made here, from the passed-through rows
      if (b->next_df != nullptr && next_keep_of(si) > 0)
        HB_CHECK(vsx_scale_mask_cast(b->g_out + r0 * C, C, b->next_row_scale != nullptr ? b->next_row_scale + b->next_scale_off + s.b0 : nullptr, N,
                                     next_keep_of(si), B16(b->next_df) + r0 * C, VSX_BF16, C, rows, C, b->next_d_b2, stream));
      continue;
    }
    const int ck = h->residual ? s.out_keep : C;
    if (b->df_ready || one_cast) continue;   // written by the LayerNorm backward of the call that produced g_out / by the launch above
    HB_CHECK(vsx_scale_mask_cast(b->g_out + r0 * C, C, (h->residual && h->row_scale != nullptr) ? h->row_scale + h->scale_off + s.b0 : nullptr, N, ck,
                                 B16(b->df) + r0 * C, VSX_BF16, C, rows, C, b->d_b2, stream));
  }
  auto for_active = [&](auto&& fn) -> int {
    for (int si = 0; si < h->num_segments; ++si) {
      const vsx_segment& s = h->segments[si];
      SegView v{&s, (long)s.b0 * N, s.b1 - s.b0, (s.b1 - s.b0) * N};
      if (v.rows <= 0 || !s.active) continue;
      HB_CHECK(fn(v));
    }
    return VSX_OK;
  };
  GemmBatch batch(stream);
  auto dxn_gemm = [&](const SegView& v, const bf16* a, long lda, int kdim, int hk) {     // gradient w.r.t. the (normalised) branch input
    const vsx_segment& s = *v.s;
    const int odt = h->pre_norm ? VSX_BF16 : VSX_F32;
    void* out = h->pre_norm ? static_cast<void*>(B16(b->dxn) + v.r0 * C) : static_cast<void*>(b->g_in + v.r0 * C);
    Gemm g(a, lda, VSX_KMAJOR, h->w1, C, VSX_MNMAJOR, v.rows, s.embed_keep, kdim, VSX_EPI_STORE, odt, out, C);
    g.d.n_out = h->pre_norm ? up8(s.embed_keep) : C;
    // attention: reduce over the kept-head windows only (the 64-wide k steps may overrun a window only into masked, i.e. zero, columns)
    if (hk >= 0 && hk < H && up64(hk * D) <= HD) g.d.k_segments = 3, g.d.k_seg_len = hk * D, g.d.k_seg_stride = HD;
    return batch.add(g.d);
  };
  if (h->kind == VSX_HALF_ATTN) {
    // phase 2: d_o[rows, hkd] = df[rows, ck] Wproj[ck, hkd]
    HB_CHECK(for_active([&](const SegView& v) -> int {
      const int ck = h->residual ? v.s->out_keep : C, hkd = (v.s->inner_keep / D) * D;
      Gemm g(B16(b->df) + v.r0 * C, C, VSX_KMAJOR, h->w2, HD, VSX_MNMAJOR, v.rows, hkd, ck, VSX_EPI_STORE, VSX_BF16, B16(b->d_act2) + v.r0 * HD, HD);
      return batch.add(g.d);
    }));
    HB_CHECK(batch.flush());
    // phase 3: attention backward (also accumulates the qkv bias gradient): one launch over the batch, or one per segment
    vsx_sample_segments stab;
    if (sample_segments(h, &stab)) {
      HB_CHECK(vsx_attn_bwd_segs(h->act1, h->act2, b->d_act2, h->lse, b->d_act1, VSX_BF16, h->batch, N, H, D, &stab, scale, VSX_ATTN_IMPL_AUTO, b->d_b1, stream));
    } else
    HB_CHECK(for_active([&](const SegView& v) -> int {
      return vsx_attn_bwd(B16(h->act1) + v.r0 * 3 * HD, B16(h->act2) + v.r0 * HD, B16(b->d_act2) + v.r0 * HD, h->lse + (long)v.s->b0 * H * N,
                          B16(b->d_act1) + v.r0 * 3 * HD, VSX_BF16, v.nb, N, H, D, v.s->inner_keep / D, scale, VSX_ATTN_IMPL_AUTO, b->d_b1, stream);
    }));
    // phase 4: weight gradients: dWproj[ck, hkd] += df^T o and dWqkv[j] += dqkv_j^T xn
    HB_CHECK(for_active([&](const SegView& v) -> int {
      const vsx_segment& s = *v.s;
      const int ck = h->residual ? s.out_keep : C, hk = s.inner_keep / D, hkd = hk * D;
      {
        Gemm g(B16(b->df) + v.r0 * C, C, VSX_MNMAJOR, B16(h->act2) + v.r0 * HD, HD, VSX_MNMAJOR, ck, hkd, v.rows, VSX_EPI_ATOMIC, VSX_F32, b->d_w2, HD);
        g.d.split_k = split_k_for(ck, hkd, v.rows);
        HB_CHECK(batch.add(g.d));
      }
      const int parts = hk < H ? 3 : 1, nrow = hk < H ? hkd : 3 * HD;
      for (int j = 0; j < parts; ++j) {
        Gemm g(B16(b->d_act1) + v.r0 * 3 * HD + j * HD, 3 * HD, VSX_MNMAJOR, B16(h->xn) + v.r0 * C, C, VSX_MNMAJOR, nrow, s.embed_keep, v.rows,
               VSX_EPI_ATOMIC, VSX_F32, b->d_w1 + (long)j * HD * C, C);
        g.d.split_k = split_k_for(nrow, s.embed_keep, v.rows);
        HB_CHECK(batch.add(g.d));
      }
      return VSX_OK;
    }));
    HB_CHECK(batch.flush());
    // phase 5: dxn[rows, ek] = dqkv[rows, 3HD] Wqkv[3HD, ek]
    HB_CHECK(for_active([&](const SegView& v) -> int { return dxn_gemm(v, B16(b->d_act1) + v.r0 * 3 * HD, 3 * HD, 3 * HD, v.s->inner_keep / D); }));
    HB_CHECK(batch.flush());
  } else {
    // phase 2: du[rows, ik] = (df[rows, ck] W2[ck, ik]) * gelu'(u); its column sums are the fc1 bias gradient (all segments add into
    // the same vector: per-CTA shared-memory partials also for the grouped launch)
    HB_CHECK(for_active([&](const SegView& v) -> int {
      const vsx_segment& s = *v.s;
      const int ck = h->residual ? s.out_keep : C;
      Gemm g(B16(b->df) + v.r0 * C, C, VSX_KMAJOR, h->w2, F, VSX_MNMAJOR, v.rows, s.inner_keep, ck, VSX_EPI_GELUGRAD, VSX_BF16, B16(b->d_act1) + v.r0 * F, F);
      g.d.n_out = up8(s.inner_keep), g.d.aux = B16(h->act1) + v.r0 * F, g.d.ld_aux = F, g.d.colsum = b->d_b1;
      return batch.add(g.d);
    }));
    HB_CHECK(batch.flush());
    // phase 3: weight gradients: dW2[ck, ik] += df^T h and dW1[ik, ek] += du^T xn
    HB_CHECK(for_active([&](const SegView& v) -> int {
      const vsx_segment& s = *v.s;
      const int ck = h->residual ? s.out_keep : C;
      {
        Gemm g(B16(b->df) + v.r0 * C, C, VSX_MNMAJOR, B16(h->act2) + v.r0 * F, F, VSX_MNMAJOR, ck, s.inner_keep, v.rows, VSX_EPI_ATOMIC, VSX_F32, b->d_w2, F);
        g.d.split_k = split_k_for(ck, s.inner_keep, v.rows);
        HB_CHECK(batch.add(g.d));
      }
      Gemm g(B16(b->d_act1) + v.r0 * F, F, VSX_MNMAJOR, B16(h->xn) + v.r0 * C, C, VSX_MNMAJOR, s.inner_keep, s.embed_keep, v.rows, VSX_EPI_ATOMIC, VSX_F32,
             b->d_w1, C);
      g.d.split_k = split_k_for(s.inner_keep, s.embed_keep, v.rows);
      return batch.add(g.d);
    }));
    HB_CHECK(batch.flush());
    // phase 4: dxn[rows, ek] = du[rows, ik] W1[ik, ek]
    HB_CHECK(for_active([&](const SegView& v) -> int { return dxn_gemm(v, B16(b->d_act1) + v.r0 * F, F, v.s->inner_keep, -1); }));
    HB_CHECK(batch.flush());
  }
  if (one_ln) {
    if (b->next_df != nullptr) {        // the cast of the consuming half block rides on this LayerNorm backward, segment by segment
      for (int i = 0; i < h->num_segments; ++i) tab_in.keep2[i] = next_keep_of(i);
      HB_CHECK(vsx_masked_ln_bwd_segs(b->dxn, VSX_BF16, C, h->x, C, h->mean, h->rstd, h->ln_w, h->residual ? b->g_out : nullptr, b->g_in, C, b->d_ln_w,
                                      b->d_ln_b, h->batch * N, C, &tab_in, b->next_df, C,
                                      b->next_row_scale != nullptr ? b->next_row_scale + b->next_scale_off : nullptr, N, b->next_d_b2, stream));
    } else
    HB_CHECK(vsx_masked_ln_bwd_segs(b->dxn, VSX_BF16, C, h->x, C, h->mean, h->rstd, h->ln_w, h->residual ? b->g_out : nullptr, b->g_in, C, b->d_ln_w, b->d_ln_b,
                                    h->batch * N, C, &tab_in, nullptr, C, nullptr, 0, nullptr, stream));
  } else if (h->pre_norm) {
    HB_CHECK(for_active([&](const SegView& v) -> int {
      if (b->next_df != nullptr && next_keep_of((int)(v.s - h->segments)) > 0)
        return vsx_masked_ln_bwd_cast(B16(b->dxn) + v.r0 * C, VSX_BF16, C, h->x + v.r0 * C, C, h->mean + v.r0, h->rstd + v.r0, h->ln_w,
                                      h->residual ? b->g_out + v.r0 * C : nullptr, b->g_in + v.r0 * C, C, b->d_ln_w, b->d_ln_b, v.rows, C,
                                      v.s->embed_keep, B16(b->next_df) + v.r0 * C, C,
                                      b->next_row_scale != nullptr ? b->next_row_scale + b->next_scale_off + v.s->b0 : nullptr, N,
                                      next_keep_of((int)(v.s - h->segments)), b->next_d_b2, stream);
      return vsx_masked_ln_bwd(B16(b->dxn) + v.r0 * C, nullptr, VSX_BF16, C, h->x + v.r0 * C, C, h->mean + v.r0, h->rstd + v.r0, h->ln_w,
                               h->residual ? b->g_out + v.r0 * C : nullptr, b->g_in + v.r0 * C, C, b->d_ln_w, b->d_ln_b, v.rows, C, v.s->embed_keep, 0, 0,
                               stream);
    }));
  }
  return VSX_OK;
}

// A run of consecutive half blocks (all transformer blocks of a stage) in ONE C-ABI call per direction: the launching thread pays one
// foreign call, one workspace allocation and one autograd node per stage instead of one per half block (the Python side of a train step
// was 12.9 ms against 14.0 ms of GPU work; tools/host_vs_gpu.py).  Descriptors are in forward order for both directions; the
// backward walks them from the last to the first, and the LayerNorm backward of half block i writes the bf16 gradient copy half block
// i - 1 starts from (next_df / df_ready links prepared by the caller).
extern "C" int vsx_stage_fwd(const vsx_half_block* halves, int count, void* stream) {
  VSX_REQUIRE(halves != nullptr && count >= 1, "vsx_stage_fwd: need at least one half block");
  for (int i = 0; i < count; ++i) HB_CHECK(vsx_half_block_fwd(&halves[i], stream));
  return VSX_OK;
}

extern "C" int vsx_stage_bwd(const vsx_half_block_grad* halves, int count, void* stream) {
  VSX_REQUIRE(halves != nullptr && count >= 1, "vsx_stage_bwd: need at least one half block");
  for (int i = count - 1; i >= 0; --i) HB_CHECK(vsx_half_block_bwd(&halves[i], stream));
  return VSX_OK;
}
