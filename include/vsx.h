/*
 * vsx.h -- C ABI of libvsx.so: the sm_100a kernels behind the ViT-Res super-network training hot path.
 *
 * The reference (yilunliao/vit-search) is pure Python on PyTorch and has NO operator/FFI layer of its own:
 * every device op is an ATen / cuBLAS / cuDNN call made from nets/*.py.  This header is therefore the
 * boundary a maintainer would bind (ctypes stub in INTEGRATION.md); each entry point names the reference
 * call site it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success or a negative VSX_ERR_* code; vsx_last_error() gives the text
 *     (thread-local).  Nothing throws, allocates device memory, or synchronises the device.
 *   - all pointers are DEVICE pointers unless named *_host; `stream` is a cudaStream_t passed as void*.
 *   - activation storage type `dtype`: VSX_BF16 (training path) or VSX_F32 (high-precision parity path);
 *     parameters, the residual stream, statistics, gradients of parameters and logits are always fp32.
 *   - a call covers one SEGMENT: a run of consecutive samples that share one sub-architecture, i.e. the same
 *     prefix keep-counts (every mask the reference can draw is a prefix mask, nets/channel_drop.py:153-157).
 *     Callers loop over segments; pointers are pre-offset to the segment's first row.
 *   - `ld*` are row pitches in ELEMENTS.
 */
#ifndef VSX_H_
#define VSX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSX_OK 0
#define VSX_ERR_ARG (-1)     /* invalid argument / unsupported shape */
#define VSX_ERR_CUDA (-2)    /* CUDA runtime / driver error at launch */
#define VSX_ERR_NO_GPU (-3)  /* no sm_100 device */

#define VSX_BF16 0
#define VSX_F32 1

#define VSX_ABI_VERSION 4

const char* vsx_last_error(void);
int vsx_abi_version(void);
/* 0 if device `dev` is an sm_100 part, VSX_ERR_NO_GPU otherwise. */
int vsx_device_ok(int dev);

/* ----------------------------------------------------------------------------------------------------
 * Masked LayerNorm -- replaces MaskedLayerNormFunc.forward/backward + the `x * mask` re-mask
 * (nets/masked_layer_norm.py:23-50, :55-88, :113-125) and F.layer_norm for keep == C (:121).
 * Statistics over the first `keep` channels only; y[:, keep:] = 0.
 * Row remap (final norm feeding cls_head / patch_head, nets/vit_sr_supernet.py:420-424): when
 * rows_per_sample > 0, input row r = b*rows_per_sample + t is written to `y` row b*split_tokens + t when
 * t < split_tokens, else to `y2` row b*(rows_per_sample-split_tokens) + (t-split_tokens).  Pass 0 / NULL
 * for the plain row-to-row mapping.  mean/rstd are indexed by the input row.
 * -------------------------------------------------------------------------------------------------- */
int vsx_masked_ln_fwd(const float* x, long ldx, const float* gamma, const float* beta, void* y, void* y2, int dtype,
                      long ldy, float* mean, float* rstd, int rows, int C, int keep, float eps, int rows_per_sample,
                      int split_tokens, void* stream);
/* g_out = (g_in ? g_in : 0) + dx on channels < keep; channels >= keep copy g_in (or 0).
 * dgamma/dbeta (fp32 [C]) are ACCUMULATED into (atomic adds); zero them once per step. */
int vsx_masked_ln_bwd(const void* dy, const void* dy2, int dtype, long lddy, const float* x, long ldx,
                      const float* mean, const float* rstd, const float* gamma, const float* g_in, float* g_out,
                      long ldg, float* dgamma, float* dbeta, int rows, int C, int keep, int rows_per_sample,
                      int split_tokens, void* stream);
/* vsx_masked_ln_bwd (no row remap) that ALSO writes cast_out = T(cast_scale[row / cast_rows_per_sample] * g_out) masked to the first
 * cast_keep channels (zeros beyond) and accumulates its column sums into cast_colsum (may be NULL): the first step of the backward of
 * the half block that consumes g_out next (vsx_scale_mask_cast fused into this kernel's store).  cast_scale may be NULL (1.0). */
int vsx_masked_ln_bwd_cast(const void* dy, int dtype, long lddy, const float* x, long ldx, const float* mean, const float* rstd,
                           const float* gamma, const float* g_in, float* g_out, long ldg, float* dgamma, float* dbeta, int rows, int C,
                           int keep, void* cast_out, long ld_cast, const float* cast_scale, int cast_rows_per_sample, int cast_keep,
                           float* cast_colsum, void* stream);

/* ----------------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05 / TMEM / TMA) -- replaces every nn.Linear on the path and its autograd:
 * Attention.qkv / .proj (nets/supernet_blocks.py:102,118), Mlp.fc1 / .fc2 (:38,:50), the SR conv as an
 * implicit GEMM and token Linear (nets/vit_sr_supernet.py:142,150), conv_proj (nets/patch_conv.py:72),
 * cls_head / patch_head (nets/vit_sr_supernet.py:440,446), plus the fused elementwise tails around them
 * (bias, GELU :39, drop-path nets/drop.py:11-26, branch mask and residual add :238-253).
 *
 *   D[M,N] = epilogue( sum_{t<terms} A_t[M,K] * B_t[N,K]^T )      bf16 operands, fp32 accumulate
 *
 * terms > 1 is the split-bf16 high-precision mode: every fp32 operand is written as a sum of bf16 parts
 * (x = x1 + x2 [+ x3], vsx_split_bf16) and the significant cross products accumulate into the same TMEM tile:
 * 3 terms (x1y1, x2y1, x1y2; ~2^-16 relative) or 6 terms (+ x2y2, x3y1, x1y3; ~2^-24, i.e. fp32-exact).  The caller
 * lists the operand pair of each term.
 * Layout of an operand: VSX_KMAJOR  -- stored [M or N rows, K contiguous]      (forward: x and W)
 *                       VSX_MNMAJOR -- stored [K rows, M or N contiguous]      (dgrad: W; wgrad: dY and x)
 * -------------------------------------------------------------------------------------------------- */
#define VSX_KMAJOR 0
#define VSX_MNMAJOR 1

#define VSX_EPI_STORE 0     /* out = acc + bias                                   (out: bf16|f32)            */
#define VSX_EPI_GELU 1      /* u = acc + bias; out = gelu'(u), out2 = gelu(u)     (bf16|f32; u itself is not stored) */
#define VSX_EPI_RESIDUAL 2  /* out = aux + [n < n_keep] * row_scale[sample] * (acc + bias)   (f32, aux f32)  */
#define VSX_EPI_GELUGRAD 3  /* out = acc * aux, aux = the gelu'(u) VSX_EPI_GELU stored (out, aux: bf16|f32)  */
#define VSX_EPI_ATOMIC 4    /* out += acc (fp32 atomics; weight gradients, split-K over the reduction)      */

typedef struct vsx_gemm_desc {
  const void* a[6]; /* bf16 A operand per term */
  const void* b[6]; /* bf16 B operand per term */
  int terms;        /* 1 .. 6 */
  long lda, ldb;
  int a_layout, b_layout;
  int M, N, K;      /* N, K are the COMPUTED extents (kept prefix); reads beyond them are zero-filled by TMA */
  int epilogue;
  int out_dtype;    /* VSX_BF16 | VSX_F32 */
  void* out;
  long ldo;
  void* out2;
  long ldo2;
  int n_out;        /* columns written, N <= n_out <= ldo: [N, n_out) is zero-filled (RESIDUAL: copied from aux) */
  const float* bias; /* [>= N] or NULL */
  const void* aux;
  long ld_aux;
  const float* row_scale; /* RESIDUAL: per-sample scale (drop-path keep / keep_prob x layer mask) or NULL (= 1) */
  int rows_per_sample;    /* RESIDUAL: rows of one sample (tokens)                                           */
  int n_keep;             /* RESIDUAL: columns < n_keep receive the branch                                    */
  int split_k;            /* ATOMIC: number of reduction splits (>= 1)                                        */
  float* colsum;          /* GELUGRAD (or STORE without bias): colsum[n] += sum_m out[m,n] (bias gradient), or NULL  */
  int k_segments;         /* 0 / 1: the reduction runs over [0, K).  s > 1: over s windows [j*k_seg_stride, j*k_seg_stride + k_seg_len),  */
  int k_seg_len;          /* j < s, all inside [0, K) -- the kept heads of a (3, heads, head_dim)-ordered qkv gradient: masked heads   */
  int k_seg_stride;       /* are skipped instead of multiplied as zeros.  Windows are walked in 64-element steps.                      */
} vsx_gemm_desc;

int vsx_gemm(const vsx_gemm_desc* d, void* stream);
/* Up to 16 independent problems with the same epilogue and output dtype in ONE launch (the q / k / v row blocks of a head-masked
 * qkv projection for every segment of a multi-architecture batch; all weight gradients of a half block): their tiles form one work
 * list for the persistent CTAs.  More than 4 problems per launch: single-term (plain bf16) problems only. */
int vsx_gemm_grouped(const vsx_gemm_desc* descs, int count, void* stream);
/* CTA tile rows: 0 = heuristic (256-row tiles sharing one B box per k block when they fill the machine), 128 / 256 = forced (tests). */
int vsx_gemm_force_tile_rows(int rows);
/* CTA group: 0 = heuristic, 1 = single-CTA tiles only, 2 = every launch on CTA pairs (tcgen05.mma.cta_group::2: a cluster of two SMs
 * computes a 256 x 256 tile, each CTA staging half of each operand).  Tests and A/B measurements. */
int vsx_gemm_force_cta_group(int cta_group);
/* Development aid: device buffer of 64 x 8 int64 clock stamps written by CTA 0 of the next GEMM launches (NULL = off). */
int vsx_gemm_debug_buffer(void* buffer);

/* ----------------------------------------------------------------------------------------------------
 * Attention core -- replaces q@k^T*scale, softmax, attn@v, the head transpose/reshape and the head ChannelDrop
 * (nets/supernet_blocks.py:103-112) and their autograd.  qkv: [batch*tokens, 3*heads*head_dim] with features
 * ordered (3, heads, head_dim) (:102); o: [batch*tokens, heads*head_dim]; lse: fp32 [batch, heads, tokens]
 * (log-sum-exp of the scaled scores, saved for the backward).  Heads >= heads_keep are not computed: their o /
 * dqkv slices are written as zeros.  impl: VSX_ATTN_IMPL_AUTO picks the fastest kernel that supports the shape:
 * the tcgen05/TMEM kernel (head_dim 64, tokens <= 288), else the mma.sync kernel (head_dim 32/48/64), else fp32 math;
 * VSX_ATTN_IMPL_FP32 forces the fp32-math kernel (always used for dtype VSX_F32), VSX_ATTN_IMPL_MMA_SYNC the legacy
 * tensor-core kernel, VSX_ATTN_IMPL_TCGEN05 the tcgen05 kernel (error when the shape is unsupported).
 * -------------------------------------------------------------------------------------------------- */
#define VSX_ATTN_IMPL_AUTO 0
#define VSX_ATTN_IMPL_FP32 1
#define VSX_ATTN_IMPL_MMA_SYNC 2
#define VSX_ATTN_IMPL_TCGEN05 3
int vsx_attn_fwd(const void* qkv, void* o, float* lse, int dtype, int batch, int tokens, int heads, int head_dim,
                 int heads_keep, float scale, int impl, void* stream);
int vsx_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int dtype, int batch,
                 int tokens, int heads, int head_dim, int heads_keep, float scale, int impl,
                 float* dbias /* NULL, or [3*heads*head_dim]: += column sums of dqkv (the qkv bias gradient) */, void* stream);
/* Development aid: device buffer of (64 x 16 + 3 x 160) int64 stamps written by the tcgen05 backward kernel (NULL = off): clock64 stamps of
 * every warp role of one CTA (VSX_ATTN_DBG_CTA, default 0) and globaltimer start / end / SM id of every CTA (tools/attn_timeline.py). */
int vsx_attn_debug_buffer(void* buffer);
/* Development aid (tests, A-B measurements): which tcgen05 launches run the LAST token (the class token of N = 2^k + 1 tokens) on the
 * CUDA-core side warps instead of a query tile / key block of its own (csrc/attn_tc.cu, "the odd token").  forward: 0 / 1 (N = 128 k + 1 > 128);
 * backward_large / backward_small (N = 64 k + 1 above / up to 128 tokens): bit 0 = as a query, bit 1 = as a key.  A negative value restores
 * the built-in choice (1, 3, 0).  Every combination computes the same result. */
int vsx_attn_odd_token_modes(int forward, int backward_large, int backward_small);

/* ----------------------------------------------------------------------------------------------------
 * One half of a transformer Block in one call (bf16 training path) -- replaces Block.forward's
 *     x = x + mask * drop_path(attn(norm1(x)))      /      x = x + mask * drop_path(mlp(norm2(x)))
 * (nets/supernet_blocks.py:213-253 with Attention.forward :100-120 and Mlp.forward :37-52) and its autograd.
 * The call only sequences the kernels declared in this header, once per SEGMENT: a run of consecutive samples [b0, b1) that
 * share one sub-architecture in this layer.  embed_keep: kept input (embedding) channels; inner_keep: kept heads * head_dim
 * (attention) or hidden channels (MLP); out_keep: channels of the branch output that reach the residual (layer & embed
 * mask); active == 0: the block is dropped for these samples (x passes through).  All buffers are caller-allocated with FULL
 * widths as pitches ([batch*tokens, width | 3*heads*head_dim | heads*head_dim | hidden]); only kept prefixes are touched.
 * -------------------------------------------------------------------------------------------------- */
#define VSX_HALF_ATTN 0
#define VSX_HALF_MLP 1
typedef struct vsx_segment {
  int b0, b1, embed_keep, inner_keep, out_keep, active;
} vsx_segment;
typedef struct vsx_half_block {
  int kind;                  /* VSX_HALF_ATTN | VSX_HALF_MLP */
  int batch, tokens, width;
  int heads, head_dim;       /* attention half */
  int hidden;                /* MLP half */
  int pre_norm, residual;    /* 0: no LayerNorm in front (input is only masked) / no residual add (SR token transform style use) */
  float eps;
  int num_segments;
  const vsx_segment* segments;   /* host memory */
  const float* x;            /* [batch*tokens, width] fp32 residual stream in */
  float* out;                /* [batch*tokens, width] fp32 residual stream out */
  const float *ln_w, *ln_b;
  const void *w1, *w2;       /* bf16 operand copies: qkv [3*H*D, width] / fc1 [hidden, width];  proj [width, H*D] / fc2 [width, hidden] */
  const float *b1, *b2;
  const float* row_scale;    /* NULL or the per-sample drop-path scale table; entry scale_off + sample is used */
  int scale_off;
  void* xn;                  /* bf16 [batch*tokens, width]: LayerNorm output (saved for the backward) */
  float *mean, *rstd;        /* [batch*tokens] */
  void* act1;                /* bf16: qkv [.., 3*H*D]  | fc1 pre-activation u [.., hidden] */
  void* act2;                /* bf16: attention output o [.., H*D] | gelu(u) [.., hidden] */
  float* lse;                /* attention: [batch, heads, tokens] */
} vsx_half_block;
typedef struct vsx_half_block_grad {
  vsx_half_block fwd;        /* the forward call's descriptor (`out` is not used) */
  const float* g_out;        /* [batch*tokens, width] fp32 gradient of `out` */
  float* g_in;               /* [batch*tokens, width] fp32 gradient of `x` */
  void *df, *dxn;            /* bf16 scratch [batch*tokens, width] */
  void* d_act1;              /* bf16 scratch: dqkv [.., 3*H*D] | du [.., hidden] */
  void* d_act2;              /* bf16 scratch: d_o [.., H*D] (attention only) */
  float *d_ln_w, *d_ln_b, *d_w1, *d_b1, *d_w2, *d_b2;   /* fp32 parameter gradients, ACCUMULATED into (zero them once per step) */
  /* Optional fusion across consecutive half blocks (pre-norm, residual calls; all zero / NULL = off).  Segments that drop the layer
   * (active == 0) take part: a dropped segment of THIS call gets its rows of next_df from a cast of the passed-through gradient, a
   * dropped segment of the consuming call is left untouched.
   * df_ready != 0: `df` already holds the scaled / masked gradient of the branch output and d_b2 its column sums -- they were written by
   * the call that produced g_out (its next_* fields) -- so the cast pass over g_out is skipped.
   * next_df != NULL: the LayerNorm backward of THIS call also writes next_df = bf16(next_row_scale[next_scale_off + sample] * g_in) masked
   * to next_keep channels and accumulates its column sums into next_d_b2: the `df` / `d_b2` of the half block that consumes g_in.
   * Several segments: next_segments points to the consuming half block's segment list (same count and sample ranges as this call's);
   * segment i is then masked to next_segments[i].out_keep channels and next_keep is ignored. */
  int df_ready;
  void* next_df;
  const float* next_row_scale;
  int next_scale_off, next_keep;
  float* next_d_b2;
  const vsx_segment* next_segments;
} vsx_half_block_grad;
int vsx_half_block_fwd(const vsx_half_block* d, void* stream);
int vsx_half_block_bwd(const vsx_half_block_grad* d, void* stream);
/* All half blocks of a stage (consecutive transformer blocks between two spatial reductions: the loop at nets/vit_sr_supernet.py:
 * 431-441) in one call per direction.  `halves` is in forward order for both; vsx_stage_bwd walks it from the last entry to the
 * first (g_out of entry i must be g_in of entry i + 1). */
int vsx_stage_fwd(const vsx_half_block* halves, int count, void* stream);
int vsx_stage_bwd(const vsx_half_block_grad* halves, int count, void* stream);
/* Kernels launched by this thread through the library so far (bench.py's gpu_launches). */
long vsx_launch_count(void);

/* ----------------------------------------------------------------------------------------------------
 * Elementwise helpers around the GEMMs.
 * vsx_split_bf16     : hi = bf16(src), lo = bf16(src - hi), lo2 = bf16(src - hi - lo) (lo / lo2 may be NULL: plain
 *                      cast / 2-way split).  Operand preparation for vsx_gemm (weights every step; activations only
 *                      in the fp32 parity mode).
 * vsx_scale_mask_cast: out[m,n] = n < n_keep ? g[m,n] * row_scale[m / rows_per_sample] : 0 -- the gradient of
 *                      `x + mask * drop_path(f)` w.r.t. f (nets/supernet_blocks.py:243-253, nets/drop.py:25);
 *                      also the forward of a stand-alone ChannelDrop (nets/channel_drop.py:80-81).
 * vsx_colsum         : out[c] += sum_r x[r,c]  (bias gradients of every nn.Linear).
 * -------------------------------------------------------------------------------------------------- */
int vsx_split_bf16(const float* src, long lds, void* hi, void* lo, void* lo2, long ldd, int rows, int cols, void* stream);
int vsx_scale_mask_cast(const float* g, long ldg, const float* row_scale, int rows_per_sample, int n_keep, void* out,
                        int dtype, long ldo, int rows, int cols, float* colsum /* NULL, or [>= n_keep]: += column sums of out */,
                        void* stream);
int vsx_colsum(const void* x, int dtype, long ldx, int rows, int cols, float* out, void* stream);

/* ----------------------------------------------------------------------------------------------------
 * Convolutions as GEMMs -- replaces the cuDNN convs of PatchConvEmbed (nets/patch_conv.py:23-36, :63-74) and the
 * SR block's patch_reduce (nets/vit_sr_supernet.py:139-143).  Feature maps are channels-last [B,H,W,C]
 * (`pix_pitch` elements between pixels, `batch_pitch` between samples, so token tensors [B,1+g*g,C] are read in
 * place).  vsx_im2col gathers k x k taps (tap-major, channel-minor columns: weights are used as [O, kh, kw, I]) and
 * fuses the producer's BatchNorm-apply + ReLU (scale/shift per channel, NULL = identity) and an optional second,
 * added input (the stem's residual, patch_conv.py:69-71).  nchw != 0: in1 is the fp32 image [B,C,H,W].
 * vsx_col2im is the transposed gather for the data gradient (optionally adding `add`).
 * -------------------------------------------------------------------------------------------------- */
int vsx_im2col(const void* in1, const float* scale1, const float* shift1, const void* in2, const float* scale2,
               const float* shift2, int in_dtype, int nchw, long batch_pitch, long pix_pitch, int B, int H, int W, int C,
               int k, int stride, int pad, void* out, int out_dtype, long ldo, void* stream);
int vsx_col2im(const void* dcol, long ldc, const void* add, int dtype, int B, int H, int W, int C, int k, int stride,
               int pad, void* din, long batch_pitch, long pix_pitch, void* stream);

/* BatchNorm2d in training mode over a channels-last map y[P,C] (C <= 32) -- nn.BatchNorm2d at nets/patch_conv.py:28.
 * stats: sums[0..C) += sum y, sums[C..2C) += sum y^2 (fp64, zero them first).  finalize: biased batch variance for
 * the normalisation (scale = gamma*rstd, shift = beta - mean*scale), running stats move with `momentum` towards
 * (mean, unbiased var), num_batches_tracked += 1 (pointers may be NULL).  Backward through relu(bn(y)):
 * bwd_stats: sums = (sum dz, sum dz*zhat) with dz = da*[bn(y) > 0]; bwd_apply: dy, and dgamma/dbeta += sums. */
int vsx_bn_stats(const void* y, int dtype, long P, int C, double* sums, void* stream);
int vsx_bn_finalize(const double* sums, long P, int C, const float* gamma, const float* beta, float eps, float momentum,
                    float* scale, float* shift, float* mean, float* rstd, float* running_mean, float* running_var,
                    long long* num_batches_tracked, void* stream);
int vsx_bn_bwd_stats(const void* da, const void* y, int dtype, long P, int C, const float* gamma, const float* beta,
                     const float* mean, const float* rstd, double* sums, void* stream);
int vsx_bn_bwd_apply(const void* da, const void* y, int dtype, long P, int C, const float* gamma, const float* beta,
                     const float* mean, const float* rstd, const double* sums, void* dy, float* dgamma, float* dbeta,
                     void* stream);

/* Direct 3x3 convolution, stride 1, pad 1, C_in = C_out = C (multiple of 8, <= 32) on channels-last bf16 maps -- the
 * 24->24 ConvBnAct layers of the stem (nets/patch_conv.py:53-54) without an im2col matrix.
 *   out[p][n] = (add ? add[p][n] : 0) + sum_{tap,k} act(in[p+tap-1][k]) * wt[tap][n][k],  wt: bf16 [9][32][32] zero padded,
 *   act = relu(x*in_scale[k] + in_shift[k]) when in_scale != NULL (the producing layer's BatchNorm + ReLU), else identity.
 * stats_mode 1: sums[0..C) += sum(out), sums[C..2C) += sum(out^2)  (this layer's BatchNorm batch statistics);
 * stats_mode 2: sums += (sum dz, sum dz*zhat), dz = out*[gamma*zhat+beta > 0], zhat = (y_prev-mean)*rstd  (out is a gradient
 * w.r.t. relu(bn(y_prev)): the reductions vsx_bn_bwd_apply needs).  The data gradient of the conv is this same call with
 * wt[tap'][ci][co] = W[co][ci][2-ky][2-kx].   vsx_conv3x3_wgrad: dw[co][tap][ci] (fp32) += sum_p dy[p][co]*act(in[p+tap-1][ci]). */
int vsx_conv3x3(const void* in, const float* in_scale, const float* in_shift, const void* wt, const void* add, void* out, int B,
                int H, int W, int C, int stats_mode, const void* y_prev, const float* gamma, const float* beta,
                const float* mean, const float* rstd, double* sums, void* stream);
int vsx_conv3x3_wgrad(const void* dy, const void* in, const float* in_scale, const float* in_shift, float* dw, int B, int H,
                      int W, int C, void* stream);
/* First convolution of the stem (nets/patch_conv.py:25-30: 3 -> 24 channels, 3x3, stride 2, pad 1) as one kernel per direction, no
 * im2col matrix in HBM.  image: fp32 NCHW [B, 3, H, W] (values are rounded to bf16 like every GEMM operand); weight: bf16 [24, ldw >= 32],
 * column (ky*3+kx)*3 + c, zero padded; y: bf16 channels-last [B, H/2, W/2, 24]; sums (may be NULL): fp64[48] += (sum y, sum y^2).
 * vsx_conv1_wgrad: dw[co][ (ky*3+kx)*3 + c ] (fp32, pitch ldw >= 27) += sum_p dy[p][co] * image[p, tap, c]. */
int vsx_conv1_fwd(const float* image, const void* weight, long ldw, void* y, int B, int H, int W, double* sums, void* stream);
int vsx_conv1_wgrad(const float* image, const void* dy, float* dw, long ldw, int B, int H, int W, void* stream);
/* Development aid: 0 = pick the kernel automatically, 1 = legacy direct kernel (csrc/conv3x3.cu), 2 = TMA warp-specialised kernel
 * (csrc/conv3x3_tma.cu; C == 24, H and W multiples of 16). */
int vsx_conv3x3_force_impl(int impl);

/* Token assembly: x0[b,t,:] = mask * ((t < num_tokens ? tokens[t] : patches[b,t-num_tokens]) + pos_embed[t])  -- replaces cat / expand /
 * add / embed ChannelDrop at nets/vit_sr_supernet.py:399-407; backward gives dpatches (activation dtype), and
 * ACCUMULATES dpos_embed [N,C] and dtokens [num_tokens,C] (1 class token; 2 with the distillation token,
 * nets/vision_transformer_supernet.py:82-84). */
int vsx_embed_assemble(const float* patches, const float* tokens, const float* pos, float* x0, int batch,
                       int tokens_per_sample, int C, int keep, int num_tokens, void* stream);
int vsx_embed_assemble_bwd(const float* g, void* dpatches, int dtype, float* dpos, float* dtokens, int batch,
                           int tokens_per_sample, int C, int keep, int num_tokens, void* stream);

/* SR block combine (nets/vit_sr_supernet.py:131-166): y = mask2 * (cat(tok, conv + pos) + zero-pad(cat(x[:,0], avgpool2x2(x[:,1:])))).
 * Backward: dconv / dtok (activation dtype), dpos ACCUMULATED, and the residual-path gradient gres [B,1+g*g,C1]. */
int vsx_sr_combine(const float* conv, const float* tok, const float* pos, const float* x, float* y, int batch, int grid_in,
                   int C1, int C2, int keep2, void* stream);
int vsx_sr_combine_bwd(const float* gy, void* dconv, void* dtok, int dtype, float* dpos, float* gres, int batch, int grid_in,
                       int C1, int C2, int keep2, void* stream);

/* ----------------------------------------------------------------------------------------------------
 * SwitchTokenMix (token_mixup.py:101-162; call site engine.py:109-110): batch augmentation in front of the train step.
 * Samples [batch, channels, height, width] fp32 -> `out` (a different buffer); hard labels (int64) -> targets [batch, num_classes] and
 * patch_targets [batch, patch_len^2, num_classes].  First batch/2 samples: the patch box [box_y0, box_y1) x [box_x0, box_x1) is pasted
 * from sample perm_patch[b] (and the per-patch targets with it), target = y*lam_patch + y[perm]*(1 - lam_patch); the remaining samples:
 * image-level mixup with partner batch/2 + perm_image[b - batch/2] and lam_image.  on_value / off_value: the smoothed one-hot levels.
 * All draws are the caller's (host RNG protocol of the reference); the arithmetic reproduces the reference's fp32 results bit for bit.
 * -------------------------------------------------------------------------------------------------- */
int vsx_token_mix(const float* samples, float* out, const long* labels, const int* perm_patch, const int* perm_image, float* targets,
                  float* patch_targets, int batch, int channels, int height, int width, int patch_len, int num_classes, int box_y0,
                  int box_y1, int box_x0, int box_x1, float on_value, float off_value, double lam_patch, double lam_image, void* stream);

/* --------------------------------------------------------------------------------------------------
 * Evaluation tail of one batch (replaces torch.nn.CrossEntropyLoss + timm accuracy(topk=(1,5)) + three .item() synchronisations per
 * batch, engine.py:195,222-233): logits [rows, cols] fp32 (row stride ld), hard labels int64.  row_loss / row_rank: caller scratch of
 * `rows` elements.  totals (fp64[5], device) accumulate: sum of per-batch MEAN losses, top-1 hits, top-5 hits, samples, batches -- the
 * numerators / denominators of the reference's meters (loss is averaged per batch, accuracies per sample).  A sample is a top-k hit
 * when fewer than k logits are strictly larger than its label's logit.  Deterministic (fixed-order reduction).
 * -------------------------------------------------------------------------------------------------- */
int vsx_eval_metrics(const float* logits, long ld, const long* labels, int rows, int cols, float* row_loss, int* row_rank, double* totals,
                     void* stream);

/* Input side of the step (engine.py:104-105 uploads fp32 images: 4 bytes per pixel).  vsx_image_normalize_u8 does torchvision's
 * ToTensor + Normalize on the device: out[i, c, p] = (in[i, c, p] / 255 - mean[c]) / std[c] for a uint8 batch [images, channels, pixels]
 * (mean / std: HOST arrays of `channels` floats), so that the per-step upload is 1 byte per pixel (engine.DeviceFeeder(normalize=...)). */
int vsx_image_normalize_u8(const void* in_u8, float* out, int images, int channels, long pixels, const float* mean, const float* stdv,
                           void* stream);

/* ----------------------------------------------------------------------------------------------------
 * Several extents in ONE launch: a batch that carries several sub-architectures (the multi-architectural sampling of
 * nets/channel_drop.py:93-111: sample i uses mask row perm[i % G]) is a sequence of row ranges ("segments": consecutive samples that
 * share one set of keep counts).  The *_segs entry points process the whole batch in one launch and look each row's extents up in the
 * table, instead of one launch per segment.  keep[i] == 0: the rows of segment i are skipped (dropped layer).  When a shape does not
 * qualify for the single-launch kernels they fall back to one launch per segment with identical results.
 * -------------------------------------------------------------------------------------------------- */
#define VSX_MAX_SEGMENTS 8
typedef struct vsx_row_segments {
  int count;                        /* 1 .. VSX_MAX_SEGMENTS */
  int row_end[VSX_MAX_SEGMENTS];    /* exclusive end row of segment i, relative to the first row of the launch; the last one = rows */
  int keep[VSX_MAX_SEGMENTS];       /* kept channels (LayerNorm: kept embedding width; cast: kept output width) */
  int keep2[VSX_MAX_SEGMENTS];      /* vsx_masked_ln_bwd_segs: kept width of the fused cast output (ignored without cast_out) */
} vsx_row_segments;
int vsx_masked_ln_fwd_segs(const float* x, long ldx, const float* gamma, const float* beta, void* y, int dtype, long ldy, float* mean,
                           float* rstd, int rows, int C, const vsx_row_segments* segs, float eps, void* stream);
int vsx_masked_ln_bwd_segs(const void* dy, int dtype, long lddy, const float* x, long ldx, const float* mean, const float* rstd,
                           const float* gamma, const float* g_in, float* g_out, long ldg, float* dgamma, float* dbeta, int rows, int C,
                           const vsx_row_segments* segs, void* cast_out, long ld_cast, const float* cast_scale, int cast_rows_per_sample,
                           float* cast_colsum, void* stream);
int vsx_scale_mask_cast_segs(const float* g, long ldg, const float* row_scale, int rows_per_sample, void* out, int dtype, long ldo,
                             int rows, int cols, const vsx_row_segments* segs, float* colsum, void* stream);

/* Attention core over a batch whose consecutive sample ranges keep different numbers of heads (heads_keep[i] == 0: samples skipped). */
typedef struct vsx_sample_segments {
  int count;                           /* 1 .. VSX_MAX_SEGMENTS */
  int sample_end[VSX_MAX_SEGMENTS];    /* exclusive end sample of segment i; the last one = batch */
  int heads_keep[VSX_MAX_SEGMENTS];
} vsx_sample_segments;
int vsx_attn_fwd_segs(const void* qkv, void* o, float* lse, int dtype, int batch, int tokens, int heads, int head_dim,
                      const vsx_sample_segments* segs, float scale, int impl, void* stream);
int vsx_attn_bwd_segs(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int dtype, int batch, int tokens,
                      int heads, int head_dim, const vsx_sample_segments* segs, float scale, int impl, float* dbias, void* stream);

/* ----------------------------------------------------------------------------------------------------
 * Loss and optimizer ends of the step (engine.py:152-157, :175-177).
 * vsx_soft_ce: *loss_sum += loss_scale * sum_rows(-sum_c t*log_softmax(x)); dlogits = grad_scale*(softmax*sum(t) - t)
 *              (timm SoftTargetCrossEntropy, main.py:392-394; dlogits may be NULL).
 * vsx_adamw  : one launch over all parameters (torch.optim.AdamW semantics: decoupled decay, bias correction); each
 *              chunk i of vsx_adamw_chunk_elems() elements belongs to tensor chunk_tensor[i] at chunk_index[i].
 *              shadow_hi / shadow_lo (bf16, may be NULL) receive the refreshed GEMM operand copies of the weight;
 *              ema (fp32, may be NULL) the updated moving average.  guard_loss_dev (may be NULL): the step's loss on the device;
 *              when it is not finite the launch updates NOTHING and increments *nonfinite_count_dev -- the device-side form of the
 *              reference's `if not math.isfinite(loss_value): sys.exit(1)` (engine.py:168-173), read by the host once per logging
 *              interval instead of once per step.
 * -------------------------------------------------------------------------------------------------- */
typedef struct vsx_adamw_tensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  void* shadow_hi;
  void* shadow_lo;
  long numel;
  float weight_decay;
  float ema_decay;   /* used when ema != NULL */
  float* ema;        /* NULL, or the exponential moving average of the parameter (timm ModelEmaV2, main.py:357-363, engine.py:179-180):
                        ema = ema_decay * ema + (1 - ema_decay) * param_new, written in the same pass */
} vsx_adamw_tensor;
int vsx_soft_ce(const float* logits, long ld, const float* target, long ldt, int rows, int cols, float loss_scale,
                float grad_scale, float* loss_sum, float* dlogits, long ldd, void* stream);
int vsx_scale_by_scalar(float* x, long n, const float* scalar_dev, void* stream);
int vsx_adamw_chunk_elems(void);
int vsx_adamw(const vsx_adamw_tensor* tensors_dev, const int* chunk_tensor_dev, const int* chunk_index_dev, int num_chunks,
              double lr, double beta1, double beta2, double eps, int step, const float* grad_scale_dev, const float* guard_loss_dev,
              int* nonfinite_count_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSX_H_ */
