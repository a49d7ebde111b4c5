"""ViT-Res (ViT with spatial reduction) super-network -- drop-in for the reference's nets/vit_sr_supernet.py.

Same classes (`FlexibleDistillVisionTransformerSR`, `SpatialReductionPatchEmbedding`, `BypassBlock`), constructor
arguments, `forward(x, patch_output_type=None)` return convention, `set_epoch`, `no_weight_decay`, attribute and
state_dict names, and the nine `@register_model` factories.  The `network_def` grammar is the reference's (:19-45).

Execution differs: the model's forward draws every ChannelDrop of the step on the host first (same CPU-RNG draw
order as the reference, so the same seeds select the same sub-architectures), permutes the batch so that samples of
one architecture group are contiguous, and then runs stem -> token assembly -> fused half-blocks -> SR blocks -> final
norm + heads on the kept prefixes only, all in libvsx.so kernels.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from .. import core, ops
from ..core import weights, _ActOperands, split_k_for, up8
from .channel_drop import ChannelDrop
from .drop import draw_scale
from .masked_layer_norm import MaskedLayerNorm
from .patch_conv import PatchConvEmbed, PatchEmbed, to_2tuple
from .registry import register_model
from .supernet_blocks import Block, run_half_blocks
from ._masks import keep_of, make_mask

_BLOCK_EMBED_INDEX, _EMBED_CHANNEL, _EMBED_CONV_MID_CHANNELS = 0, 1, 2
_BLOCK_HEAD_INDEX, _HEAD_OUT_CHANNEL, _HEAD_IN_CHANNEL = -1, 2, 1
_BLOCK_TYPE = 0
_TYPE_IS_EMBED, _TYPE_IS_TRANS, _TYPE_IS_HEAD, _TYPE_IS_SR, _TYPE_IS_CONV_EMBED, _TYPE_IS_FLEXIBLE_CONV_EMBED = 0, 1, 2, 3, 4, 5
_NUM_WARMUP_EPOCHS = 15


def _cfg(url='', **kwargs):
    return {'url': url, 'num_classes': 1000, 'input_size': (3, 224, 224), 'pool_size': None, 'crop_pct': .9,
            'interpolation': 'bicubic', 'mean': (0.485, 0.456, 0.406), 'std': (0.229, 0.224, 0.225),
            'first_conv': 'patch_embed.proj', 'classifier': 'head', **kwargs}


def trunc_normal_(t, std=.02):
    return nn.init.trunc_normal_(t, mean=0., std=std, a=-2., b=2.)


def _runs(keep, batch, full):
    if keep is None:
        return [(0, batch, full)]
    out, b = [], 0
    while b < batch:
        e = b
        while e < batch and keep[e] == keep[b]:
            e += 1
        out.append((b, e, int(keep[b])))
        b = e
    return out


class BypassBlock(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, x, embed_mask=None, layer_mask=None):
        return x, embed_mask, None


# =====================================================================================================================
# token assembly: cat(tokens, patches) + pos_embed, embed ChannelDrop          (reference :399-407)
# =====================================================================================================================
class _EmbedAssembleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, patches, tokens, pos, keep):
        core.require_cuda(patches, 'FlexibleDistillVisionTransformerSR')
        B, Np, C = patches.shape
        T = tokens.shape[1]                  # 1 class token (+ 1 distillation token in the patch16 network)
        patches = patches.contiguous()
        x0 = torch.empty(B, Np + T, C, device=patches.device)
        for b0, b1, k in _runs(keep, B, C):
            ops.call('embed_assemble', (patches, b0 * Np * C), tokens, pos, (x0, b0 * (Np + T) * C), b1 - b0, Np + T, C, k, T)
        ctx.keep, ctx.shape = keep, (B, Np, C, T)
        return x0

    @staticmethod
    def backward(ctx, g):
        B, Np, C, T = ctx.shape
        g = g.contiguous()
        dpatches = torch.empty(B, Np, C, device=g.device)
        dpos = torch.zeros(1, Np + T, C, device=g.device)
        dtok = torch.zeros(1, T, C, device=g.device)
        for b0, b1, k in _runs(ctx.keep, B, C):
            ops.call('embed_assemble_bwd', (g, b0 * (Np + T) * C), (dpatches, b0 * Np * C), ops.F32, dpos, dtok, b1 - b0, Np + T, C, k, T)
        return dpatches, dtok, dpos, None


# =====================================================================================================================
# spatial-reduction block                                                      (reference :59-172)
# =====================================================================================================================
class _SRFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta, x, ln_w, ln_b, conv_w, conv_b, tok_w, tok_b, pos):
        core.require_cuda(x, 'SpatialReductionPatchEmbedding')
        x = x.contiguous()
        g, C1, C2, eps, segs = meta          # segs: (b0, b1, keep_in, keep_out)
        B, N1, _ = x.shape
        g2 = g // 2
        N2 = 1 + g2 * g2
        T = core.act_dtype()
        dt = ops._DT[T]
        dev = x.device
        x2 = x.view(B * N1, C1)
        xn = torch.empty(B * N1, C1, device=dev, dtype=T)
        mean, rstd = torch.empty(B * N1, device=dev), torch.empty(B * N1, device=dev)
        A = torch.empty(B * g2 * g2, 9 * C1, device=dev, dtype=T)
        conv = torch.empty(B * g2 * g2, C2, device=dev)
        tok = torch.empty(B, C2, device=dev)
        y = torch.empty(B, N2, C2, device=dev)
        wc, wt = weights.get(conv_w, 'ohwi'), weights.get(tok_w)
        acts = _ActOperands()
        for b0, b1, k1, k2 in segs:
            nb = b1 - b0
            r1, rc = b0 * N1, b0 * g2 * g2
            ops.masked_ln_fwd(x2, C1, ln_w, ln_b, xn, C1, mean, rstd, nb * N1, C1, k1, eps, x_off=r1 * C1, y_off=r1 * C1, stat_off=r1)
            ops.call('im2col', (xn, r1 * C1 + C1), None, None, None, None, None, dt, 0, N1 * C1, C1, nb, g, g, C1, 3, 2, 1,
                     (A, rc * 9 * C1), dt, 9 * C1)
            ops.gemm(acts.get(A, 9 * C1, rc, nb * g2 * g2, 9 * C1), wc, 9 * C1, core.ld_of(wc), nb * g2 * g2, k2, 9 * C1, ops.EPI_STORE,
                     conv, C2, a_off=rc * 9 * C1, out_off=rc * C2, n_out=up8(k2), bias=conv_b)
            # class-token Linear on the normalised token rows (row pitch N1*C1)
            ops.gemm(acts.get(xn, C1, r1, nb * N1, k1), wt, N1 * C1, C1, nb, k2, k1, ops.EPI_STORE, tok, C2, a_off=r1 * C1,
                     out_off=b0 * C2, n_out=up8(k2), bias=tok_b)
            ops.call('sr_combine', (conv, rc * C2), (tok, b0 * C2), pos, (x, r1 * C1), (y, b0 * N2 * C2), nb, g, C1, C2, k2)
        ctx.save_for_backward(x, ln_w, conv_w, tok_w)
        ctx.meta, ctx.stuff = meta, (xn, mean, rstd, A)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, ln_w, conv_w, tok_w = ctx.saved_tensors
        xn, mean, rstd, A = ctx.stuff
        ctx.stuff = None
        g, C1, C2, eps, segs = ctx.meta
        B, N1, _ = x.shape
        g2 = g // 2
        N2 = 1 + g2 * g2
        T = core.act_dtype()
        dt = ops._DT[T]
        dev = x.device
        gy = gy.contiguous()
        x2 = x.view(B * N1, C1)
        dconv = torch.empty(B * g2 * g2, C2, device=dev, dtype=T)
        dtok = torch.empty(B, C2, device=dev, dtype=T)
        gres = torch.empty(B * N1, C1, device=dev)
        dA = torch.empty(B * g2 * g2, 9 * C1, device=dev, dtype=T)
        dxn = torch.empty(B * N1, C1, device=dev, dtype=T)
        g_in = torch.empty_like(x)
        d_lnw, d_lnb, d_cw, d_cb, d_tb, d_tw, dpos = core.zeros_like_many(
            ln_w, ln_w, torch.empty(C2, 9 * C1, device='meta'), torch.empty(C2, device='meta'), torch.empty(C2, device='meta'), tok_w,
            torch.empty(1, g2 * g2, C2, device='meta'))
        wc, wt = weights.get(conv_w, 'ohwi'), weights.get(tok_w)
        acts = _ActOperands()
        for b0, b1, k1, k2 in segs:
            nb = b1 - b0
            r1, rc, R = b0 * N1, b0 * g2 * g2, nb * g2 * g2
            ops.call('sr_combine_bwd', (gy, b0 * N2 * C2), (dconv, rc * C2), (dtok, b0 * C2), dt, dpos, (gres, r1 * C1), nb, g, C1, C2, k2)
            ops.colsum(dconv, C2, R, k2, d_cb, x_off=rc * C2)
            ops.colsum(dtok, C2, nb, k2, d_tb, x_off=b0 * C2)
            a_dc = acts.get(dconv, C2, rc, R, k2)
            a_A = acts.get(A, 9 * C1, rc, R, 9 * C1)
            ops.gemm(a_dc, a_A, C2, 9 * C1, k2, 9 * C1, R, ops.EPI_ATOMIC, d_cw, 9 * C1, a_off=rc * C2, b_off=rc * 9 * C1,
                     a_layout=ops.MNMAJOR, b_layout=ops.MNMAJOR, split_k=split_k_for(k2, 9 * C1, R))
            ops.gemm(a_dc, wc, C2, core.ld_of(wc), R, 9 * C1, k2, ops.EPI_STORE, dA, 9 * C1, a_off=rc * C2, out_off=rc * 9 * C1,
                     b_layout=ops.MNMAJOR)
            ops.call('col2im', (dA, rc * 9 * C1), 9 * C1, None, dt, nb, g, g, C1, 3, 2, 1, (dxn, r1 * C1 + C1), N1 * C1, C1)
            a_dt = acts.get(dtok, C2, b0, nb, k2)
            a_xn = acts.get(xn, C1, r1, nb * N1, k1)
            ops.gemm(a_dt, a_xn, C2, N1 * C1, k2, k1, nb, ops.EPI_ATOMIC, d_tw, C1, a_off=b0 * C2, b_off=r1 * C1, a_layout=ops.MNMAJOR,
                     b_layout=ops.MNMAJOR, split_k=1)
            ops.gemm(a_dt, wt, C2, C1, nb, k1, k2, ops.EPI_STORE, dxn, N1 * C1, a_off=b0 * C2, out_off=r1 * C1, n_out=up8(k1),
                     b_layout=ops.MNMAJOR)
            ops.masked_ln_bwd(dxn, C1, x2, C1, mean, rstd, ln_w, gres, g_in.view(B * N1, C1), C1, d_lnw, d_lnb, nb * N1, C1, k1,
                              dy_off=r1 * C1, x_off=r1 * C1, stat_off=r1, g_off=r1 * C1)
        d_cw = d_cw.view(C2, 3, 3, C1).permute(0, 3, 1, 2).contiguous()
        return None, g_in, d_lnw, d_lnb, d_cw, d_cb, d_tw, d_tb, dpos


class SpatialReductionPatchEmbedding(nn.Module):
    def __init__(self, img_size, in_features, out_features, patch_size=2, distill_token=True, num_channels_to_keep=None,
                 num_warmup_epochs=_NUM_WARMUP_EPOCHS, example_per_arch=None, single_arch=False):
        super().__init__()
        img_size, patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        assert patch_size == (2, 2), 'the SR kernels implement the reference\'s fixed 2x2 reduction'
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0])
        self.distill_token = distill_token
        self.num_tokens = 2 if distill_token else 1
        self.norm = MaskedLayerNorm(num_channels=in_features)
        self.patch_reduce = nn.Conv2d(in_features, out_features, kernel_size=patch_size[0] + 1, stride=patch_size[0],
                                      padding=patch_size[0] // 2)
        self.patch_pool = nn.AvgPool2d(kernel_size=patch_size[0], stride=patch_size[0])   # kept for repr parity; fused
        assert out_features >= in_features
        self.token_transform = nn.Linear(in_features, out_features)
        self.in_features, self.out_features = in_features, out_features
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, out_features))
        trunc_normal_(self.pos_embed, std=.02)
        self.channel_drop = None
        if num_channels_to_keep is not None:
            self.channel_drop = ChannelDrop(num_channels_to_keep=num_channels_to_keep, num_warmup_epochs=num_warmup_epochs,
                                            example_per_arch=example_per_arch, single_arch=single_arch)

    def draw(self, batch, like=None):
        return None if self.channel_drop is None else self.channel_drop.keeps(batch, self.out_features, like)

    def forward_keeps(self, x, keep_in, keep_out):
        assert self.num_tokens == 1, 'distillation-token SR path is outside the hot path (SURVEY.md §2: KD out of scope)'
        B = x.shape[0]
        C1, C2 = self.in_features, self.out_features
        segs, b = [], 0
        while b < B:
            k1 = C1 if keep_in is None else int(keep_in[b])
            k2 = C2 if keep_out is None else int(keep_out[b])
            e = b
            while e < B and (C1 if keep_in is None else int(keep_in[e])) == k1 and (C2 if keep_out is None else int(keep_out[e])) == k2:
                e += 1
            segs.append((b, e, k1, k2))
            b = e
        meta = (self.img_size[0], C1, C2, self.norm.eps, segs)
        return _SRFn.apply(meta, x, self.norm.weight, self.norm.bias, self.patch_reduce.weight, self.patch_reduce.bias,
                           self.token_transform.weight, self.token_transform.bias, self.pos_embed)

    def forward(self, x, embed_mask=None, layer_mask=None):
        keep_out = self.draw(x.shape[0], None)
        y = self.forward_keeps(x, keep_of(embed_mask), keep_out)
        new_mask = None if keep_out is None else make_mask(keep_out, self.out_features, x.device)
        return y, new_mask, None


# =====================================================================================================================
# final norm + classifier heads                                               (reference :420-452)
# =====================================================================================================================
class _HeadFn(torch.autograd.Function):
    """LN over all rows (training + patch_output) or the token rows only, then cls_head on token 0 and patch_head on
    the patch rows.  Returns (cls [B,K], patch [B,N-1,K] or an empty tensor)."""

    @staticmethod
    def forward(ctx, meta, x, ln_w, ln_b, cw, cb, pw, pb):
        core.require_cuda(x, 'FlexibleDistillVisionTransformerSR')
        keep, with_patches, eps = meta
        x = x.contiguous()
        B, N, C = x.shape
        K = cw.shape[0]
        T = core.act_dtype()
        dev = x.device
        x2 = x.view(B * N, C)
        tokf = torch.empty(B, C, device=dev, dtype=T)
        patchf = torch.empty(B * (N - 1), C, device=dev, dtype=T) if with_patches else None
        mean, rstd = torch.empty(B * N, device=dev), torch.empty(B * N, device=dev)
        cls = torch.empty(B, K, device=dev)
        patch = torch.empty(B * (N - 1), K, device=dev) if with_patches else torch.empty(0, device=dev)
        wcl = weights.get(cw)
        wpa = weights.get(pw) if with_patches else None
        acts = _ActOperands()
        for b0, b1, k in _runs(keep, B, C):
            nb = b1 - b0
            if with_patches:
                ops.call('masked_ln_fwd', (x2, b0 * N * C), C, ln_w, ln_b, (tokf, b0 * C), (patchf, b0 * (N - 1) * C), ops._DT[T], C,
                         (mean, b0 * N), (rstd, b0 * N), nb * N, C, k, eps, N, 1)
                ops.gemm(acts.get(patchf, C, b0 * (N - 1), nb * (N - 1), k), wpa, C, C, nb * (N - 1), K, k, ops.EPI_STORE, patch, K,
                         a_off=b0 * (N - 1) * C, out_off=b0 * (N - 1) * K, bias=pb)
            else:   # token rows only: row pitch N*C, statistics stored densely per sample
                ops.call('masked_ln_fwd', (x2, b0 * N * C), N * C, ln_w, ln_b, (tokf, b0 * C), None, ops._DT[T], C, (mean, b0), (rstd, b0),
                         nb, C, k, eps, 0, 0)
            ops.gemm(acts.get(tokf, C, b0, nb, k), wcl, C, C, nb, K, k, ops.EPI_STORE, cls, K, a_off=b0 * C, out_off=b0 * K, bias=cb)
        ctx.save_for_backward(x, ln_w, cw, pw)
        ctx.meta, ctx.stuff = meta, (tokf, patchf, mean, rstd)
        if with_patches:
            patch = patch.view(B, N - 1, K)
        return cls, patch

    @staticmethod
    def backward(ctx, gcls, gpatch):
        x, ln_w, cw, pw = ctx.saved_tensors
        tokf, patchf, mean, rstd = ctx.stuff
        ctx.stuff = None
        keep, with_patches, eps = ctx.meta
        B, N, C = x.shape
        K = cw.shape[0]
        T = core.act_dtype()
        dev = x.device
        x2 = x.view(B * N, C)
        gcls = gcls.contiguous()
        dc = torch.empty(B, K, device=dev, dtype=T)
        d_cw, d_cb, d_pw, d_pb, d_lnw, d_lnb = core.zeros_like_many(cw, torch.empty(K, device='meta'), pw, torch.empty(K, device='meta'),
                                                                    ln_w, ln_w)
        ops.scale_mask_cast(gcls, K, None, 1, K, dc, K, B, K, colsum=d_cb)           # cast + head bias gradient in one pass
        dtokf = torch.empty(B, C, device=dev, dtype=T)
        g_in = torch.zeros_like(x) if not with_patches else torch.empty_like(x)
        wcl = weights.get(cw)
        acts = _ActOperands()
        if with_patches:
            R = B * (N - 1)
            dp = torch.empty(R, K, device=dev, dtype=T)
            ops.scale_mask_cast(gpatch.contiguous().view(R, K), K, None, 1, K, dp, K, R, K, colsum=d_pb)
            dpatchf = torch.empty(R, C, device=dev, dtype=T)
            wpa = weights.get(pw)
        for b0, b1, k in _runs(keep, B, C):
            nb = b1 - b0
            a_dc = acts.get(dc, K, b0, nb, K)
            ops.gemm(a_dc, acts.get(tokf, C, b0, nb, k), K, C, K, k, nb, ops.EPI_ATOMIC, d_cw, C, a_off=b0 * K, b_off=b0 * C,
                     a_layout=ops.MNMAJOR, b_layout=ops.MNMAJOR, split_k=1)
            ops.gemm(a_dc, wcl, K, C, nb, k, K, ops.EPI_STORE, dtokf, C, a_off=b0 * K, out_off=b0 * C, n_out=up8(k), b_layout=ops.MNMAJOR)
            if with_patches:
                r0, rows = b0 * (N - 1), nb * (N - 1)
                a_dp = acts.get(dp, K, r0, rows, K)
                ops.gemm(a_dp, acts.get(patchf, C, r0, rows, k), K, C, K, k, rows, ops.EPI_ATOMIC, d_pw, C, a_off=r0 * K, b_off=r0 * C,
                         a_layout=ops.MNMAJOR, b_layout=ops.MNMAJOR, split_k=split_k_for(K, k, rows))
                ops.gemm(a_dp, wpa, K, C, rows, k, K, ops.EPI_STORE, dpatchf, C, a_off=r0 * K, out_off=r0 * C, n_out=up8(k),
                         b_layout=ops.MNMAJOR)
                ops.call('masked_ln_bwd', (dtokf, b0 * C), (dpatchf, r0 * C), ops._DT[T], C, (x2, b0 * N * C), C, (mean, b0 * N),
                         (rstd, b0 * N), ln_w, None, (g_in, b0 * N * C), C, d_lnw, d_lnb, nb * N, C, k, N, 1)
            else:
                ops.call('masked_ln_bwd', (dtokf, b0 * C), None, ops._DT[T], C, (x2, b0 * N * C), N * C, (mean, b0), (rstd, b0), ln_w,
                         None, (g_in, b0 * N * C), N * C, d_lnw, d_lnb, nb, C, k, 0, 0)
        return None, g_in, d_lnw, d_lnb, d_cw, d_cb, (d_pw if with_patches else None), (d_pb if with_patches else None)


# =====================================================================================================================
# the model
# =====================================================================================================================
class FlexibleDistillVisionTransformerSR(nn.Module):
    def __init__(self, img_size=224, patch_size=14, in_chans=3, num_classes=1000, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., norm_layer=MaskedLayerNorm, distill_token=True, network_def=None, supernet=False,
                 num_channels_to_keep=None, example_per_arch=None, num_warmup_epochs=_NUM_WARMUP_EPOCHS, single_arch=False,
                 hybrid_arch=False, patch_output=False):
        super().__init__()
        assert patch_size == 14
        assert drop_rate == 0. and attn_drop_rate == 0., 'dropout is always 0 on the reference path (scripts/vit-sr-nas)'
        self.network_def = network_def
        self.num_classes = num_classes
        assert network_def[_BLOCK_HEAD_INDEX][_HEAD_OUT_CHANNEL] == num_classes
        embed_dim = network_def[_BLOCK_EMBED_INDEX][_EMBED_CHANNEL]
        self.num_features = self.embed_dim = embed_dim
        etype = network_def[_BLOCK_EMBED_INDEX][_BLOCK_TYPE]
        if etype == _TYPE_IS_FLEXIBLE_CONV_EMBED:
            self.patch_embed = PatchConvEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                              mid_chans=network_def[_BLOCK_EMBED_INDEX][_EMBED_CONV_MID_CHANNELS])
        elif etype == _TYPE_IS_CONV_EMBED:
            self.patch_embed = PatchConvEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        else:
            self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        img_size = img_size // patch_size
        self.distill_token = distill_token
        self.num_tokens = 2 if distill_token else 1
        self.tokens = nn.Parameter(torch.zeros(1, self.num_tokens, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + self.num_tokens, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.patch_output = patch_output

        self.embed_channel_drop = None
        if supernet:
            assert num_channels_to_keep is not None, 'Super-network numbers of channels to keep error'
            assert (example_per_arch is not None) or single_arch, 'Super-network forward-backward architecture error'
            assert isinstance(num_channels_to_keep, list), 'Num of channels to keep type error'
            assert len(num_channels_to_keep) == len(network_def), 'Lengths of num_channels_to_keep and network_def are not the same'
            self.embed_channel_drop = ChannelDrop(num_channels_to_keep=num_channels_to_keep[0], num_warmup_epochs=num_warmup_epochs,
                                                  example_per_arch=example_per_arch, single_arch=(single_arch or hybrid_arch))
        depth = sum(1 for d in network_def if d[_BLOCK_TYPE] == _TYPE_IS_TRANS)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        blocks, depth = [], 0
        for i, d in enumerate(network_def):
            if d[_BLOCK_TYPE] not in (_TYPE_IS_SR, _TYPE_IS_TRANS):
                continue
            keep_attn = keep_mlp = keep_layer = keep_block = None
            if supernet:
                keep_block = num_channels_to_keep[i]
                if d[_BLOCK_TYPE] == _TYPE_IS_TRANS:
                    assert isinstance(keep_block, dict)
                    keep_attn, keep_mlp, keep_layer = keep_block['attn'], keep_block['mlp'], keep_block['layer']
                else:
                    assert isinstance(keep_block, np.ndarray)
            if d[_BLOCK_TYPE] == _TYPE_IS_TRANS:
                assert d[1][0] == d[2][0], 'Block {}: embedding dim mismatch'.format(depth)
                assert d[1][0] == embed_dim, 'Block {}: embedding dim is not consistent with patch embedding'.format(depth)
                cls = Block if d[3] else BypassBlock
                blocks.append(cls(dim=embed_dim, num_heads=d[1][1], head_dim=d[1][2], mlp_features=d[2][1], drop_path=dpr[depth],
                                  num_chs_to_keep_attn=keep_attn, num_chs_to_keep_mlp=keep_mlp, num_chs_to_keep_block=keep_layer,
                                  num_warmup_epochs=num_warmup_epochs, example_per_arch=example_per_arch, single_arch=single_arch))
                depth += 1
            else:
                assert d[1] == embed_dim, 'Block {}: SR input embedding size error'.format(i)
                blocks.append(SpatialReductionPatchEmbedding(img_size=img_size, in_features=d[1], out_features=d[2],
                                                             num_channels_to_keep=keep_block, num_warmup_epochs=num_warmup_epochs,
                                                             example_per_arch=example_per_arch, single_arch=(single_arch or hybrid_arch),
                                                             distill_token=distill_token))
                embed_dim = d[2]
                img_size = blocks[-1].img_size[0] // blocks[-1].patch_size[0]
        self.blocks = nn.ModuleList(blocks)
        self.norm = norm_layer(embed_dim)
        assert embed_dim == network_def[_BLOCK_HEAD_INDEX][_HEAD_IN_CHANNEL]
        self.cls_head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.dst_head = (nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()) if distill_token else None
        self.patch_head = None
        if patch_output:
            assert not distill_token, 'Currently support only either ShiftTokenMixup or Distillation.'
            self.patch_head = nn.Linear(embed_dim, num_classes)
        trunc_normal_(self.pos_embed, std=.02)
        trunc_normal_(self.tokens, std=.02)
        self.apply(self._init_weights)
        self.num_warmup_epochs = num_warmup_epochs
        self.epoch_now = None
        self.is_supernet = supernet
        self.example_per_arch = example_per_arch
        self.last_keeps = None       # per-entry keep dicts of the most recent forward (original sample order)
        self.active_subnet = None    # candidate evaluation: a dense sub-network definition run on THESE weights (set_active_subnet)
        self.weights_resident = False

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, (nn.LayerNorm, MaskedLayerNorm)):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        no_wd = ['tokens']
        for name, _ in self.blocks.named_parameters():
            if name.endswith(tuple(no_wd)):
                no_wd.append(name)
        return set(no_wd)

    def get_classifier(self):
        return self.cls_head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.cls_head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.dst_head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    # ------------------------------------------------------------------ sub-architecture sampling
    def sample_keeps(self, batch):
        """All ChannelDrop draws of one forward, in the reference's module execution order:
        embed drop; per Block attn, layer, mlp; each SR block's drop at its end.  Returns a list aligned with
        network_def of {'embed'|'attn'|'layer'|'mlp': [keep per sample]} (empty dicts where nothing is drawn)."""
        out, j = [], 0
        for i, d in enumerate(self.network_def):
            if i == 0:
                cd = self.embed_channel_drop
                out.append({} if cd is None else {'embed': cd.keeps(batch, self.embed_dim)})
            elif d[_BLOCK_TYPE] == _TYPE_IS_TRANS:
                blk = self.blocks[j]
                j += 1
                out.append({k: v for k, v in blk.draw(batch).items() if v is not None} if isinstance(blk, Block) else {})
            elif d[_BLOCK_TYPE] == _TYPE_IS_SR:
                k = self.blocks[j].draw(batch)
                j += 1
                out.append({} if k is None else {'embed': k})
            else:
                out.append({})
        return out

    @staticmethod
    def _group_layout(keeps, batch):
        """-> (order, bounds): the order that makes samples with identical keep tuples contiguous (stable; None = already so) and the
        sample indices (in the new order) at which the sub-architecture changes (None = one architecture).  Signatures are built with
        zip (C speed): this runs once per step on ~50 keep lists of `batch` integers."""
        cols = [v for k in keeps for v in k.values()]
        if not cols:
            return None, None
        sig = list(zip(*cols))
        if len(set(sig)) <= 1:
            return None, None
        first = {}
        for b, s in enumerate(sig):
            first.setdefault(s, b)
        order = sorted(range(batch), key=lambda b: (first[sig[b]], b))
        bounds = {i for i in range(1, batch) if sig[order[i]] != sig[order[i - 1]]}
        return (None if order == list(range(batch)) else order), bounds

    @staticmethod
    def _group_permutation(keeps, batch):
        """Order that makes samples with identical keep tuples contiguous (stable).  Identity when there is one group."""
        return FlexibleDistillVisionTransformerSR._group_layout(keeps, batch)[0]

    # ------------------------------------------------------------------ candidate evaluation on resident weights
    def subnet_extents(self, sub_network_def):
        """Prefix extents that make this network compute exactly what the reference's evolutionary search evaluates for
        `sub_network_def` (evo_search.py:256-273: a freshly built dense sub-network loaded with nets/net_utils.get_sub_state_dict
        prefix slices of these weights): embed / SR widths, heads * head_dim, MLP features; sub-network blocks with exists=0 are
        skipped (BypassBlock).  -> list aligned with network_def of {'embed' | 'attn' + 'mlp' | 'skip'}."""
        sup = self.network_def
        if len(sub_network_def) != len(sup):
            raise ValueError('sub-network definition has %d entries, this network %d' % (len(sub_network_def), len(sup)))
        out, width_sup, width = [], None, None
        for i, (d, u) in enumerate(zip(sub_network_def, sup)):
            if i == 0:
                if d[_BLOCK_TYPE] != u[_BLOCK_TYPE] or d[1] > u[1] or tuple(d[2:]) != tuple(u[2:]):
                    raise ValueError('entry 0: embedding %s does not fit %s' % (d, u))
                width, width_sup = d[1], u[1]
                out.append({'embed': width})
            elif u[_BLOCK_TYPE] == _TYPE_IS_TRANS:
                if d[_BLOCK_TYPE] != _TYPE_IS_TRANS or d[1][0] != width or d[2][0] != width:
                    raise ValueError('entry %d: %s is not a transformer block of width %d' % (i, d, width))
                if not d[3]:
                    out.append({'skip': True})
                    continue
                if not u[3] or d[1][2] != u[1][2] or d[1][1] > u[1][1] or d[2][1] > u[2][1]:
                    raise ValueError('entry %d: block %s does not fit inside %s' % (i, d, u))
                out.append({'attn': d[1][1] * d[1][2], 'mlp': d[2][1]})
            elif u[_BLOCK_TYPE] == _TYPE_IS_SR:
                if d[_BLOCK_TYPE] != _TYPE_IS_SR or d[1] != width or d[2] > u[2]:
                    raise ValueError('entry %d: SR block %s does not fit %s at width %d' % (i, d, u, width))
                width, width_sup = d[2], u[2]
                out.append({'embed': width})
            else:
                if d[_BLOCK_TYPE] != u[_BLOCK_TYPE] or d[1] != width or d[2] != u[2]:
                    raise ValueError('entry %d: head %s does not fit %s at width %d' % (i, d, u, width))
                out.append({})
        return out

    def set_active_subnet(self, sub_network_def, weights_resident=True):
        """Evaluate `sub_network_def` (None: back to the network's own behaviour) with eval-mode forwards of THIS module: no model
        build, no state-dict slicing or copy, no device upload per candidate.  weights_resident: the
        parameters do not change between forwards, so the bf16 GEMM operand copies are derived once, not once per forward."""
        self.active_subnet = None if sub_network_def is None else self.subnet_extents(sub_network_def)
        self.weights_resident = bool(weights_resident) and sub_network_def is not None
        if self.weights_resident:
            core.weights.generation += 1

    # ------------------------------------------------------------------ forward
    def forward_features(self, x):
        core.require_cuda(x, 'FlexibleDistillVisionTransformerSR')
        assert self.num_tokens == 1, 'distillation-token path is outside the hot path (SURVEY.md §2)'
        core.reset_half_chain()
        B = x.shape[0]
        if self.active_subnet is not None:
            if self.training:
                raise RuntimeError('set_active_subnet() is the evaluation path of the evolutionary search: call model.eval() first')
            keeps = [{k: ([v] * B if k != 'skip' else v) for k, v in e.items()} for e in self.active_subnet]
        else:
            keeps = self.sample_keeps(B) if self.is_supernet else [{} for _ in self.network_def]
        if self.training or not self.weights_resident:
            core.weights.generation += 1      # weights may have been updated since the last forward: re-derive operand copies
        self.last_keeps = keeps
        perm, bounds = (None, None) if self.active_subnet is not None else self._group_layout(keeps, B)
        if perm is not None:       # make architecture groups contiguous; undone on the logits
            idx = core.h2d(perm, x.device)
            x = x.index_select(0, idx)
            keeps = [{k: [v[p] for p in perm] for k, v in kd.items()} for kd in keeps]
        depth = sum(1 for b in self.blocks if isinstance(b, (Block, BypassBlock)))
        rates = [getattr(getattr(b, 'drop_path', None), 'drop_prob', 0.) or 0. for b in self.blocks if isinstance(b, (Block, BypassBlock))]
        dp = None
        if self.training and any(r > 0 for r in rates):
            u = torch.rand((depth, 2, B), device=x.device)               # CUDA generator, like nets/drop.py:20
            keep_prob = 1.0 - core.h2d(rates, x.device).view(depth, 1, 1)
            dp = ((keep_prob + u).floor_() / keep_prob).view(depth * 2, B).contiguous()

        h = self.patch_embed(x)
        embed_keep = keeps[0].get('embed')
        h = _EmbedAssembleFn.apply(h, self.tokens, self.pos_embed, embed_keep)
        if core.trunk_grads_ready_hook is not None and h.requires_grad:
            # fires in backward once every transformer / SR block has produced its parameter gradients and only the stem is left:
            # engine.TrainStep starts the gradient all-reduce of that part of the gradient pool there (overlaps the stem backward)
            h.register_hook(core.trunk_grads_ready_hook)
        layer_keep = None
        j = t = 0
        run_metas, run_params = [], []          # consecutive transformer blocks of a stage run as ONE autograd node (core.StageFn)

        def flush(h):
            if run_metas:
                h = run_half_blocks(h, list(run_metas), tuple(run_params))
                run_metas.clear()
                run_params.clear()
            return h
        for i, d in enumerate(self.network_def):
            if d[_BLOCK_TYPE] == _TYPE_IS_TRANS:
                blk = self.blocks[j]
                if isinstance(blk, Block) and not keeps[i].get('skip'):
                    use_dp = dp if rates[t] > 0 else None
                    meta_a, meta_m, layer_keep = blk.half_metas(B, h.shape[1], h.shape[2], embed_keep, layer_keep, keeps[i], use_dp, 2 * t, bounds, True)
                    run_metas.extend((meta_a, meta_m))
                    run_params.extend(blk.half_params())
                else:
                    layer_keep = None
                j += 1
                t += 1
            elif d[_BLOCK_TYPE] == _TYPE_IS_SR:
                h = flush(h)
                new_keep = keeps[i].get('embed')
                h = self.blocks[j].forward_keeps(h, embed_keep, new_keep)
                embed_keep, layer_keep = new_keep, None
                j += 1
        h = flush(h)
        return h, embed_keep, perm

    def forward(self, x, patch_output_type=None):
        h, embed_keep, perm = self.forward_features(x)
        with_patches = bool(self.training and self.patch_output)
        if with_patches and patch_output_type not in ('seq', None):
            raise ValueError("only the 'seq' patch output of the reference's training path is implemented")
        pw = self.patch_head.weight if self.patch_head is not None else self.cls_head.weight
        pb = self.patch_head.bias if self.patch_head is not None else self.cls_head.bias
        cls_pred, patch_pred = _HeadFn.apply((embed_keep, with_patches, self.norm.eps), h, self.norm.weight, self.norm.bias,
                                             self.cls_head.weight, self.cls_head.bias, pw, pb)
        if perm is not None:
            inv = torch.empty(len(perm), dtype=torch.long)
            inv[torch.tensor(perm)] = torch.arange(len(perm))
            inv = core.h2d(inv, cls_pred.device)
            cls_pred = cls_pred.index_select(0, inv)
            if with_patches:
                patch_pred = patch_pred.index_select(0, inv)
        if self.patch_output:
            return (cls_pred, patch_pred) if self.training else cls_pred
        return cls_pred

    def set_epoch(self, epoch):
        self.epoch_now = epoch
        for m in self.modules():
            if isinstance(m, ChannelDrop):
                m.set_epoch(epoch)
        if self.is_supernet and self.num_warmup_epochs >= self.epoch_now:
            for m in self.modules():
                if isinstance(m, Block):
                    m.rewiring()


def _factory(distill, supernet, patch_output, img_size=224):
    def make(pretrained=False, **kwargs):
        kw = dict(patch_size=14, distill_token=distill, patch_output=patch_output)
        if supernet:
            kw['supernet'] = True
        if img_size != 224:
            kw['img_size'] = img_size
        model = FlexibleDistillVisionTransformerSR(**kw, **kwargs)
        model.default_cfg = _cfg()
        return model
    return make


flexible_vit_sr_distill_patch14_224 = register_model('flexible_vit_sr_distill_patch14_224', _factory(True, False, False))
flexible_vit_sr_patch14_224 = register_model('flexible_vit_sr_patch14_224', _factory(False, False, False))
flexible_vit_sr_distill_patch14_224_supernet = register_model('flexible_vit_sr_distill_patch14_224_supernet', _factory(True, True, False))
flexible_vit_sr_patch14_224_supernet = register_model('flexible_vit_sr_patch14_224_supernet', _factory(False, True, False))
flexible_vit_sr_patch14_224_patch_output = register_model('flexible_vit_sr_patch14_224_patch_output', _factory(False, False, True))
flexible_vit_sr_patch14_224_patch_output_supernet = register_model('flexible_vit_sr_patch14_224_patch_output_supernet',
                                                                   _factory(False, True, True))
flexible_vit_sr_patch14_280_patch_output = register_model('flexible_vit_sr_patch14_280_patch_output', _factory(False, False, True, 280))
flexible_vit_sr_patch14_336_patch_output = register_model('flexible_vit_sr_patch14_336_patch_output', _factory(False, False, True, 336))
flexible_vit_sr_patch14_392_patch_output = register_model('flexible_vit_sr_patch14_392_patch_output', _factory(False, False, True, 392))
