"""Patch-16 ViT super-network with a distillation token -- drop-in for the reference's nets/vision_transformer_supernet.py
(SURVEY.md §8(f) row 4): same class (`FlexibleDistillVisionTransformer`), constructor arguments, `forward` return convention
(`cls_pred` or `(cls_pred, dst_pred)`), `set_epoch`, `no_weight_decay`, state_dict names and the four `@register_model` factories.

It is the ViT-Res machinery without spatial reduction: timm `PatchEmbed` (16x16 stride-16 conv as im2col + GEMM), class + distillation
token, the same `Block` half-block kernels over N = 196 + num_tokens rows, final masked LayerNorm evaluated on the token rows only (the
reference normalises all rows and keeps the first `num_tokens`, :205-206 -- LayerNorm is per row, so the result is identical), one
Linear head per token.
"""
import torch
import torch.nn as nn

from .. import core, ops
from ..core import weights, _ActOperands, up8
from .channel_drop import ChannelDrop
from .masked_layer_norm import MaskedLayerNorm
from .patch_conv import PatchEmbed
from .registry import register_model
from .supernet_blocks import Block, run_half_blocks
from .vit_sr_supernet import BypassBlock, FlexibleDistillVisionTransformerSR, _EmbedAssembleFn, _cfg, _runs, trunc_normal_

_BLOCK_EMBED_INDEX, _EMBED_CHANNEL = 0, 1
_BLOCK_HEAD_INDEX, _HEAD_CHANNEL = -1, 2
_BLOCK_TYPE, _TYPE_IS_EMBED, _TYPE_IS_TRANS, _TYPE_IS_HEAD = 0, 0, 1, 2
_NUM_WARMUP_EPOCHS = 15


class _TokenHeadsFn(torch.autograd.Function):
    """Final norm on the token rows + one Linear head per token (reference :205-221).  x [B, N, C] fp32; returns one [B, K] tensor per
    token.  meta = (embed keep per sample or None, eps, number of tokens)."""

    @staticmethod
    def forward(ctx, meta, x, ln_w, ln_b, *heads):
        core.require_cuda(x, 'FlexibleDistillVisionTransformer')
        keep, eps, T = meta
        x = x.contiguous()
        B, N, C = x.shape
        K = heads[0].shape[0]
        AT = core.act_dtype()
        dev = x.device
        x2 = x.view(B * N, C)
        tokf = torch.empty(T, B, C, device=dev, dtype=AT)
        mean, rstd = torch.empty(T, B, device=dev), torch.empty(T, B, device=dev)
        outs = [torch.empty(B, K, device=dev) for _ in range(T)]
        acts = _ActOperands()
        for j in range(T):
            wj = weights.get(heads[2 * j])
            for b0, b1, k in _runs(keep, B, C):
                nb = b1 - b0
                # token row j of every sample: row pitch N*C, statistics stored densely per (token, sample)
                ops.call('masked_ln_fwd', (x2, (b0 * N + j) * C), N * C, ln_w, ln_b, (tokf, (j * B + b0) * C), None, ops._DT[AT], C,
                         (mean, j * B + b0), (rstd, j * B + b0), nb, C, k, eps, 0, 0)
                ops.gemm(acts.get(tokf, C, j * B + b0, nb, k), wj, C, C, nb, K, k, ops.EPI_STORE, outs[j], K, a_off=(j * B + b0) * C,
                         out_off=b0 * K, bias=heads[2 * j + 1])
        ctx.save_for_backward(x, ln_w, *[heads[2 * j] for j in range(T)])
        ctx.meta, ctx.stuff = meta, (tokf, mean, rstd)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        x, ln_w, *hw = ctx.saved_tensors
        tokf, mean, rstd = ctx.stuff
        ctx.stuff = None
        keep, eps, T = ctx.meta
        B, N, C = x.shape
        K = hw[0].shape[0]
        AT = core.act_dtype()
        dev = x.device
        x2 = x.view(B * N, C)
        g_in = torch.zeros_like(x)                       # only the token rows receive a gradient
        d_lnw, d_lnb = core.zeros_like_many(ln_w, ln_w)
        acts = _ActOperands()
        grads = []
        for j in range(T):
            d_w, d_b = core.zeros_like_many(hw[j], torch.empty(K, device='meta'))
            dc = torch.empty(B, K, device=dev, dtype=AT)
            ops.scale_mask_cast(gouts[j].contiguous(), K, None, 1, K, dc, K, B, K, colsum=d_b)       # cast + head bias gradient
            dtokf = torch.empty(B, C, device=dev, dtype=AT)
            wj = weights.get(hw[j])
            for b0, b1, k in _runs(keep, B, C):
                nb = b1 - b0
                a_dc = acts.get(dc, K, b0, nb, K)
                ops.gemm(a_dc, acts.get(tokf, C, j * B + b0, nb, k), K, C, K, k, nb, ops.EPI_ATOMIC, d_w, C, a_off=b0 * K,
                         b_off=(j * B + b0) * C, a_layout=ops.MNMAJOR, b_layout=ops.MNMAJOR, split_k=1)
                ops.gemm(a_dc, wj, K, C, nb, k, K, ops.EPI_STORE, dtokf, C, a_off=b0 * K, out_off=b0 * C, n_out=up8(k), b_layout=ops.MNMAJOR)
                ops.call('masked_ln_bwd', (dtokf, b0 * C), None, ops._DT[AT], C, (x2, (b0 * N + j) * C), N * C, (mean, j * B + b0),
                         (rstd, j * B + b0), ln_w, None, (g_in, (b0 * N + j) * C), N * C, d_lnw, d_lnb, nb, C, k, 0, 0)
            grads += [d_w, d_b]
        return (None, g_in, d_lnw, d_lnb) + tuple(grads)


class FlexibleDistillVisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=MaskedLayerNorm, distill_token=True, network_def=None, supernet=False, num_channels_to_keep=None,
                 example_per_arch=None, num_warmup_epochs=_NUM_WARMUP_EPOCHS, single_arch=False):
        super().__init__()
        assert drop_rate == 0. and attn_drop_rate == 0., 'dropout is always 0 on the reference path'
        self.network_def = network_def
        self.num_classes = num_classes
        assert network_def[_BLOCK_HEAD_INDEX][_HEAD_CHANNEL] == num_classes
        embed_dim = network_def[_BLOCK_EMBED_INDEX][_EMBED_CHANNEL]
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.num_tokens = 2 if distill_token else 1
        self.tokens = nn.Parameter(torch.zeros(1, self.num_tokens, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + self.num_tokens, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.embed_channel_drop = None
        if supernet:
            assert num_channels_to_keep is not None, 'Super-network numbers of channels to keep error'
            assert (example_per_arch is not None) or single_arch, 'Super-network forward-backward architecture error'
            assert isinstance(num_channels_to_keep, list), 'Num of channels to keep type error'
            self.embed_channel_drop = ChannelDrop(num_channels_to_keep=num_channels_to_keep[0], num_warmup_epochs=num_warmup_epochs,
                                                  example_per_arch=example_per_arch, single_arch=single_arch)
        depth = sum(1 for d in network_def if d[_BLOCK_TYPE] == _TYPE_IS_TRANS)
        assert depth == len(network_def) - 2, 'Block number error'
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        blocks, depth = [], 0
        for i, d in enumerate(network_def):
            if d[_BLOCK_TYPE] != _TYPE_IS_TRANS:
                continue
            assert d[1][0] == d[2][0], 'Block {}: embedding dim mismatch'.format(depth)
            assert d[1][0] == embed_dim, 'Block {}: embedding dim is not consistent with patch embedding'.format(depth)
            k = num_channels_to_keep[i] if supernet else {'attn': None, 'mlp': None, 'layer': None}
            cls = Block if d[3] else BypassBlock
            blocks.append(cls(dim=embed_dim, num_heads=d[1][1], head_dim=d[1][2], mlp_features=d[2][1], drop_path=dpr[depth],
                              num_chs_to_keep_attn=k['attn'], num_chs_to_keep_mlp=k['mlp'], num_chs_to_keep_block=k['layer'],
                              num_warmup_epochs=num_warmup_epochs, example_per_arch=example_per_arch, single_arch=single_arch))
            depth += 1
        self.blocks = nn.ModuleList(blocks)
        self.norm = norm_layer(embed_dim)
        self.cls_head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.dst_head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        trunc_normal_(self.pos_embed, std=.02)
        trunc_normal_(self.tokens, std=.02)
        self.apply(self._init_weights)
        self.num_warmup_epochs = num_warmup_epochs
        self.epoch_now = None
        self.is_supernet = supernet
        self.last_keeps = None

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
        return {'pos_embed', 'tokens'}

    def get_classifier(self):
        return self.cls_head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.cls_head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.dst_head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def sample_keeps(self, batch):
        """All ChannelDrop draws of one forward in the reference's module execution order (:193-203): embed drop, then per Block attn,
        layer, mlp.  -> list aligned with self.blocks (index 0 = the embedding)."""
        cd = self.embed_channel_drop
        out = [{} if cd is None else {'embed': cd.keeps(batch, self.embed_dim)}]
        for blk in self.blocks:
            out.append({k: v for k, v in blk.draw(batch).items() if v is not None} if isinstance(blk, Block) else {})
        return out

    _group_permutation = staticmethod(FlexibleDistillVisionTransformerSR._group_permutation)
    _group_layout = staticmethod(FlexibleDistillVisionTransformerSR._group_layout)

    def forward(self, x):
        core.require_cuda(x, 'FlexibleDistillVisionTransformer')
        core.weights.generation += 1
        core.reset_half_chain()
        B = x.shape[0]
        keeps = self.sample_keeps(B) if self.is_supernet else [{} for _ in range(len(self.blocks) + 1)]
        self.last_keeps = keeps
        perm, bounds = self._group_layout(keeps, B)
        if perm is not None:       # make architecture groups contiguous; undone on the logits
            x = x.index_select(0, core.h2d(perm, x.device))
            keeps = [{k: [v[p] for p in perm] for k, v in kd.items()} for kd in keeps]
        rates = [getattr(getattr(b, 'drop_path', None), 'drop_prob', 0.) or 0. for b in self.blocks]
        depth = len(self.blocks)
        dp = None
        if self.training and any(r > 0 for r in rates):
            u = torch.rand((depth, 2, B), device=x.device)
            keep_prob = 1.0 - core.h2d(rates, x.device).view(depth, 1, 1)
            dp = ((keep_prob + u).floor_() / keep_prob).view(depth * 2, B).contiguous()
        h = self.patch_embed(x)
        embed_keep = keeps[0].get('embed')
        h = _EmbedAssembleFn.apply(h, self.tokens, self.pos_embed, embed_keep)
        layer_keep = None
        run_metas, run_params = [], []          # all transformer blocks run as ONE autograd node (core.StageFn)
        for t, blk in enumerate(self.blocks):
            if isinstance(blk, Block):
                meta_a, meta_m, layer_keep = blk.half_metas(B, h.shape[1], h.shape[2], embed_keep, layer_keep, keeps[t + 1],
                                                            dp if rates[t] > 0 else None, 2 * t, bounds, True)
                run_metas.extend((meta_a, meta_m))
                run_params.extend(blk.half_params())
            else:
                layer_keep = None
        if run_metas:
            h = run_half_blocks(h, run_metas, tuple(run_params))
        heads = [self.cls_head.weight, self.cls_head.bias]
        if self.num_tokens == 2:
            heads += [self.dst_head.weight, self.dst_head.bias]
        outs = _TokenHeadsFn.apply((embed_keep, self.norm.eps, self.num_tokens), h, self.norm.weight, self.norm.bias, *heads)
        if perm is not None:
            inv = torch.empty(len(perm), dtype=torch.long)
            inv[torch.tensor(perm)] = torch.arange(len(perm))
            inv = core.h2d(inv, x.device)
            outs = tuple(o.index_select(0, inv) for o in outs)
        return outs[0] if self.num_tokens == 1 else (outs[0], outs[1])

    def set_epoch(self, epoch):
        self.epoch_now = epoch
        for m in self.modules():
            if isinstance(m, ChannelDrop):
                m.set_epoch(epoch)
        if self.is_supernet and self.num_warmup_epochs >= self.epoch_now:
            for m in self.modules():
                if isinstance(m, Block):
                    m.rewiring()


def _factory(supernet, img_size=224):
    def make(pretrained=False, **kwargs):
        kw = dict(patch_size=16, distill_token=True)
        if supernet:
            kw['supernet'] = True
        if img_size != 224:
            kw['img_size'] = img_size
        model = FlexibleDistillVisionTransformer(**kw, **kwargs)
        model.default_cfg = _cfg()
        return model
    return make


flexible_vit_patch16_224 = register_model('flexible_vit_patch16_224', _factory(False))
flexible_vit_patch16_224_supernet = register_model('flexible_vit_patch16_224_supernet', _factory(True))
flexible_vit_patch16_192 = register_model('flexible_vit_patch16_192', _factory(False, 192))
flexible_vit_patch16_192_supernet = register_model('flexible_vit_patch16_192_supernet', _factory(True, 192))
