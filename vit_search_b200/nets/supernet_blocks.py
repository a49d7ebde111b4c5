"""Mlp / Attention / Block of the ViT-Res super-network -- drop-in for the reference's nets/supernet_blocks.py.

Same class names, constructor arguments, sub-module / parameter names (hence state_dict keys), `forward`
signatures and `rewiring()` semantics.  The arithmetic runs in libvsx.so: each half of a Block
(LN -> qkv -> attention -> proj, LN -> fc1 -> GELU -> fc2) is one autograd node (core.HalfBlockFn) whose kernels
work on the kept prefixes only -- masked heads, hidden channels, embedding channels and dropped layers are
skipped, not computed and multiplied by zero as in the reference (:37-52, :100-120, :209-255).
"""
import torch
import torch.nn as nn

from .. import core
from .channel_drop import ChannelDrop
from .drop import DropPath, draw_scale
from .masked_layer_norm import MaskedLayerNorm
from ._masks import keep_of, and_keep, make_mask

_NUM_WARMUP_EPOCHS_CHANNEL = 15
_EXAMPLE_PER_ARCH = 16


def _cd(num_channels_to_keep, num_warmup_epochs, example_per_arch, single_arch):
    if num_channels_to_keep is None:
        return None
    return ChannelDrop(num_channels_to_keep=num_channels_to_keep, num_warmup_epochs=num_warmup_epochs,
                       example_per_arch=example_per_arch, single_arch=single_arch)


def run_half_blocks(x, metas, params):
    """x through a run of half blocks (`params`: six tensors per half block).  bf16 training path: ONE autograd node and one C-ABI call
    per direction for the whole run (core.StageFn); parity / instrumented modes: one node per half block (core.HalfBlockFn)."""
    if core.stage_native_ok(metas, params):
        return core.StageFn.apply(metas, x, *params)
    for i, meta in enumerate(metas):
        x = core.HalfBlockFn.apply(meta, x, *params[6 * i:6 * i + 6])
    return x


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.,
                 num_channels_to_keep=None, num_warmup_epochs=_NUM_WARMUP_EPOCHS_CHANNEL,
                 example_per_arch=_EXAMPLE_PER_ARCH, single_arch=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        assert act_layer is nn.GELU, 'the fused fc1 epilogue implements exact GELU (the only activation the reference uses)'
        assert drop == 0., 'dropout inside Mlp is always 0 on the reference path (nets/vit_sr_supernet.py:299-309)'
        assert out_features == in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        self.channel_drop_layer = _cd(num_channels_to_keep, num_warmup_epochs, example_per_arch, single_arch)

    def draw(self, batch, like=None):
        cd = self.channel_drop_layer
        return None if cd is None else cd.keeps(batch, self.fc1.out_features, like)

    def forward(self, x):
        """Stand-alone fc1 -> GELU -> hidden ChannelDrop -> fc2 (reference :37-52): same kernels, no LN / residual."""
        B, N, C = x.shape
        segs = core.make_segments(B, C, None, self.draw(B), self.fc1.out_features, None)
        meta = core.HalfMeta('mlp', segs, N, C, hidden=self.fc1.out_features, pre_norm=False, residual=False)
        dummy = self.fc1.bias
        return core.HalfBlockFn.apply(meta, x, dummy, dummy, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)

    def rewiring(self):
        """Sort hidden channels by L1 magnitude, largest first (reference :55-71); plain torch on the weights, once
        per warm-up epoch."""
        w = self.fc2.weight.data.abs().sum(dim=0) + self.fc1.weight.data.abs().sum(dim=1) + self.fc1.bias.data.abs()
        _, idx = torch.sort(w, descending=True)
        self.fc1.weight.data = self.fc1.weight.data[idx, :]
        self.fc1.bias.data = self.fc1.bias.data[idx]
        self.fc2.weight.data = self.fc2.weight.data[:, idx]


class Attention(nn.Module):
    def __init__(self, dim, num_heads, head_dim=64, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.,
                 num_channels_to_keep=None, num_warmup_epochs=_NUM_WARMUP_EPOCHS_CHANNEL,
                 example_per_arch=_EXAMPLE_PER_ARCH, single_arch=False):
        super().__init__()
        assert qkv_bias and qk_scale is None and attn_drop == 0. and proj_drop == 0., \
            'the reference path always uses qkv_bias=True, scale=head_dim**-0.5 and no dropout'
        self.num_heads = num_heads
        self.head_dim = head_dim
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, num_heads * head_dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(num_heads * head_dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.channel_drop_layer = _cd(num_channels_to_keep, num_warmup_epochs, example_per_arch, single_arch)

    def draw(self, batch, like=None):
        cd = self.channel_drop_layer
        if cd is None:
            return None
        keep = cd.keeps(batch, self.num_heads * self.head_dim, like)
        assert all(k % self.head_dim == 0 for k in keep), 'attention keep counts must be whole heads'
        return keep

    def forward(self, x):
        """Stand-alone qkv -> softmax attention -> head ChannelDrop -> proj (reference :100-120)."""
        B, N, C = x.shape
        hd = self.num_heads * self.head_dim
        segs = core.make_segments(B, C, None, self.draw(B), hd, None)
        meta = core.HalfMeta('attn', segs, N, C, heads=self.num_heads, head_dim=self.head_dim, pre_norm=False, residual=False)
        dummy = self.proj.bias
        return core.HalfBlockFn.apply(meta, x, dummy, dummy, self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias)

    def rewiring(self):
        """Sort heads by L1 magnitude (reference :123-161)."""
        H, D = self.num_heads, self.head_dim
        qkv_w, qkv_b, proj_w = self.qkv.weight.data, self.qkv.bias.data, self.proj.weight.data
        score = qkv_w.abs().sum(dim=1).reshape(3, H, D).sum(dim=(0, 2))
        score = score + qkv_b.abs().reshape(3, H, D).sum(dim=(0, 2))
        score = proj_w.abs().sum(dim=0).reshape(H, D).sum(dim=1) + score
        _, idx = torch.sort(score, descending=True)
        self.qkv.weight.data = qkv_w.reshape(3, H, D, -1)[:, idx].reshape(3 * H * D, -1)
        self.qkv.bias.data = qkv_b.reshape(3, H, D)[:, idx].reshape(3 * H * D)
        self.proj.weight.data = proj_w.reshape(-1, H, D)[:, idx].reshape(-1, H * D)


class Block(nn.Module):
    def __init__(self, dim, num_heads, head_dim, mlp_features, qkv_bias=True, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, num_chs_to_keep_attn=None, num_chs_to_keep_mlp=None,
                 num_chs_to_keep_block=None, num_warmup_epochs=_NUM_WARMUP_EPOCHS_CHANNEL,
                 example_per_arch=_EXAMPLE_PER_ARCH, single_arch=False):
        super().__init__()
        self.layer_drop = _cd(num_chs_to_keep_block, num_warmup_epochs, example_per_arch, single_arch)
        self.norm1 = MaskedLayerNorm(dim)
        self.attn = Attention(dim, num_heads=num_heads, head_dim=head_dim, qkv_bias=qkv_bias, qk_scale=qk_scale,
                              attn_drop=attn_drop, proj_drop=drop, num_channels_to_keep=num_chs_to_keep_attn,
                              num_warmup_epochs=num_warmup_epochs, example_per_arch=example_per_arch, single_arch=single_arch)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = MaskedLayerNorm(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_features, act_layer=act_layer, drop=drop,
                       num_channels_to_keep=num_chs_to_keep_mlp, num_warmup_epochs=num_warmup_epochs,
                       example_per_arch=example_per_arch, single_arch=single_arch)
        self.dim = dim

    # ---- keep-count interface used by the model's fused forward (no mask tensors, no device work)
    def draw(self, batch, like=None):
        """Draws in the reference's order: attn, layer, mlp."""
        k = {'attn': self.attn.draw(batch, like)}
        if self.layer_drop is not None:
            k['layer'] = self.layer_drop.keeps(batch, self.dim, like)
        k['mlp'] = self.mlp.draw(batch, like)
        return k

    def half_metas(self, B, N, C, embed_keep, layer_keep_in, keeps, dp_scale=None, dp_off=0, bounds=None, grouped=False):
        """Static descriptions of the two half blocks for one batch (host integers only).  -> (attention meta, MLP meta, current layer keep)."""
        cur = attn_ck = None
        if keeps.get('layer') is not None:                      # reference :220-223: layer_drop(f_x) masks the attention branch with
            attn_ck = list(keeps['layer'])                      # this block's OWN layer mask; the incoming one only joins `cur`
            cur = and_keep(attn_ck, layer_keep_in)
        if embed_keep is not None:                              # reference :238-243: with an embed mask the attention branch is
            cur = and_keep(cur, embed_keep)                     # multiplied by the full AND as well
            attn_ck = cur
        a, m = self.attn, self.mlp
        hd = a.num_heads * a.head_dim
        segs = core.make_segments(B, C, embed_keep, keeps.get('attn'), hd, attn_ck, bounds, grouped)
        meta_a = core.HalfMeta('attn', segs, N, C, heads=a.num_heads, head_dim=a.head_dim, row_scale=dp_scale, scale_off=dp_off * B,
                               eps=self.norm1.eps)
        segs = core.make_segments(B, C, embed_keep, keeps.get('mlp'), m.fc1.out_features, cur, bounds, grouped)
        meta_m = core.HalfMeta('mlp', segs, N, C, hidden=m.fc1.out_features, row_scale=dp_scale, scale_off=(dp_off + 1) * B,
                               eps=self.norm2.eps)
        return meta_a, meta_m, cur

    def half_params(self):
        a, m = self.attn, self.mlp
        return (self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias,
                self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias)

    def forward_keeps(self, x, embed_keep, layer_keep_in, keeps, dp_scale=None, dp_off=0):
        """x [B,N,C] fp32.  keeps = self.draw(B).  dp_scale: optional fp32 device table whose rows dp_off and dp_off+1
        hold the per-sample drop-path scales of the attention and MLP branches.  Returns (x, current_layer_keep)."""
        B, N, C = x.shape
        meta_a, meta_m, cur = self.half_metas(B, N, C, embed_keep, layer_keep_in, keeps, dp_scale, dp_off)
        return run_half_blocks(x, [meta_a, meta_m], self.half_params()), cur

    def forward(self, x, embed_mask=None, layer_mask=None):
        """Reference signature (:209-255): masks are [B,1,C] bool prefix masks; returns (x, embed_mask, current_layer_mask)."""
        B = x.shape[0]
        keeps = self.draw(B, x)
        dp = None
        p = getattr(self.drop_path, 'drop_prob', 0.) or 0.
        if self.training and p > 0.:
            dp = draw_scale(B, p, x.device, n=2)
        x, cur = self.forward_keeps(x, keep_of(embed_mask), keep_of(layer_mask), keeps, dp, 0)
        cur_mask = None if cur is None else make_mask(cur, self.dim, x.device)
        return x, embed_mask, cur_mask

    def rewiring(self):
        self.attn.rewiring()
        self.mlp.rewiring()
