"""ChannelDrop -- sub-architecture sampling by prefix channel masks.

Drop-in for the reference's nets/channel_drop.py (same constructor, attributes and methods).  The table logic and
the per-step draw protocol follow the reference line by line in behaviour (table: :114-157, draw: :93-111, epoch
reset: :160-162, eval all-true mask: :84-88) but work on integer keep counts on the host; one `torch.randperm` on
the global CPU generator is consumed per training forward exactly like the reference, so seeding
`torch.manual_seed(epoch*10000+iter)` (engine.py:122) selects the same sub-architectures.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from .. import core, ops
from ._masks import make_mask

_NUM_WARMUP_EPOCHS_CHANNEL = 5


class _PrefixMaskFn(torch.autograd.Function):
    """y[b] = x[b] * prefix_mask(keep[b]) through the scale_mask_cast kernel, segment by segment."""

    @staticmethod
    def forward(ctx, x, keep):
        core.require_cuda(x, 'ChannelDrop')
        x = x.contiguous().float()
        ctx.keep = keep
        return _apply(x, keep)

    @staticmethod
    def backward(ctx, g):
        return _apply(g.contiguous().float(), ctx.keep), None


def _apply(x, keep):
    B = x.shape[0]
    C = x.shape[-1]
    rows_per = x.numel() // (B * C)
    y = torch.empty_like(x)
    b = 0
    while b < B:
        e = b
        while e < B and keep[e] == keep[b]:
            e += 1
        off = b * rows_per * C
        ops.scale_mask_cast(x, C, None, 1, keep[b], y, C, (e - b) * rows_per, C, g_off=off, out_off=off)
        b = e
    return y


class ChannelDrop(nn.Module):
    def __init__(self, num_channels_to_keep=None, num_warmup_epochs=_NUM_WARMUP_EPOCHS_CHANNEL, example_per_arch=None,
                 single_arch=False):
        super().__init__()
        assert num_channels_to_keep is not None
        assert example_per_arch is not None
        assert isinstance(num_channels_to_keep, np.ndarray), 'num_channels_to_keep data type error'
        self.num_channels_to_keep = np.sort(num_channels_to_keep)[::-1]
        self.epoch_now = None
        self.num_warmup_epochs = num_warmup_epochs
        self.example_per_arch = example_per_arch
        self.single_arch = single_arch
        self.keep_table = None        # host-side table: keep count of every mask row
        self.mask = None              # materialised lazily, only if somebody asks for the reference's attribute
        self.mask_all_true = None
        self.num_layer_config = None
        self.fixed_mask = None
        self.fixed_keep = None

    # ------------------------------------------------------------------ table (per epoch)
    def set_mask(self, inputs):
        B, C = inputs.shape[0], inputs.shape[-1]
        ch = [int(c) for c in self.num_channels_to_keep]
        assert B % self.example_per_arch == 0, 'Batch size is not divisible by sub-batch size (examples per arch).'
        assert all(c <= C for c in ch), 'Some elements in num_channels_to_keep is larger than channel size.'
        assert max(ch) == C, 'Maximum channel not in num_channels_to_keep'
        assert B >= len(ch), 'The batch size is smaller than the number of channels to keep.'
        if self.num_warmup_epochs == 0:
            n = len(ch)
        else:
            n = min(1 + math.floor(self.epoch_now * (len(ch) - 1) / self.num_warmup_epochs), len(ch))
            n = max(n, 1)
        self.num_layer_config = n
        cycles = 1 if self.single_arch else math.ceil((B // self.example_per_arch) / n)
        self.keep_table = [ch[r % n] for r in range(n * cycles)]
        self.mask = None

    def table_mask(self, device='cuda'):
        """The reference's `self.mask` tensor ([rows,1,C] bool), built on demand."""
        if self.mask is None and self.keep_table is not None:
            self.mask = make_mask(self.keep_table, int(max(self.num_channels_to_keep)), device)
        return self.mask

    # ------------------------------------------------------------------ draw (per step)
    def draw(self, batch, width, like=None):
        """Per-sample keep counts for one training forward; consumes one CPU randperm like the reference."""
        if self.fixed_keep is not None:
            return [self.fixed_keep] * batch
        if self.keep_table is None:
            self.set_mask(torch.empty(batch, 1, width, device='meta'))      # the table depends on (batch, width) only; `like` may be a narrower tensor
        perm = torch.randperm(len(self.keep_table)).tolist()
        if self.single_arch:
            return [self.keep_table[perm[0]]] * batch
        assert batch % self.example_per_arch == 0, 'In forward(), batch size is not divisible by sub-batch size (examples per arch).'
        g = batch // self.example_per_arch
        return [self.keep_table[perm[i % g]] for i in range(batch)]

    def keeps(self, batch, width, like=None):
        """Keep counts for the current mode: a fresh draw in training, all channels in eval (:84-88)."""
        if self.fixed_keep is not None:
            return [self.fixed_keep] * batch
        if self.training:
            return self.draw(batch, width, like)
        return [width] * batch

    def forward(self, x):
        """x [B,N,C] -> (masked x, mask [B,1,C] bool) -- stand-alone use, as in the reference's signature."""
        B, C = x.shape[0], x.shape[-1]
        keep = self.keeps(B, C, x)
        mask = make_mask(keep, C, x.device)
        if self.training or self.fixed_keep is not None:
            x = _PrefixMaskFn.apply(x, keep)
        return x, mask

    def forward_mask(self, x):
        return make_mask(self.draw(x.shape[0], x.shape[-1], x), x.shape[-1], x.device)

    # ------------------------------------------------------------------ epoch / debug hooks
    def set_epoch(self, epoch_now):
        self.epoch_now = epoch_now
        self.reset_mask()

    def reset_mask(self):
        self.mask = None
        self.keep_table = None
        self.mask_all_true = None
        self.fixed_mask = None
        self.fixed_keep = None
        self.num_layer_config = None

    def set_fixed_mask(self, mask):
        assert len(mask.shape) == 3 and mask.shape[0] == 1
        from ._masks import keep_of
        self.fixed_keep = keep_of(mask)[0]
        self.fixed_mask = mask

    def set_random_fixed_mask(self):
        width = int(max(self.num_channels_to_keep))
        if self.keep_table is None:
            self.set_mask(torch.empty(self.example_per_arch * len(self.num_channels_to_keep), 1, width, device='meta'))
        perm = torch.randperm(len(self.keep_table)).tolist()
        self.fixed_keep = self.keep_table[perm[0]]
        self.fixed_mask = None

    def extra_repr(self):
        s = 'num_channels_to_keep={}, num_warmup_epochs={}, example_per_arch={}'.format(
            self.num_channels_to_keep, self.num_warmup_epochs, self.example_per_arch)
        if self.single_arch:
            s += ', single_arch={}'.format(self.single_arch)
        return s
