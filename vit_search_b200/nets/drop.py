"""DropPath (stochastic depth per sample) -- drop-in for the reference's nets/drop.py.

Inside Block the per-sample scale `floor(keep_prob + U) / keep_prob` is folded into the proj / fc2 GEMM epilogue;
this module form exists for stand-alone use and draws U on the CUDA generator like the reference (:20-22)."""
import torch
import torch.nn as nn

from .. import core, ops


def draw_scale(batch, drop_prob, device, n=1):
    """[n, batch] fp32 table of per-sample scales: floor(keep_prob + U) / keep_prob  (nets/drop.py:18-25)."""
    keep_prob = 1.0 - drop_prob
    u = torch.rand((n, batch), dtype=torch.float32, device=device)
    return (keep_prob + u).floor_().div_(keep_prob)


class _RowScaleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        core.require_cuda(x, 'DropPath')
        x = x.contiguous().float()
        ctx.scale = scale
        return _scale(x, scale)

    @staticmethod
    def backward(ctx, g):
        return _scale(g.contiguous().float(), ctx.scale), None


def _scale(x, scale):
    B, C = x.shape[0], x.shape[-1]
    rows_per = x.numel() // (B * C)
    y = torch.empty_like(x)
    ops.scale_mask_cast(x, C, scale, rows_per, C, y, C, B * rows_per, C)
    return y


def drop_path(x, drop_prob: float = 0., training: bool = False):
    if drop_prob == 0. or not training:
        return x
    return _RowScaleFn.apply(x, draw_scale(x.shape[0], drop_prob, x.device).view(-1))


class DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)

    def extra_repr(self):
        return 'drop_prob={}'.format(self.drop_prob)
