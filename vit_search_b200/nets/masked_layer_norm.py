"""MaskedLayerNorm -- drop-in for the reference's nets/masked_layer_norm.py (same constructor, parameters
`weight` / `bias`, `forward(x, mask=None)`), backed by the one-pass sm_100a kernels in csrc/norm.cu."""
import torch
import torch.nn as nn

from .. import core, ops
from ._masks import keep_of


class MaskedLayerNormFunc(torch.autograd.Function):
    """y = LN over each sample's kept prefix, zero elsewhere (reference :23-50 + :124; backward :55-88).
    Output is fp32 like the reference's (it is forced to fp32 under autocast, :22)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, keep):
        core.require_cuda(x, 'MaskedLayerNorm')
        shape = x.shape
        C = shape[-1]
        B = shape[0]
        x2 = x.contiguous().float().view(-1, C)
        rows = x2.shape[0]
        rps = rows // B
        y = torch.empty_like(x2)
        mean = torch.empty(rows, device=x.device)
        rstd = torch.empty(rows, device=x.device)
        for b0, b1, k in _runs(keep, B, C):
            off = b0 * rps
            ops.masked_ln_fwd(x2, C, weight, bias, y, C, mean, rstd, (b1 - b0) * rps, C, k, eps, x_off=off * C, y_off=off * C,
                              stat_off=off)
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.keep, ctx.shape, ctx.rps = keep, shape, rps
        return y.view(shape)

    @staticmethod
    def backward(ctx, g):
        x2, weight, mean, rstd = ctx.saved_tensors
        C = x2.shape[1]
        B = ctx.shape[0]
        g2 = g.contiguous().float().view(-1, C)
        gx = torch.empty_like(x2)
        dw, db = torch.zeros_like(weight), torch.zeros_like(weight)
        for b0, b1, k in _runs(ctx.keep, B, C):
            off = b0 * ctx.rps
            ops.masked_ln_bwd(g2, C, x2, C, mean, rstd, weight, None, gx, C, dw, db, (b1 - b0) * ctx.rps, C, k, dy_off=off * C,
                              x_off=off * C, stat_off=off, g_off=off * C)
        return gx.view(ctx.shape), dw, db, None, None


def _runs(keep, B, C):
    if keep is None:
        return [(0, B, C)]
    out, b = [], 0
    while b < B:
        e = b
        while e < B and keep[e] == keep[b]:
            e += 1
        out.append((b, e, keep[b]))
        b = e
    return out


class MaskedLayerNorm(nn.Module):
    def __init__(self, num_channels, eps=1e-6):
        super().__init__()
        self.register_parameter('weight', nn.Parameter(torch.ones(num_channels)))
        self.register_parameter('bias', nn.Parameter(torch.zeros(num_channels)))
        self.eps = eps
        self.num_channels = num_channels
        self.normalized_shape = (num_channels,)

    def forward(self, x, mask=None):
        """x [B,N,C]; mask [B,1,C] bool prefix mask or None (None = plain LayerNorm, reference :119-122)."""
        return MaskedLayerNormFunc.apply(x, self.weight, self.bias, self.eps, keep_of(mask))

    def extra_repr(self):
        return 'num_channels={}, eps={}'.format(self.num_channels, self.eps)
