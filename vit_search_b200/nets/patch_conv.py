"""PatchConvEmbed (conv stem) and the plain PatchEmbed -- drop-in for the reference's nets/patch_conv.py and for
timm 0.3.2's PatchEmbed (embed type 0).  Same sub-module names (conv1/conv2/conv3 = ConvBnAct{conv,bn,act},
conv_proj), hence the same state_dict keys and BatchNorm buffers.

Every convolution runs as im2col + the tcgen05 GEMM of libvsx.so; BatchNorm(train) statistics, the BN-apply + ReLU of
each layer (fused into the next layer's gather), the residual add, and the whole backward are hand-written kernels
(csrc/spatial.cu).  The stem is one autograd node.
"""
import torch
import torch.nn as nn

from .. import core, ops
from ..core import weights, _ActOperands, split_k_for

BN_MOMENTUM = 0.1


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class ConvBnAct(nn.Module):
    """Parameter holder with the reference's layout (nets/patch_conv.py:23-36); executed by _StemFn."""

    def __init__(self, in_channels, out_channels, kernel_size=(3, 3), padding=(1, 1), stride=(1, 1)):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=padding, stride=stride, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)
        self.act = nn.ReLU()


def _bn_train(y, P, C, bn, dt):
    sums = torch.zeros(2 * C, device=y.device, dtype=torch.float64)
    ops.call('bn_stats', y, dt, P, C, sums)
    return _bn_finalize(sums, P, C, bn)


def _bn_finalize(sums, P, C, bn):
    dev = sums.device
    scale, shift, mean, rstd = (torch.empty(C, device=dev) for _ in range(4))
    track = bn.track_running_stats and bn.running_mean is not None
    ops.call('bn_finalize', sums, P, C, bn.weight, bn.bias, float(bn.eps), BN_MOMENTUM if bn.momentum is None else float(bn.momentum),
             scale, shift, mean, rstd, bn.running_mean if track else None, bn.running_var if track else None,
             bn.num_batches_tracked if track else None)
    return scale, shift, mean, rstd


def _bn_eval(bn):
    rstd = torch.rsqrt(bn.running_var + bn.eps)
    scale = bn.weight.detach() * rstd
    return scale, bn.bias.detach() - bn.running_mean * scale, bn.running_mean, rstd


class _StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, w1, g1, b1, w2, g2, b2, w3, g3, b3, wp, bp):
        core.require_cuda(x, 'PatchConvEmbed')
        x = x.contiguous().float()
        B, Cin, H, W = x.shape
        T = core.act_dtype()
        dt = ops._DT[T]
        dev = x.device
        Cm = w1.shape[0]
        H1, W1 = H // 2, W // 2
        P = B * H1 * W1
        k = wp.shape[-1]
        Hp, Wp = H1 // k, W1 // k
        C = wp.shape[0]
        acts = _ActOperands()
        bns = (mod.conv1.bn, mod.conv2.bn, mod.conv3.bn)
        train = mod.training

        def conv(a_col, kdim, w, y):
            wc = weights.get(w, 'ohwi')
            a = acts.get(a_col, a_col.shape[1], 0, a_col.shape[0], kdim)
            ops.gemm(a, wc, a_col.shape[1], core.ld_of(wc), a_col.shape[0], y.shape[1], kdim, ops.EPI_STORE, y, y.shape[1])

        k1 = 9 * Cin
        y1 = torch.empty(P, Cm, device=dev, dtype=T)
        # conv1 as one kernel (csrc/conv1.cu): image -> y1 (+ its batch statistics), no im2col matrix in HBM
        direct1 = T == torch.bfloat16 and Cin == 3 and Cm == 24 and H % 2 == 0 and W % 2 == 0 and x.dtype == torch.float32
        if direct1:
            A1 = None
            wc1 = weights.get(w1, 'ohwi')
            sums1 = torch.zeros(2 * Cm, device=dev, dtype=torch.float64) if train else None
            ops.call('conv1_fwd', x, wc1, core.ld_of(wc1), y1, B, H, W, sums1)
            s1 = _bn_finalize(sums1, P, Cm, bns[0]) if train else _bn_eval(bns[0])
        else:
            A1 = torch.empty(P, core.up8(k1), device=dev, dtype=T)
            ops.call('im2col', x, None, None, None, None, None, ops.F32, 1, Cin * H * W, 1, B, H, W, Cin, 3, 2, 1, A1, dt, A1.shape[1])
            conv(A1, k1, w1, y1)
            s1 = _bn_train(y1, P, Cm, bns[0], dt) if train else _bn_eval(bns[0])
        direct = T == torch.bfloat16 and Cm % 8 == 0 and H1 % 8 == 0 and W1 % 16 == 0     # tensor-core direct conv (csrc/conv3x3.cu)
        A2 = A3 = None
        y2 = torch.empty(P, Cm, device=dev, dtype=T)
        y3 = torch.empty(P, Cm, device=dev, dtype=T)

        def conv3(y_in, s_in, w, y_out, bn):
            """3x3 conv of relu(bn(y_in)) with this layer's batch statistics fused into the epilogue."""
            sums = torch.zeros(2 * Cm, device=dev, dtype=torch.float64) if train else None
            ops.call('conv3x3', y_in, s_in[0], s_in[1], weights.get(w, 'conv3x3_fwd'), None, y_out, B, H1, W1, Cm, 1 if train else 0,
                     None, None, None, None, None, sums)
            if not train:
                return _bn_eval(bn)
            return _bn_finalize(sums, P, Cm, bn)

        if direct:
            s2 = conv3(y1, s1, w2, y2, bns[1])
            s3 = conv3(y2, s2, w3, y3, bns[2])
        else:
            A2 = torch.empty(P, 9 * Cm, device=dev, dtype=T)
            ops.call('im2col', y1, s1[0], s1[1], None, None, None, dt, 0, H1 * W1 * Cm, Cm, B, H1, W1, Cm, 3, 1, 1, A2, dt, 9 * Cm)
            conv(A2, 9 * Cm, w2, y2)
            s2 = _bn_train(y2, P, Cm, bns[1], dt) if train else _bn_eval(bns[1])
            A3 = torch.empty(P, 9 * Cm, device=dev, dtype=T)
            ops.call('im2col', y2, s2[0], s2[1], None, None, None, dt, 0, H1 * W1 * Cm, Cm, B, H1, W1, Cm, 3, 1, 1, A3, dt, 9 * Cm)
            conv(A3, 9 * Cm, w3, y3)
            s3 = _bn_train(y3, P, Cm, bns[2], dt) if train else _bn_eval(bns[2])
        kp = k * k * Cm
        A4 = torch.empty(B * Hp * Wp, core.up8(kp), device=dev, dtype=T)
        ops.call('im2col', y3, s3[0], s3[1], y1, s1[0], s1[1], dt, 0, H1 * W1 * Cm, Cm, B, H1, W1, Cm, k, k, 0, A4, dt, A4.shape[1])
        out = torch.empty(B * Hp * Wp, C, device=dev)
        wpc = weights.get(wp, 'ohwi')
        a4 = acts.get(A4, A4.shape[1], 0, A4.shape[0], kp)
        ops.gemm(a4, wpc, A4.shape[1], core.ld_of(wpc), A4.shape[0], C, kp, ops.EPI_STORE, out, C, bias=bp)
        ctx.save_for_backward(w1, g1, b1, w2, g2, b2, w3, g3, b3, wp)
        ctx.stuff = (A1 if A1 is not None else x, A2, A3, A4, y1, y2, y3, s1, s2, s3, (B, Cin, H1, W1, Cm, k, Hp, Wp, C, k1, kp, A1 is None))
        return out.view(B, Hp * Wp, C)

    @staticmethod
    def backward(ctx, g):
        w1, g1, b1, w2, g2, b2, w3, g3, b3, wp = ctx.saved_tensors
        A1, A2, A3, A4, y1, y2, y3, s1, s2, s3, dims = ctx.stuff
        ctx.stuff = None
        B, Cin, H1, W1, Cm, k, Hp, Wp, C, k1, kp, direct1 = dims
        T = core.act_dtype()
        dt = ops._DT[T]
        dev = g.device
        P = B * H1 * W1
        R = B * Hp * Wp
        acts = _ActOperands()
        g2d = g.contiguous().view(R, C)
        dtok = torch.empty(R, C, device=dev, dtype=T)
        d_bp = torch.zeros(C, device=dev)
        ops.scale_mask_cast(g2d, C, None, 1, C, dtok, C, R, C, colsum=d_bp)        # cast + conv_proj bias gradient in one pass

        def wgrad(dy, a_col, kdim, like):
            """dW (ohwi layout [O, kdim]) = dy^T a_col, then back to the parameter's [O, I, kh, kw] layout."""
            O = dy.shape[1]
            ldw = (kdim + 3) // 4 * 4                  # 16-byte row pitch for the TMA reduce-add
            dw = torch.zeros(O, ldw, device=dev)
            a = acts.get(dy, O, 0, dy.shape[0], O)
            bcol = acts.get(a_col, a_col.shape[1], 0, a_col.shape[0], kdim)
            ops.gemm(a, bcol, O, a_col.shape[1], O, kdim, dy.shape[0], ops.EPI_ATOMIC, dw, ldw, a_layout=ops.MNMAJOR,
                     b_layout=ops.MNMAJOR, split_k=split_k_for(O, kdim, dy.shape[0]))
            kh = like.shape[-1]
            return dw[:, :kdim].reshape(O, kh, kh, like.shape[1]).permute(0, 3, 1, 2).contiguous()

        def dgrad(dy, w, kdim, out):
            wc = weights.get(w, 'ohwi')
            a = acts.get(dy, dy.shape[1], 0, dy.shape[0], dy.shape[1])
            ops.gemm(a, wc, dy.shape[1], core.ld_of(wc), dy.shape[0], kdim, dy.shape[1], ops.EPI_STORE, out, out.shape[1],
                     n_out=core.up8(kdim), b_layout=ops.MNMAJOR)

        def bn_bwd(da, y, gam, bet, st):
            sums = torch.zeros(2 * Cm, device=dev, dtype=torch.float64)
            ops.call('bn_bwd_stats', da, y, dt, P, Cm, gam, bet, st[2], st[3], sums)
            dy = torch.empty(P, Cm, device=dev, dtype=T)
            dgam, dbet = torch.zeros(Cm, device=dev), torch.zeros(Cm, device=dev)
            ops.call('bn_bwd_apply', da, y, dt, P, Cm, gam, bet, st[2], st[3], sums, dy, dgam, dbet)
            return dy, dgam, dbet

        d_wp = wgrad(dtok, A4, kp, wp)
        dA4 = torch.empty(R, A4.shape[1], device=dev, dtype=T)
        dgrad(dtok, wp, kp, dA4)
        d_out = torch.empty(P, Cm, device=dev, dtype=T)      # gradient of (a3 + a1)
        ops.call('col2im', dA4, A4.shape[1], None, dt, B, H1, W1, Cm, k, k, 0, d_out, H1 * W1 * Cm, Cm)
        dy3, d_g3, d_b3 = bn_bwd(d_out, y3, g3, b3, s3)
        d_a2 = torch.empty(P, Cm, device=dev, dtype=T)
        d_a1 = torch.empty(P, Cm, device=dev, dtype=T)
        if A3 is None:     # direct tensor-core convs: weight gradient, and data gradient with the BN-backward reductions fused
            def wgrad3(dy, y_in, s_in, like):
                dw = torch.zeros(Cm, 9 * Cm, device=dev)
                ops.call('conv3x3_wgrad', dy, y_in, s_in[0], s_in[1], dw, B, H1, W1, Cm)
                return dw.view(Cm, 3, 3, like.shape[1]).permute(0, 3, 1, 2).contiguous()

            def dgrad3(dy, w, add, d_a, y_prev, gam, bet, st):
                sums = torch.zeros(2 * Cm, device=dev, dtype=torch.float64)
                ops.call('conv3x3', dy, None, None, weights.get(w, 'conv3x3_bwd'), add, d_a, B, H1, W1, Cm, 2, y_prev, gam, bet, st[2],
                         st[3], sums)
                dyp = torch.empty(P, Cm, device=dev, dtype=T)
                dgam, dbet = torch.zeros(Cm, device=dev), torch.zeros(Cm, device=dev)
                ops.call('bn_bwd_apply', d_a, y_prev, dt, P, Cm, gam, bet, st[2], st[3], sums, dyp, dgam, dbet)
                return dyp, dgam, dbet

            d_w3 = wgrad3(dy3, y2, s2, w3)
            dy2, d_g2, d_b2 = dgrad3(dy3, w3, None, d_a2, y2, g2, b2, s2)
            d_w2 = wgrad3(dy2, y1, s1, w2)
            dy1, d_g1, d_b1 = dgrad3(dy2, w2, d_out, d_a1, y1, g1, b1, s1)
        else:
            d_w3 = wgrad(dy3, A3, 9 * Cm, w3)
            dA = torch.empty(P, 9 * Cm, device=dev, dtype=T)
            dgrad(dy3, w3, 9 * Cm, dA)
            ops.call('col2im', dA, 9 * Cm, None, dt, B, H1, W1, Cm, 3, 1, 1, d_a2, H1 * W1 * Cm, Cm)
            dy2, d_g2, d_b2 = bn_bwd(d_a2, y2, g2, b2, s2)
            d_w2 = wgrad(dy2, A2, 9 * Cm, w2)
            acts.invalidate(dA)
            dgrad(dy2, w2, 9 * Cm, dA)
            ops.call('col2im', dA, 9 * Cm, d_out, dt, B, H1, W1, Cm, 3, 1, 1, d_a1, H1 * W1 * Cm, Cm)
            dy1, d_g1, d_b1 = bn_bwd(d_a1, y1, g1, b1, s1)
        if direct1:            # A1 holds the image: weight gradient straight from it (csrc/conv1.cu)
            dw1 = torch.zeros(Cm, 28, device=dev)
            ops.call('conv1_wgrad', A1, dy1, dw1, 28, B, 2 * H1, 2 * W1)
            d_w1 = dw1[:, :k1].reshape(Cm, 3, 3, Cin).permute(0, 3, 1, 2).contiguous()
        else:
            d_w1 = wgrad(dy1, A1, k1, w1)
        return (None, None, d_w1, d_g1, d_b1, d_w2, d_g2, d_b2, d_w3, d_g3, d_b3, d_wp, d_bp)


class PatchConvEmbed(nn.Module):
    def __init__(self, embed_dim, img_size=224, patch_size=14, in_chans=3, mid_chans=24):
        super().__init__()
        img_size, patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.patch_grid = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.patch_grid[0] * self.patch_grid[1]
        assert mid_chans % 4 == 0 and mid_chans <= 32, 'stem kernels support mid_chans in {4..32}, multiple of 4'
        self.conv1 = ConvBnAct(in_chans, mid_chans, stride=(2, 2))
        self.conv2 = ConvBnAct(mid_chans, mid_chans)
        self.conv3 = ConvBnAct(mid_chans, mid_chans)
        assert self.patch_size[0] % 2 == 0 and self.patch_size[1] % 2 == 0 and self.patch_size[0] == self.patch_size[1]
        self.conv_proj = nn.Conv2d(mid_chans, embed_dim, kernel_size=(patch_size[0] // 2, patch_size[1] // 2),
                                   stride=(patch_size[0] // 2, patch_size[1] // 2))

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1]
        c1, c2, c3 = self.conv1, self.conv2, self.conv3
        return _StemFn.apply(self, x, c1.conv.weight, c1.bn.weight, c1.bn.bias, c2.conv.weight, c2.bn.weight, c2.bn.bias,
                             c3.conv.weight, c3.bn.weight, c3.bn.bias, self.conv_proj.weight, self.conv_proj.bias)


class _PatchProjFn(torch.autograd.Function):
    """Conv2d(k = stride = patch) on the image as im2col + GEMM (timm PatchEmbed, embed type 0)."""

    @staticmethod
    def forward(ctx, x, w, b):
        core.require_cuda(x, 'PatchEmbed')
        x = x.contiguous().float()
        B, Cin, H, W = x.shape
        k = w.shape[-1]
        T = core.act_dtype()
        dt = ops._DT[T]
        R, kd = B * (H // k) * (W // k), Cin * k * k
        A = torch.empty(R, core.up8(kd), device=x.device, dtype=T)
        ops.call('im2col', x, None, None, None, None, None, ops.F32, 1, Cin * H * W, 1, B, H, W, Cin, k, k, 0, A, dt, A.shape[1])
        out = torch.empty(R, w.shape[0], device=x.device)
        wc = weights.get(w, 'ohwi')
        ops.gemm(_ActOperands().get(A, A.shape[1], 0, R, kd), wc, A.shape[1], core.ld_of(wc), R, w.shape[0], kd, ops.EPI_STORE, out,
                 w.shape[0], bias=b)
        ctx.save_for_backward(w)
        ctx.stuff = (A, kd, B)
        return out.view(B, -1, w.shape[0])

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        A, kd, B = ctx.stuff
        C = w.shape[0]
        R = A.shape[0]
        T = core.act_dtype()
        acts = _ActOperands()
        dtok = torch.empty(R, C, device=g.device, dtype=T)
        db = torch.zeros(C, device=g.device)
        ops.scale_mask_cast(g.contiguous().view(R, C), C, None, 1, C, dtok, C, R, C, colsum=db)
        ldw = (kd + 3) // 4 * 4
        dw = torch.zeros(C, ldw, device=g.device)
        ops.gemm(acts.get(dtok, C, 0, R, C), acts.get(A, A.shape[1], 0, R, kd), C, A.shape[1], C, kd, R, ops.EPI_ATOMIC, dw, ldw,
                 a_layout=ops.MNMAJOR, b_layout=ops.MNMAJOR, split_k=split_k_for(C, kd, R))
        k = w.shape[-1]
        return None, dw[:, :kd].reshape(C, k, k, w.shape[1]).permute(0, 3, 1, 2).contiguous(), db


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size, patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0])
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return _PatchProjFn.apply(x, self.proj.weight, self.proj.bias)
