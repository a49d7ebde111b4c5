"""Host-side orchestration of the kernels: precision modes, weight operand cache, segments, and the fused
half-block forward/backward routines (attention half, MLP half) used by nets/supernet_blocks.py.

All arithmetic happens in libvsx.so (see ops.py).  torch is used for allocation, streams and autograd wiring.
"""
import math
import weakref
from contextlib import contextmanager

import torch

from . import ops

# ------------------------------------------------------------------------------------------------ precision
# 'bf16' : activations stored in bf16, single-term bf16 tensor-core GEMMs (the training path; bench.py).
# 'fp32' : activations stored in fp32, every GEMM operand split into three bf16 parts (x = x1+x2+x3, 24 mantissa bits)
#          and contracted as six tensor-core terms (all cross products down to 2^-24), attention in fp32 math.  Same
#          kernels, fp32-level accuracy -- the path that proves parity with the fp32 reference to the north-star's 1e-3.
_precision = 'bf16'


def set_precision(p):
    global _precision
    assert p in ('bf16', 'fp32')
    _precision = p


def get_precision():
    return _precision


@contextmanager
def precision(p):
    old = get_precision()
    set_precision(p)
    try:
        yield
    finally:
        set_precision(old)


def act_dtype():
    return torch.bfloat16 if _precision == 'bf16' else torch.float32


def h2d(data, device, dtype=None):
    """Small host -> device upload that never synchronises the stream: the data goes through a pinned staging tensor (PyTorch's
    caching host allocator recycles it only after the copy has completed) and an asynchronous copy.  A plain
    `torch.tensor(data, device=...)` / `.to(device)` from pageable memory blocks the Python thread until the GPU has drained,
    which keeps the launch queue empty at the start of every step."""
    t = data if isinstance(data, torch.Tensor) else torch.tensor(data, dtype=dtype)
    if torch.device(device).type != 'cuda':
        return t.to(device)
    return t.pin_memory().to(device, non_blocking=True)


def require_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError('%s: vit_search_b200 runs on a B200 through libvsx.so only -- got a %s tensor; there is no CPU '
                           'fallback (the CPU restatement lives in oracle/ and is test-only)' % (what, getattr(t, 'device', type(t))))


# ------------------------------------------------------------------------------------------------ operand preparation
class _WeightCache:
    """bf16 operand copies of fp32 parameters.  An entry is valid only for the SAME live parameter object (weak reference:
    Python re-uses ids and the allocator re-uses addresses once a model dies), the same storage pointer (`rewiring`
    re-allocates weights, nets/supernet_blocks.py:55-71,123-161), the same autograd version counter (in-place optimizer
    updates) and the same `generation` (bumped by every model forward and by optimizers that write through raw pointers), so
    a weight is cast once per step and a stale copy can never be served."""

    def __init__(self):
        self.entries = {}
        self.generation = 0          # bumped by optimizers that update parameters through raw pointers

    def get(self, w, layout=None):
        key = (id(w), layout)
        tag = (w.data_ptr(), w._version, self.generation, _precision, tuple(w.shape))
        e = self.entries.get(key)
        if e is not None and e[2]() is w:
            if e[0] == tag:
                return e[1]
            # shadow copies are rewritten by the fused optimizer kernel in the same pass that updates the fp32 master, so they do
            # not depend on `generation`; everything else in the tag must still match
            if len(e) == 4 and e[0][:2] == tag[:2] and e[0][3:] == tag[3:]:
                return e[1]
        ref = weakref.ref(w)
        src = w.detach()
        if layout in ('conv3x3_fwd', 'conv3x3_bwd'):
            # direct-conv operand [9 taps][32 n][32 k] bf16, zero padded (csrc/conv3x3.cu).  fwd: n = out ch, k = in ch;
            # bwd (data gradient): taps flipped, n = in ch, k = out ch
            o, i = src.shape[0], src.shape[1]
            t9 = src.permute(2, 3, 0, 1).reshape(9, o, i) if layout == 'conv3x3_fwd' else src.flip(2, 3).permute(2, 3, 1, 0).reshape(9, i, o)
            val = torch.zeros(9, 32, 32, device=w.device, dtype=torch.bfloat16)
            val[:, :t9.shape[1], :t9.shape[2]] = t9.to(torch.bfloat16)
            self.entries[key] = (tag, val, ref)
            return val
        if layout == 'ohwi':                       # conv weight [O, I, kh, kw] -> [O, kh*kw*I]
            src = src.permute(0, 2, 3, 1)
        src = src.reshape(src.shape[0], -1).contiguous()
        rows, cols = src.shape
        colsp = (cols + 7) // 8 * 8                # TMA needs 16-byte row pitch
        hi = torch.zeros(rows, colsp, device=w.device, dtype=torch.bfloat16) if colsp != cols else \
            torch.empty(rows, cols, device=w.device, dtype=torch.bfloat16)
        lo = torch.zeros_like(hi) if _precision == 'fp32' else None
        lo2 = torch.zeros_like(hi) if _precision == 'fp32' else None
        ops.split_bf16(src, cols, hi, lo, colsp, rows, cols, lo2=lo2) if cols % 4 == 0 else _split_slow(src, hi, lo, lo2)
        val = (hi, lo, lo2) if lo is not None else hi
        if len(self.entries) > 4096:            # dead models leave entries behind: drop them wholesale
            self.entries = {k: v for k, v in self.entries.items() if v[2]() is not None}
        self.entries[key] = (tag, val, ref)
        return val

    def adopt_shadow(self, w, shadow):
        """Register `shadow` (bf16, same [rows, cols] layout) as the GEMM operand copy of parameter w.  The caller (FusedAdamW)
        guarantees that it rewrites the shadow whenever it rewrites w through raw pointers."""
        tag = (w.data_ptr(), w._version, self.generation, 'bf16', tuple(w.shape))
        self.entries[(id(w), None)] = (tag, shadow, weakref.ref(w), True)

    def clear(self):
        self.entries.clear()


def _split_slow(src, hi, lo, lo2):
    """Widths that are not a multiple of 4 (only conv1's [mid, 27] weight): same split with torch ops, once per step."""
    cols = src.shape[1]
    hi[:, :cols] = src.to(torch.bfloat16)
    if lo is not None:
        r = src - hi[:, :cols].float()
        lo[:, :cols] = r.to(torch.bfloat16)
        lo2[:, :cols] = (r - lo[:, :cols].float()).to(torch.bfloat16)


weights = _WeightCache()


def ld_of(op):
    return (op[0] if isinstance(op, tuple) else op).shape[-1]


class _ActOperands:
    """Per-call cache of activation GEMM operands.  bf16 mode: the tensor itself.  fp32 mode: hi/lo bf16 buffers of the
    same [rows, ld] shape, filled segment by segment right before use."""

    def __init__(self):
        self.buf = {}

    def get(self, t, ld, row0, rows, cols):
        if t.dtype == torch.bfloat16:
            return t
        key = id(t)
        if key not in self.buf:
            self.buf[key] = tuple(torch.empty(t.shape, device=t.device, dtype=torch.bfloat16) for _ in range(3)) + (set(),)
        hi, lo, lo2, done = self.buf[key]
        c4 = (cols + 3) // 4 * 4
        if (row0, rows, c4) not in done:
            ops.split_bf16(t, ld, hi, lo, ld, rows, min(c4, ld), src_off=row0 * ld, dst_off=row0 * ld, lo2=lo2)
            done.add((row0, rows, c4))
        return (hi, lo, lo2)

    def invalidate(self, t):
        self.buf.pop(id(t), None)


def split_k_for(m_rows, n_cols, red_rows):
    tiles = math.ceil(m_rows / 128) * math.ceil(n_cols / 128)
    kb = max(1, math.ceil(red_rows / 64))
    return max(1, min(kb, (2 * 148) // max(tiles, 1)))


def up8(n):
    return (n + 7) // 8 * 8


class _GradPool:
    """One zero-filled fp32 buffer per training step for ALL parameter gradients (one allocation + one memset instead of one per
    half block).  engine.TrainStep opens it right before `backward()`; outside of that window (tests, ad-hoc backward calls)
    zeros_like_many falls back to its own allocation."""

    def __init__(self):
        self.flat, self.off = None, 0

    def begin(self, numel, device):
        # one PERSISTENT buffer, cleared with a memset: gradient addresses are then the same in every step (the optimizer's pointer table
        # stays valid, the all-reduce always works on the same registered memory) and the caching allocator is not asked for 280 MB per step
        buf = getattr(self, 'buf', None)
        if buf is None or buf.numel() != numel or buf.device != torch.device(device):
            self.buf = buf = torch.zeros(numel, device=device, dtype=torch.float32)
        else:
            buf.zero_()
        self.flat, self.off = buf, 0

    def end(self):
        self.flat, self.off = None, 0

    def take(self, numel, device):
        f = self.flat
        if f is None or f.device != device or self.off + numel > f.numel():
            return None
        v = f[self.off:self.off + numel]
        self.off += numel
        return v


grad_pool = _GradPool()
trunk_grads_ready_hook = None    # set by engine.TrainStep for data-parallel runs (see nets/vit_sr_supernet.py forward_features)
pool_prefix_ready_hook = None    # set by engine.TrainStep for data-parallel runs: called (no arguments) whenever a stage's backward has been
                                 # queued, i.e. grad_pool.flat[:grad_pool.off] is final -- the staged gradient exchange starts there


def zeros_like_many(*tensors):
    """Zero-initialised fp32 gradient buffers for several parameters from ONE allocation + ONE memset (each view 16-byte
    aligned, as the TMA reduce-add of the weight-gradient GEMMs requires); carved out of the per-step gradient pool when
    engine.TrainStep has opened one."""
    sizes = [(t.numel() + 3) // 4 * 4 for t in tensors]
    flat = grad_pool.take(sum(sizes), tensors[0].device)
    if flat is None:
        flat = torch.zeros(sum(sizes), device=tensors[0].device, dtype=torch.float32)
    out, off = [], 0
    for t, n in zip(tensors, sizes):
        out.append(flat[off:off + t.numel()].view(t.shape))
        off += n
    return out


# ------------------------------------------------------------------------------------------------ segments
class Seg:
    """A run of consecutive samples sharing one sub-architecture in this layer."""
    __slots__ = ('b0', 'b1', 'ek', 'ik', 'ck', 'active')

    def __init__(self, b0, b1, ek, ik, ck, active=True):
        self.b0, self.b1 = b0, b1
        self.ek = ek          # embedding channels kept (input width of the branch)
        self.ik = ik          # inner width kept: heads*head_dim for attention, hidden channels for the MLP
        self.ck = ck          # channels of the branch output that reach the residual (layer & embed mask)
        self.active = active  # False: the whole block is dropped for these samples

    def key(self):
        return (self.ek, self.ik, self.ck, self.active)


def make_segments(batch, width, embed_keep, inner_keep, inner_full, cur_keep, bounds=None, grouped=False):
    """Per-sample keep lists (or None) -> list of Seg with maximal runs of identical keeps.  `bounds`: sample indices at which a new
    segment starts even if the extents do not change (the architecture-group boundaries of the batch: every half block of a step then
    has the SAME sample ranges, so consecutive half blocks can hand gradients over segment by segment).  grouped=True: the caller
    guarantees that the keeps only change at `bounds` (None: one architecture), so only the first sample of every group is inspected."""
    segs = []
    if grouped:
        starts = [0] + (sorted(bounds) if bounds else [])
        for i, b in enumerate(starts):
            ek = width if embed_keep is None else int(embed_keep[b])
            ik = inner_full if inner_keep is None else int(inner_keep[b])
            ck = width if cur_keep is None else int(cur_keep[b])
            segs.append(Seg(b, starts[i + 1] if i + 1 < len(starts) else batch, ek, ik, ck, ck > 0 and ik > 0 and ek > 0))
        return segs
    for b in range(batch):
        ek = width if embed_keep is None else int(embed_keep[b])
        ik = inner_full if inner_keep is None else int(inner_keep[b])
        ck = width if cur_keep is None else int(cur_keep[b])
        act = ck > 0 and ik > 0 and ek > 0
        if segs and segs[-1].key() == (ek, ik, ck, act) and not (bounds is not None and b in bounds):
            segs[-1].b1 = b + 1
        else:
            segs.append(Seg(b, b + 1, ek, ik, ck, act))
    return segs


class HalfMeta:
    """Static description of one half-block call."""

    def __init__(self, kind, segs, tokens, width, heads=0, head_dim=0, hidden=0, row_scale=None, scale_off=0,
                 pre_norm=True, residual=True, eps=1e-6):
        self.kind, self.segs, self.N, self.C = kind, segs, tokens, width
        self.H, self.D, self.F = heads, head_dim, hidden
        self.row_scale, self.scale_off = row_scale, scale_off
        self.pre_norm, self.residual, self.eps = pre_norm, residual, eps


# ------------------------------------------------------------------------------------------------ attention half
def _pre(meta, x2, g, b, xn, mean, rstd, s, r0, rows):
    C = meta.C
    if meta.pre_norm:
        ops.masked_ln_fwd(x2, C, g, b, xn, C, mean, rstd, rows, C, s.ek, meta.eps, x_off=r0 * C, y_off=r0 * C, stat_off=r0)
    else:
        ops.scale_mask_cast(x2, C, None, 1, s.ek, xn, C, rows, C, g_off=r0 * C, out_off=r0 * C)


def attn_half_forward(meta, x, ln_w, ln_b, qkv_w, qkv_b, proj_w, proj_b):
    B, N, C = x.shape
    M, H, D = B * N, meta.H, meta.D
    HD = H * D
    T = act_dtype()
    dev = x.device
    x2 = x.view(M, C)
    xn = torch.empty(M, C, device=dev, dtype=T)
    mean = torch.empty(M, device=dev)
    rstd = torch.empty(M, device=dev)
    qkv = torch.empty(M, 3 * HD, device=dev, dtype=T)
    o = torch.empty(M, HD, device=dev, dtype=T)
    lse = torch.empty(B, H, N, device=dev)
    out = torch.empty_like(x)
    out2 = out.view(M, C)
    wq, wp = weights.get(qkv_w), weights.get(proj_w)
    acts = _ActOperands()
    scale = D ** -0.5
    for s in meta.segs:
        r0, rows, nb = s.b0 * N, (s.b1 - s.b0) * N, s.b1 - s.b0
        if not s.active:
            if meta.residual:
                out[s.b0:s.b1].copy_(x[s.b0:s.b1])
            else:
                out[s.b0:s.b1].zero_()
            continue
        hk = s.ik // D
        _pre(meta, x2, ln_w, ln_b, xn, mean, rstd, s, r0, rows)
        a = acts.get(xn, C, r0, rows, s.ek)
        if hk == H:
            ops.gemm(a, wq, C, C, rows, 3 * HD, s.ek, ops.EPI_STORE, qkv, 3 * HD, a_off=r0 * C, out_off=r0 * 3 * HD, bias=qkv_b)
        else:
            # q / k / v row blocks of the kept heads only (features ordered (3,H,D)): three problems, one launch
            ops.gemm_grouped([((a, wq, C, C, rows, hk * D, s.ek, ops.EPI_STORE, qkv, 3 * HD),
                               dict(a_off=r0 * C, b_off=j * HD * C, out_off=r0 * 3 * HD + j * HD, bias=qkv_b, bias_off=j * HD)) for j in range(3)])
        ops.attn_fwd(qkv, o, lse, nb, N, H, D, hk, scale, qkv_off=r0 * 3 * HD, o_off=r0 * HD, lse_off=s.b0 * H * N)
        ao = acts.get(o, HD, r0, rows, hk * D)
        if meta.residual:
            ops.gemm(ao, wp, HD, HD, rows, s.ck, hk * D, ops.EPI_RESIDUAL, out2, C, a_off=r0 * HD, out_off=r0 * C, n_out=C,
                     bias=proj_b, aux=x2, ld_aux=C, aux_off=r0 * C, row_scale=meta.row_scale,
                     row_scale_off=meta.scale_off + s.b0, rows_per_sample=N, n_keep=s.ck)
        else:
            ops.gemm(ao, wp, HD, HD, rows, C, hk * D, ops.EPI_STORE, out2, C, a_off=r0 * HD, out_off=r0 * C, n_out=C, bias=proj_b)
    return out, (xn, mean, rstd, qkv, o, lse)


def attn_half_backward(meta, g_out, saved, x, ln_w, ln_b, qkv_w, qkv_b, proj_w, proj_b):
    xn, mean, rstd, qkv, o, lse = saved
    B, N, C = x.shape
    M, H, D = B * N, meta.H, meta.D
    HD = H * D
    T = act_dtype()
    dev = x.device
    x2, g2 = x.view(M, C), g_out.view(M, C)
    g_in = torch.empty_like(x)
    gi2 = g_in.view(M, C)
    df = torch.empty(M, C, device=dev, dtype=T)
    d_o = torch.empty(M, HD, device=dev, dtype=T)
    dqkv = torch.empty(M, 3 * HD, device=dev, dtype=T)
    dxn = torch.empty(M, C, device=dev, dtype=T) if meta.pre_norm else None
    d_lnw, d_lnb, d_qw, d_qb, d_pw, d_pb = zeros_like_many(ln_w, ln_b, qkv_w, qkv_b, proj_w, proj_b)
    wq, wp = weights.get(qkv_w), weights.get(proj_w)
    acts = _ActOperands()
    scale = D ** -0.5
    for s in meta.segs:
        r0, rows, nb = s.b0 * N, (s.b1 - s.b0) * N, s.b1 - s.b0
        if not s.active:
            if meta.residual:
                g_in[s.b0:s.b1].copy_(g_out[s.b0:s.b1])
            else:
                g_in[s.b0:s.b1].zero_()
            continue
        hk = s.ik // D
        hkd = hk * D
        ck = s.ck if meta.residual else C
        ops.scale_mask_cast(g2, C, meta.row_scale if meta.residual else None, N, ck, df, C, rows, C, g_off=r0 * C, out_off=r0 * C,
                            scale_off=meta.scale_off + s.b0, colsum=d_pb)
        a_df = acts.get(df, C, r0, rows, ck)
        a_o = acts.get(o, HD, r0, rows, hkd)
        # dWproj[ck, hkd] += df^T o : launched together with the qkv weight gradients below (one grouped launch per half block)
        wgrads = [((a_df, a_o, C, HD, ck, hkd, rows, ops.EPI_ATOMIC, d_pw, HD),
                   dict(a_off=r0 * C, b_off=r0 * HD, a_layout=ops.MNMAJOR, b_layout=ops.MNMAJOR, split_k=split_k_for(ck, hkd, rows)))]
        # d_o[rows, hkd] = df[rows, ck] Wproj[ck, hkd]
        ops.gemm(a_df, wp, C, HD, rows, hkd, ck, ops.EPI_STORE, d_o, HD, a_off=r0 * C, out_off=r0 * HD, b_layout=ops.MNMAJOR)
        ops.attn_bwd(qkv, o, d_o, lse, dqkv, nb, N, H, D, hk, scale, qkv_off=r0 * 3 * HD, o_off=r0 * HD, lse_off=s.b0 * H * N, dbias=d_qb)
        a_dq = acts.get(dqkv, 3 * HD, r0, rows, 3 * HD)
        a_xn = acts.get(xn, C, r0, rows, s.ek)
        for j in (range(3) if hk < H else range(1)):
            nrow = hkd if hk < H else 3 * HD
            wgrads.append(((a_dq, a_xn, 3 * HD, C, nrow, s.ek, rows, ops.EPI_ATOMIC, d_qw, C),
                           dict(a_off=r0 * 3 * HD + j * HD, b_off=r0 * C, out_off=j * HD * C, a_layout=ops.MNMAJOR, b_layout=ops.MNMAJOR,
                                split_k=split_k_for(nrow, s.ek, rows))))
        ops.gemm_grouped(wgrads)
        # dxn[rows, ek] = dqkv[rows, 3HD] Wqkv[3HD, ek]: the reduction walks the three windows of kept heads only
        # (the 64-wide k steps may overrun a window only into masked, i.e. zero, columns of dqkv)
        kseg = dict(k_segments=3, k_seg_len=hkd, k_seg_stride=HD) if (hk < H and (hkd + 63) // 64 * 64 <= HD) else {}
        if meta.pre_norm:
            ops.gemm(a_dq, wq, 3 * HD, C, rows, s.ek, 3 * HD, ops.EPI_STORE, dxn, C, a_off=r0 * 3 * HD, out_off=r0 * C,
                     n_out=up8(s.ek), b_layout=ops.MNMAJOR, **kseg)
            ops.masked_ln_bwd(dxn, C, x2, C, mean, rstd, ln_w, g2 if meta.residual else None, gi2, C, d_lnw, d_lnb, rows, C, s.ek,
                              dy_off=r0 * C, x_off=r0 * C, stat_off=r0, g_off=r0 * C)
        else:
            ops.gemm(a_dq, wq, 3 * HD, C, rows, s.ek, 3 * HD, ops.EPI_STORE, gi2, C, a_off=r0 * 3 * HD, out_off=r0 * C, n_out=C,
                     b_layout=ops.MNMAJOR, **kseg)
    return g_in, (d_lnw, d_lnb, d_qw, d_qb, d_pw, d_pb)


# ------------------------------------------------------------------------------------------------ MLP half
def mlp_half_forward(meta, x, ln_w, ln_b, fc1_w, fc1_b, fc2_w, fc2_b):
    B, N, C = x.shape
    M, F = B * N, meta.F
    T = act_dtype()
    dev = x.device
    x2 = x.view(M, C)
    xn = torch.empty(M, C, device=dev, dtype=T)
    mean = torch.empty(M, device=dev)
    rstd = torch.empty(M, device=dev)
    u = torch.empty(M, F, device=dev, dtype=T)
    h = torch.empty(M, F, device=dev, dtype=T)
    out = torch.empty_like(x)
    out2 = out.view(M, C)
    w1, w2 = weights.get(fc1_w), weights.get(fc2_w)
    acts = _ActOperands()
    for s in meta.segs:
        r0, rows = s.b0 * N, (s.b1 - s.b0) * N
        if not s.active:
            if meta.residual:
                out[s.b0:s.b1].copy_(x[s.b0:s.b1])
            else:
                out[s.b0:s.b1].zero_()
            continue
        _pre(meta, x2, ln_w, ln_b, xn, mean, rstd, s, r0, rows)
        a = acts.get(xn, C, r0, rows, s.ek)
        ops.gemm(a, w1, C, C, rows, s.ik, s.ek, ops.EPI_GELU, u, F, a_off=r0 * C, out_off=r0 * F, n_out=up8(s.ik), out2=h,
                 ldo2=F, out2_off=r0 * F, bias=fc1_b)
        ah = acts.get(h, F, r0, rows, s.ik)
        if meta.residual:
            ops.gemm(ah, w2, F, F, rows, s.ck, s.ik, ops.EPI_RESIDUAL, out2, C, a_off=r0 * F, out_off=r0 * C, n_out=C, bias=fc2_b,
                     aux=x2, ld_aux=C, aux_off=r0 * C, row_scale=meta.row_scale, row_scale_off=meta.scale_off + s.b0,
                     rows_per_sample=N, n_keep=s.ck)
        else:
            ops.gemm(ah, w2, F, F, rows, C, s.ik, ops.EPI_STORE, out2, C, a_off=r0 * F, out_off=r0 * C, n_out=C, bias=fc2_b)
    return out, (xn, mean, rstd, u, h)


def mlp_half_backward(meta, g_out, saved, x, ln_w, ln_b, fc1_w, fc1_b, fc2_w, fc2_b):
    xn, mean, rstd, u, h = saved
    B, N, C = x.shape
    M, F = B * N, meta.F
    T = act_dtype()
    dev = x.device
    x2, g2 = x.view(M, C), g_out.view(M, C)
    g_in = torch.empty_like(x)
    gi2 = g_in.view(M, C)
    df = torch.empty(M, C, device=dev, dtype=T)
    du = torch.empty(M, F, device=dev, dtype=T)
    dxn = torch.empty(M, C, device=dev, dtype=T) if meta.pre_norm else None
    d_lnw, d_lnb, d_w1, d_b1, d_w2, d_b2 = zeros_like_many(ln_w, ln_b, fc1_w, fc1_b, fc2_w, fc2_b)
    w1, w2 = weights.get(fc1_w), weights.get(fc2_w)
    acts = _ActOperands()
    for s in meta.segs:
        r0, rows = s.b0 * N, (s.b1 - s.b0) * N
        if not s.active:
            if meta.residual:
                g_in[s.b0:s.b1].copy_(g_out[s.b0:s.b1])
            else:
                g_in[s.b0:s.b1].zero_()
            continue
        ck = s.ck if meta.residual else C
        ops.scale_mask_cast(g2, C, meta.row_scale if meta.residual else None, N, ck, df, C, rows, C, g_off=r0 * C, out_off=r0 * C,
                            scale_off=meta.scale_off + s.b0, colsum=d_b2)
        a_df = acts.get(df, C, r0, rows, ck)
        a_h = acts.get(h, F, r0, rows, s.ik)
        # dW2[ck, ik] += df^T h : launched together with dW1 below
        wgrads = [((a_df, a_h, C, F, ck, s.ik, rows, ops.EPI_ATOMIC, d_w2, F),
                   dict(a_off=r0 * C, b_off=r0 * F, a_layout=ops.MNMAJOR, b_layout=ops.MNMAJOR, split_k=split_k_for(ck, s.ik, rows)))]
        # du[rows, ik] = (df[rows, ck] W2[ck, ik]) * gelu'(pre-activation): `u` holds that derivative (the fc1 epilogue stores it
        # instead of the pre-activation)
        ops.gemm(a_df, w2, C, F, rows, s.ik, ck, ops.EPI_GELUGRAD, du, F, a_off=r0 * C, out_off=r0 * F, n_out=up8(s.ik), aux=u,
                 ld_aux=F, aux_off=r0 * F, b_layout=ops.MNMAJOR, colsum=d_b1)
        a_du = acts.get(du, F, r0, rows, s.ik)
        a_xn = acts.get(xn, C, r0, rows, s.ek)
        # dW1[ik, ek] += du^T xn
        wgrads.append(((a_du, a_xn, F, C, s.ik, s.ek, rows, ops.EPI_ATOMIC, d_w1, C),
                       dict(a_off=r0 * F, b_off=r0 * C, a_layout=ops.MNMAJOR, b_layout=ops.MNMAJOR, split_k=split_k_for(s.ik, s.ek, rows))))
        ops.gemm_grouped(wgrads)
        # dxn[rows, ek] = du[rows, ik] W1[ik, ek]
        if meta.pre_norm:
            ops.gemm(a_du, w1, F, C, rows, s.ek, s.ik, ops.EPI_STORE, dxn, C, a_off=r0 * F, out_off=r0 * C, n_out=up8(s.ek),
                     b_layout=ops.MNMAJOR)
            ops.masked_ln_bwd(dxn, C, x2, C, mean, rstd, ln_w, g2 if meta.residual else None, gi2, C, d_lnw, d_lnb, rows, C, s.ek,
                              dy_off=r0 * C, x_off=r0 * C, stat_off=r0, g_off=r0 * C)
        else:
            ops.gemm(a_du, w1, F, C, rows, s.ek, s.ik, ops.EPI_STORE, gi2, C, a_off=r0 * F, out_off=r0 * C, n_out=C,
                     b_layout=ops.MNMAJOR)
    return g_in, (d_lnw, d_lnb, d_w1, d_b1, d_w2, d_b2)


# ------------------------------------------------------------------------------------------------ native fast path
# bf16 training path: one C-ABI call per half block and direction (csrc/half_block.cu sequences exactly the launches that the Python
# routines above issue).  The Python routines remain the implementation of the fp32 parity mode (split-bf16 operands) and of the
# instrumented runs (ops.PROFILE), and are the readable specification of the native one.
import ctypes as _C

from . import _lib

USE_NATIVE_HALF = True


def _segments_c(meta):
    arr = getattr(meta, '_cseg', None)
    if arr is None:
        arr = (_lib.Segment * len(meta.segs))(*[(s.b0, s.b1, s.ek, s.ik, s.ck, 1 if s.active else 0) for s in meta.segs])
        meta._cseg = arr
    return arr


def _half_desc(meta, x, out, params, ws16, ws32):
    """vsx_half_block for this call.  ws16: bf16 workspace [xn | act1 | act2]; ws32: fp32 [mean | rstd | lse]."""
    B, N, C = x.shape
    M = B * N
    ln_w, ln_b, w1, b1, w2, b2 = params
    attn = meta.kind == 'attn'
    inner = 3 * meta.H * meta.D if attn else meta.F
    a2 = meta.H * meta.D if attn else meta.F
    p16, p32 = ws16.data_ptr(), ws32.data_ptr()
    rs = meta.row_scale
    wa, wb = weights.get(w1), weights.get(w2)
    return _lib.HalfBlock(0 if attn else 1, B, N, C, meta.H, meta.D, meta.F, 1 if meta.pre_norm else 0, 1 if meta.residual else 0,
                          meta.eps, len(meta.segs), _C.cast(_segments_c(meta), _C.c_void_p),
                          x.data_ptr(), out.data_ptr() if out is not None else None, ln_w.data_ptr(), ln_b.data_ptr(), wa.data_ptr(), wb.data_ptr(),
                          None if b1 is None else b1.data_ptr(), None if b2 is None else b2.data_ptr(),
                          None if rs is None else rs.data_ptr(), meta.scale_off,
                          p16, p32, p32 + 4 * M, p16 + 2 * M * C, p16 + 2 * M * (C + inner), (p32 + 8 * M) if attn else None), a2


def native_half_forward(meta, x, params):
    B, N, C = x.shape
    M = B * N
    attn = meta.kind == 'attn'
    inner = 3 * meta.H * meta.D if attn else meta.F
    a2 = meta.H * meta.D if attn else meta.F
    dev = x.device
    out = torch.empty_like(x)
    ws16 = torch.empty(M * (C + inner + a2), device=dev, dtype=torch.bfloat16)
    ws32 = torch.empty(2 * M + (B * meta.H * N if attn else 0), device=dev, dtype=torch.float32)
    d, _ = _half_desc(meta, x, out, params, ws16, ws32)
    ops._ck(_lib.lib().vsx_half_block_fwd(_C.byref(d), ops._stream()))
    return out, (ws16, ws32)


FUSE_CAST = True      # LayerNorm backward of half block i also writes the scaled / masked bf16 gradient half block i-1 starts from


def _fusable(meta):
    """Single active segment covering the batch, pre-norm, residual: the shape of every half block of a 1-arch step."""
    if len(meta.segs) != 1:
        return False
    return bool(meta.segs[0].active) and meta.residual and meta.pre_norm


def _fusable_pair(consumer, producer):
    """The backward of `producer` (the later half block) may write the bf16 gradient copy `consumer` (the earlier one) starts from:
    both pre-norm + residual with identical sample ranges.  Segments that drop the layer take part (vsx_half_block_bwd: a dropped
    producer segment casts the passed-through gradient, a dropped consumer segment is left alone); with several segments the producer
    needs at least two active ones (its one-launch LayerNorm backward carries the per-segment extents of the cast)."""
    for m in (consumer, producer):
        if not (m.residual and m.pre_norm and 1 <= len(m.segs) <= 8 and all(s.b1 > s.b0 for s in m.segs)):
            return False
    if not any(s.active for s in consumer.segs):
        return False
    if len(producer.segs) == 1 and not (producer.segs[0].active and consumer.segs[0].active):
        return False
    return [(s.b0, s.b1) for s in consumer.segs] == [(s.b0, s.b1) for s in producer.segs]


def _half_backward_buffers(meta, x, params):
    B, N, C = x.shape
    M = B * N
    attn = meta.kind == 'attn'
    inner = 3 * meta.H * meta.D if attn else meta.F
    a2 = meta.H * meta.D if attn else 0
    sc = torch.empty(M * (2 * C + inner + a2), device=x.device, dtype=torch.bfloat16)      # [df | dxn | d_act1 | d_act2]
    return sc, zeros_like_many(*params)


def native_half_backward(meta, g_out, saved, x, params, ctx=None):
    ws16, ws32 = saved
    B, N, C = x.shape
    M = B * N
    attn = meta.kind == 'attn'
    inner = 3 * meta.H * meta.D if attn else meta.F
    g_in = torch.empty_like(x)
    pre = getattr(ctx, 'prefetched', None) if ctx is not None else None
    df_ready = 0
    if pre is not None:
        # buffers were allocated (in this order, so gradient-pool offsets are unchanged) by the half block that produced g_out
        sc, grads, g_ref, g_ver = pre
        ctx.prefetched = None
        if g_out.data_ptr() == g_ref.data_ptr() and g_out._version == g_ver and g_out.shape == g_ref.shape:
            df_ready = 1
        else:                       # the gradient was re-materialised or accumulated into: redo the cast from the tensor we were given
            grads[5].zero_()
    else:
        sc, grads = _half_backward_buffers(meta, x, params)
    fd, _ = _half_desc(meta, x, None, params, ws16, ws32)
    ps = sc.data_ptr()
    nxt = (None, None, 0, 0, None, None)
    prev = getattr(ctx, 'prev', None) if ctx is not None else None
    if FUSE_CAST and prev is not None and _fusable(meta) and _fusable(prev.meta) and prev.saved is not None:
        px, *pparams = prev.saved_tensors
        if px.shape == x.shape:
            psc, pgrads = _half_backward_buffers(prev.meta, px, pparams)
            pm = prev.meta
            rs = pm.row_scale
            nxt = (psc.data_ptr(), None if rs is None else rs.data_ptr(), pm.scale_off if rs is not None else 0, pm.segs[0].ck,
                   pgrads[5].data_ptr(), None)
            prev.prefetched = (psc, pgrads, g_in, None)
    d = _lib.HalfBlockGrad(fd, g_out.data_ptr(), g_in.data_ptr(), ps, ps + 2 * M * C, ps + 4 * M * C, (ps + 2 * M * (2 * C + inner)) if attn else None,
                           *[t.data_ptr() for t in grads], df_ready, *nxt)
    ops._ck(_lib.lib().vsx_half_block_bwd(_C.byref(d), ops._stream()))
    if nxt[0] is not None:
        psc, pgrads, g_ref, _ = prev.prefetched
        prev.prefetched = (psc, pgrads, g_ref, g_in._version)
    return g_in, tuple(grads)


_chain_tail = None     # (data_ptr of the last native half block's output, its ctx, shape, version)


def reset_half_chain():
    """Forget the previous half block: called at the start of every model forward so that a chain never spans two forwards (the caching
    allocator may hand the new step's first activation the address of the previous step's last one)."""
    global _chain_tail
    _chain_tail = None


class HalfBlockFn(torch.autograd.Function):
    """x_out = x + mask * drop_path(branch(LN(x)))  for one half of a Block (nets/supernet_blocks.py:213-253)."""

    @staticmethod
    def forward(ctx, meta, x, *params):
        require_cuda(x, 'Block')
        if not x.is_contiguous():
            x = x.contiguous()
        native = USE_NATIVE_HALF and _precision == 'bf16' and ops.PROFILE is None and all(p is not None for p in params)
        if native:
            out, saved = native_half_forward(meta, x, params)
        else:
            fwd = attn_half_forward if meta.kind == 'attn' else mlp_half_forward
            out, saved = fwd(meta, x, *params)
        ctx.meta, ctx.saved, ctx.native = meta, saved, native
        ctx.save_for_backward(x, *params)
        # chain of consecutive half blocks (this one consumes exactly what the previous one produced): lets the backward of this node
        # hand the previous node its scaled / masked gradient copy (FUSE_CAST)
        global _chain_tail
        ctx.prev = ctx.prefetched = None
        if native and _chain_tail is not None and _chain_tail[0] == x.data_ptr() and _chain_tail[2] == x.shape and _chain_tail[3] == x._version:
            ctx.prev = _chain_tail[1]
        _chain_tail = (out.data_ptr(), ctx, out.shape, out._version) if (native and any(ctx.needs_input_grad)) else None
        return out

    @staticmethod
    def backward(ctx, g):
        x, *params = ctx.saved_tensors
        g = g if g.is_contiguous() else g.contiguous()
        if ctx.native:
            g_in, pg = native_half_backward(ctx.meta, g, ctx.saved, x, params, ctx)
        else:
            bwd = attn_half_backward if ctx.meta.kind == 'attn' else mlp_half_backward
            g_in, pg = bwd(ctx.meta, g, ctx.saved, x, *params)
        ctx.saved = ctx.prev = None
        return (None, g_in) + tuple(pg)


# ------------------------------------------------------------------------------------------------ stage-level native path
# All half blocks of a stage (the transformer blocks between two spatial reductions) as ONE autograd node: one workspace allocation,
# one descriptor array and one C-ABI call per direction (vsx_stage_fwd / vsx_stage_bwd) instead of one of each per half block.
USE_NATIVE_STAGE = True


def stage_native_ok(metas, params):
    return (USE_NATIVE_STAGE and USE_NATIVE_HALF and _precision == 'bf16' and ops.PROFILE is None and len(metas) > 0
            and all(p is not None for p in params))


def _half_sizes(meta, M, C):
    attn = meta.kind == 'attn'
    inner = 3 * meta.H * meta.D if attn else meta.F
    a2 = meta.H * meta.D if attn else meta.F
    return attn, inner, a2


class StageFn(torch.autograd.Function):
    """x -> half block 0 -> ... -> half block n-1 (each x_out = x + mask * drop_path(branch(LN(x))), nets/supernet_blocks.py:213-253)."""

    @staticmethod
    def forward(ctx, metas, x, *params):
        require_cuda(x, 'Block')
        if not x.is_contiguous():
            x = x.contiguous()
        n = len(metas)
        B, N, C = x.shape
        M = B * N
        dev = x.device
        outs = torch.empty((n, B, N, C), device=dev, dtype=torch.float32)
        n16 = n32 = 0
        sizes = []
        for meta in metas:
            attn, inner, a2 = _half_sizes(meta, M, C)
            sizes.append((attn, inner, a2, n16, n32))
            n16 += M * (C + inner + a2)
            n32 += 2 * M + (B * meta.H * N if attn else 0)
        ws16 = torch.empty(n16, device=dev, dtype=torch.bfloat16)
        ws32 = torch.empty(n32, device=dev, dtype=torch.float32)
        p16, p32, po, px = ws16.data_ptr(), ws32.data_ptr(), outs.data_ptr(), x.data_ptr()
        ostride = 4 * M * C
        descs = (_lib.HalfBlock * n)()
        keep_alive = []
        for i, meta in enumerate(metas):
            attn, inner, a2, o16, o32 = sizes[i]
            ln_w, ln_b, w1, b1, w2, b2 = params[6 * i:6 * i + 6]
            wa, wb = weights.get(w1), weights.get(w2)
            keep_alive.append((wa, wb))
            rs = meta.row_scale
            q16, q32 = p16 + 2 * o16, p32 + 4 * o32
            descs[i] = _lib.HalfBlock(0 if attn else 1, B, N, C, meta.H, meta.D, meta.F, 1 if meta.pre_norm else 0, 1 if meta.residual else 0,
                                      meta.eps, len(meta.segs), _C.cast(_segments_c(meta), _C.c_void_p),
                                      px if i == 0 else po + (i - 1) * ostride, po + i * ostride, ln_w.data_ptr(), ln_b.data_ptr(), wa.data_ptr(), wb.data_ptr(),
                                      None if b1 is None else b1.data_ptr(), None if b2 is None else b2.data_ptr(),
                                      None if rs is None else rs.data_ptr(), meta.scale_off,
                                      q16, q32, q32 + 4 * M, q16 + 2 * M * C, q16 + 2 * M * (C + inner), (q32 + 8 * M) if attn else None)
        ops._ck(_lib.lib().vsx_stage_fwd(descs, n, ops._stream()))
        ctx.metas, ctx.descs, ctx.sizes = metas, descs, sizes
        ctx.keep = (outs, ws16, ws32, keep_alive)
        ctx.save_for_backward(x, *params)
        return outs[n - 1]

    @staticmethod
    def backward(ctx, g):
        x, *params = ctx.saved_tensors
        metas, fdescs, sizes = ctx.metas, ctx.descs, ctx.sizes
        n = len(metas)
        B, N, C = x.shape
        M = B * N
        dev = x.device
        g = g if g.is_contiguous() else g.contiguous()
        gbuf = torch.empty((2, B, N, C), device=dev, dtype=torch.float32)       # g_in of half block i = g_out of half block i - 1: ping-pong
        sc_elems = max(M * (2 * C + inner + (a2 if attn else 0)) for attn, inner, a2, _, _ in sizes)
        sc = torch.empty(2 * sc_elems, device=dev, dtype=torch.bfloat16)        # [df | dxn | d_act1 | d_act2] of half block i in buffer i & 1
        grads = zeros_like_many(*params)
        pg, pgb, psc = g.data_ptr(), gbuf.data_ptr(), sc.data_ptr()
        gstride = 4 * M * C
        descs = (_lib.HalfBlockGrad * n)()
        fuse = [False] * n          # fuse[i]: half block i's df is written by the LayerNorm backward of half block i + 1
        if FUSE_CAST:
            for i in range(n - 1):
                fuse[i] = _fusable_pair(metas[i], metas[i + 1])
        for i, meta in enumerate(metas):
            attn, inner, a2, _, _ = sizes[i]
            ps = psc + 2 * sc_elems * (i & 1)
            g_out = pg if i == n - 1 else pgb + gstride * ((i + 1) & 1)
            g_in = pgb + gstride * (i & 1)
            nxt = (None, None, 0, 0, None, None)
            if i > 0 and fuse[i - 1]:
                pm = metas[i - 1]
                rs = pm.row_scale
                nxt = (psc + 2 * sc_elems * ((i - 1) & 1), None if rs is None else rs.data_ptr(), pm.scale_off if rs is not None else 0, pm.segs[0].ck,
                       grads[6 * (i - 1) + 5].data_ptr(), _C.cast(_segments_c(pm), _C.c_void_p))
            gi = grads[6 * i:6 * i + 6]
            descs[i] = _lib.HalfBlockGrad(fdescs[i], g_out, g_in, ps, ps + 2 * M * C, ps + 4 * M * C, (ps + 2 * M * (2 * C + inner)) if attn else None,
                                          *[t.data_ptr() for t in gi], 1 if fuse[i] else 0, *nxt)
        ops._ck(_lib.lib().vsx_stage_bwd(descs, n, ops._stream()))
        ctx.keep = None
        if pool_prefix_ready_hook is not None:
            pool_prefix_ready_hook()        # every gradient carved from the pool so far is final: the data-parallel exchange of that prefix may start
        return (None, gbuf[0]) + tuple(grads)
