"""Tensor-level wrappers over the C ABI (include/vsx.h).  PyTorch is used for device memory and streams only.

Every function takes CUDA tensors, computes raw pointers (+ element offsets for segment / column windows) and
calls into libvsx.so on the current stream.  No arithmetic happens here.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import BF16, F32, KMAJOR, MNMAJOR, EPI_STORE, EPI_GELU, EPI_RESIDUAL, EPI_GELUGRAD, EPI_ATOMIC  # noqa: F401

_DT = {torch.bfloat16: BF16, torch.float32: F32}

LAUNCHES = 0      # kernels launched through the C ABI (every entry point launches exactly one kernel)
PROFILE = None    # when a list: (start_event, end_event, algorithmic_flops) per tensor-core GEMM launch (bench.py roofline)


def _ck(rc):
    global LAUNCHES
    LAUNCHES += 1
    _lib.check(rc)


_raw_stream = torch._C._cuda_getCurrentRawStream if hasattr(torch._C, '_cuda_getCurrentRawStream') else None
_get_device = torch._C._cuda_getDevice if hasattr(torch._C, '_cuda_getDevice') else None


def _stream():
    """cudaStream_t of torch's current stream on torch's CURRENT device (fast path: two C-level queries, ~0.4 us; current_stream()
    costs ~14 us).  The device is re-read on every call, so a process that calls torch.cuda.set_device(local_rank) late, or switches
    devices, never enqueues work on another GPU's stream."""
    if _raw_stream is None or _get_device is None:
        return torch.cuda.current_stream().cuda_stream
    return _raw_stream(_get_device())


def set_device(index):
    """Make `index` the current CUDA device of this process (one process per GPU).  Kept for callers of the round-1 API; the
    bindings follow torch's current device by themselves."""
    torch.cuda.set_device(index)


_ESIZE = {torch.bfloat16: 2, torch.float32: 4, torch.float64: 8, torch.int32: 4, torch.int64: 8, torch.uint8: 1, torch.bool: 1}


def _ptr(t, offset=0):
    """Device pointer of element `offset` of tensor t (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('libvsx operates on CUDA tensors only (no CPU fallback)')
    return t.data_ptr() + offset * _ESIZE[t.dtype]


def dt(t):
    return _DT[t.dtype]


# ------------------------------------------------------------------------------------------------ layer norm
def masked_ln_fwd(x, ldx, gamma, beta, y, ldy, mean, rstd, rows, C_, keep, eps, x_off=0, y_off=0, stat_off=0,
                  y2=None, rows_per_sample=0, split_tokens=0):
    _ck(_lib.lib().vsx_masked_ln_fwd(
        _ptr(x, x_off), ldx, _ptr(gamma), _ptr(beta), _ptr(y, y_off), _ptr(y2), dt(y), ldy,
        _ptr(mean, stat_off), _ptr(rstd, stat_off), rows, C_, keep, eps, rows_per_sample, split_tokens, _stream()))


def masked_ln_bwd(dy, lddy, x, ldx, mean, rstd, gamma, g_in, g_out, ldg, dgamma, dbeta, rows, C_, keep,
                  dy_off=0, x_off=0, stat_off=0, g_off=0, dy2=None, rows_per_sample=0, split_tokens=0):
    _ck(_lib.lib().vsx_masked_ln_bwd(
        _ptr(dy, dy_off), _ptr(dy2), dt(dy), lddy, _ptr(x, x_off), ldx, _ptr(mean, stat_off), _ptr(rstd, stat_off),
        _ptr(gamma), _ptr(g_in, g_off), _ptr(g_out, g_off), ldg, _ptr(dgamma), _ptr(dbeta), rows, C_, keep,
        rows_per_sample, split_tokens, _stream()))


# ------------------------------------------------------------------------------------------------ GEMM
_P6 = C.c_void_p * 6
_PAIRS3 = [(0, 0), (1, 0), (0, 1)]
_PAIRS6 = [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (0, 2)]

def _gemm_desc(a, b, lda, ldb, M, N, K, epilogue, out, ldo, *, a_off=0, b_off=0, out_off=0, a_layout=KMAJOR,
               b_layout=KMAJOR, n_out=None, out2=None, ldo2=0, out2_off=0, bias=None, bias_off=0, aux=None, ld_aux=0,
               aux_off=0, row_scale=None, row_scale_off=0, rows_per_sample=1, n_keep=0, split_k=1, colsum=None, colsum_off=0,
               k_segments=0, k_seg_len=0, k_seg_stride=0):
    """-> (vsx_gemm_desc, profile record).  a, b: a bf16 tensor, or tuples of 2 (3 product terms, ~2^-16) or 3 (6 terms, fp32-exact)
    bf16 parts whose sum is the fp32 operand (split-bf16 high-precision mode)."""
    if isinstance(a, (tuple, list)):
        n = min(len(a), len(b))
        pairs = _PAIRS3 if n == 2 else _PAIRS6
        pa = _P6(*[_ptr(a[i], a_off) for i, _ in pairs])
        pb = _P6(*[_ptr(b[j], b_off) for _, j in pairs])
        nterms = len(pairs)
    else:
        pa, pb, nterms = _P6(_ptr(a, a_off)), _P6(_ptr(b, b_off)), 1
    # positional construction: one C-level initialisation instead of ~25 Python-level field stores
    d = _lib.GemmDesc(pa, pb, nterms, lda, ldb, a_layout, b_layout, M, N, K, epilogue, _DT[out.dtype], _ptr(out, out_off), ldo,
                      _ptr(out2, out2_off), ldo2, N if n_out is None else n_out, _ptr(bias, bias_off), _ptr(aux, aux_off), ld_aux,
                      _ptr(row_scale, row_scale_off), rows_per_sample, n_keep, split_k, _ptr(colsum, colsum_off), k_segments, k_seg_len,
                      k_seg_stride)
    if PROFILE is None:
        return d, None
    # algorithmic HBM bytes of the launch: both operands once + the output tile(s) + the aux tile the epilogue reads
    es = 4 if out.dtype == torch.float32 else 2
    nbytes = 2.0 * nterms * (M * K + N * K) + es * M * d.n_out * (2 if epilogue == EPI_GELU else 1)
    if epilogue in (EPI_RESIDUAL, EPI_GELUGRAD):
        nbytes += es * M * d.n_out
    if epilogue == EPI_ATOMIC:
        nbytes += es * M * d.n_out          # read-modify-write of the fp32 gradient tile
    keff = k_segments * k_seg_len if k_segments > 1 else K        # algorithmic reduction length
    return d, (2.0 * M * N * keff, (M, N, keff, epilogue, a_layout, b_layout, d.n_out, split_k, nterms), nbytes)


def gemm(*args, **kw):
    """One GEMM launch (see _gemm_desc for the arguments)."""
    d, rec = _gemm_desc(*args, **kw)
    if rec is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _ck(_lib.lib().vsx_gemm(C.byref(d), _stream()))
        e1.record()
        PROFILE.append((e0, e1, rec[0], rec[1], rec[2]))
        return
    _ck(_lib.lib().vsx_gemm(C.byref(d), _stream()))


def gemm_grouped(problems):
    """Up to 4 problems (each a (args, kwargs) pair for _gemm_desc) with the same epilogue / output dtype in ONE launch."""
    if len(problems) == 1:
        return gemm(*problems[0][0], **problems[0][1])
    built = [_gemm_desc(*a, **k) for a, k in problems]
    arr = (_lib.GemmDesc * len(built))(*[b[0] for b in built])
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _ck(_lib.lib().vsx_gemm_grouped(arr, len(built), _stream()))
        e1.record()
        PROFILE.append((e0, e1, sum(b[1][0] for b in built), (sum(b[1][1][0] for b in built),) + built[0][1][1][1:], sum(b[1][2] for b in built)))
        return
    _ck(_lib.lib().vsx_gemm_grouped(arr, len(built), _stream()))


# ------------------------------------------------------------------------------------------------ attention core
ATTN_AUTO, ATTN_FP32, ATTN_MMA_SYNC, ATTN_TCGEN05 = 0, 1, 2, 3


def attn_fwd(qkv, o, lse, batch, tokens, heads, head_dim, heads_keep, scale, *, qkv_off=0, o_off=0, lse_off=0, impl=ATTN_AUTO):
    _ck(_lib.lib().vsx_attn_fwd(_ptr(qkv, qkv_off), _ptr(o, o_off), _ptr(lse, lse_off), dt(qkv), batch, tokens, heads,
                                       head_dim, heads_keep, scale, impl, _stream()))


def attn_bwd(qkv, o, d_o, lse, dqkv, batch, tokens, heads, head_dim, heads_keep, scale, *, qkv_off=0, o_off=0, lse_off=0,
             impl=ATTN_AUTO, dbias=None):
    _ck(_lib.lib().vsx_attn_bwd(_ptr(qkv, qkv_off), _ptr(o, o_off), _ptr(d_o, o_off), _ptr(lse, lse_off),
                                       _ptr(dqkv, qkv_off), dt(qkv), batch, tokens, heads, head_dim, heads_keep, scale, impl,
                                       _ptr(dbias), _stream()))


# ------------------------------------------------------------------------------------------------ elementwise
def split_bf16(src, lds, hi, lo, ldd, rows, cols, src_off=0, dst_off=0, lo2=None):
    _ck(_lib.lib().vsx_split_bf16(_ptr(src, src_off), lds, _ptr(hi, dst_off), _ptr(lo, dst_off), _ptr(lo2, dst_off), ldd, rows, cols,
                                  _stream()))


def scale_mask_cast(g, ldg, row_scale, rows_per_sample, n_keep, out, ldo, rows, cols, g_off=0, out_off=0, scale_off=0, colsum=None):
    _ck(_lib.lib().vsx_scale_mask_cast(_ptr(g, g_off), ldg, _ptr(row_scale, scale_off), rows_per_sample, n_keep,
                                              _ptr(out, out_off), dt(out), ldo, rows, cols, _ptr(colsum), _stream()))


def colsum(x, ldx, rows, cols, out, x_off=0, out_off=0):
    _ck(_lib.lib().vsx_colsum(_ptr(x, x_off), dt(x), ldx, rows, cols, _ptr(out, out_off), _stream()))


# ------------------------------------------------------------------------------------------------ generic caller
def call(name, *args):
    """Call `vsx_<name>` with tensors converted to device pointers ((tensor, element_offset) tuples allowed, None -> NULL);
    the current CUDA stream is appended as the last argument."""
    conv = []
    for a in args:
        if isinstance(a, torch.Tensor):
            conv.append(_ptr(a))
        elif isinstance(a, tuple):
            conv.append(_ptr(a[0], a[1]))
        else:
            conv.append(a)
    _ck(getattr(_lib.lib(), 'vsx_' + name)(*conv, _stream()))
