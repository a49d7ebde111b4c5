"""ctypes binding of libvsx.so (the C ABI declared in include/vsx.h).

The library is the product: there is NO fallback.  If it is missing, stale against the header, or the device is
not an sm_100 part, loading / calling fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libvsx.so')
ABI_VERSION = 4

BF16, F32 = 0, 1
KMAJOR, MNMAJOR = 0, 1
EPI_STORE, EPI_GELU, EPI_RESIDUAL, EPI_GELUGRAD, EPI_ATOMIC = 0, 1, 2, 3, 4

_p, _i, _l, _f = C.c_void_p, C.c_int, C.c_long, C.c_float


class GemmDesc(C.Structure):
    _fields_ = [('a', _p * 6), ('b', _p * 6), ('terms', _i), ('lda', _l), ('ldb', _l), ('a_layout', _i), ('b_layout', _i),
                ('M', _i), ('N', _i), ('K', _i), ('epilogue', _i), ('out_dtype', _i), ('out', _p), ('ldo', _l),
                ('out2', _p), ('ldo2', _l), ('n_out', _i), ('bias', _p), ('aux', _p), ('ld_aux', _l),
                ('row_scale', _p), ('rows_per_sample', _i), ('n_keep', _i), ('split_k', _i), ('colsum', _p),
                ('k_segments', _i), ('k_seg_len', _i), ('k_seg_stride', _i)]


# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/vsx.h one to one
SIGNATURES = {
    'vsx_last_error': [],
    'vsx_abi_version': [],
    'vsx_device_ok': [_i],
    'vsx_masked_ln_fwd': [_p, _l, _p, _p, _p, _p, _i, _l, _p, _p, _i, _i, _i, _f, _i, _i, _p],
    'vsx_masked_ln_bwd_cast': [_p, _i, _l, _p, _l, _p, _p, _p, _p, _p, _l, _p, _p, _i, _i, _i, _p, _l, _p, _i, _i, _p, _p],
    'vsx_masked_ln_bwd': [_p, _p, _i, _l, _p, _l, _p, _p, _p, _p, _p, _l, _p, _p, _i, _i, _i, _i, _i, _p],
    'vsx_masked_ln_fwd_segs': [_p, _l, _p, _p, _p, _i, _l, _p, _p, _i, _i, _p, _f, _p],
    'vsx_masked_ln_bwd_segs': [_p, _i, _l, _p, _l, _p, _p, _p, _p, _p, _l, _p, _p, _i, _i, _p, _p, _l, _p, _i, _p, _p],
    'vsx_scale_mask_cast_segs': [_p, _l, _p, _i, _p, _i, _l, _i, _i, _p, _p, _p],
    'vsx_gemm': [C.POINTER(GemmDesc), _p],
    'vsx_attn_fwd': [_p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _p],
    'vsx_attn_bwd': [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _p, _p],
    'vsx_attn_fwd_segs': [_p, _p, _p, _i, _i, _i, _i, _i, _p, _f, _i, _p],
    'vsx_attn_bwd_segs': [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _f, _i, _p, _p],
    'vsx_attn_debug_buffer': [_p],
    'vsx_attn_odd_token_modes': [_i, _i, _i],
    'vsx_gemm_force_tile_rows': [_i],
    'vsx_gemm_force_cta_group': [_i],
    'vsx_gemm_grouped': [_p, _i, _p],
    'vsx_half_block_fwd': [_p, _p],
    'vsx_half_block_bwd': [_p, _p],
    'vsx_stage_fwd': [_p, _i, _p],
    'vsx_stage_bwd': [_p, _i, _p],
    'vsx_launch_count': [],
    'vsx_token_mix': [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, C.c_double, C.c_double, _p],
    'vsx_conv3x3_force_impl': [_i],
    'vsx_conv1_fwd': [_p, _p, _l, _p, _i, _i, _i, _p, _p],
    'vsx_conv1_wgrad': [_p, _p, _p, _l, _i, _i, _i, _p],
    'vsx_eval_metrics': [_p, _l, _p, _i, _i, _p, _p, _p, _p],
    'vsx_gemm_debug_buffer': [_p],
    'vsx_split_bf16': [_p, _l, _p, _p, _p, _l, _i, _i, _p],
    'vsx_scale_mask_cast': [_p, _l, _p, _i, _i, _p, _i, _l, _i, _i, _p, _p],
    'vsx_colsum': [_p, _i, _l, _i, _i, _p, _p],
    'vsx_image_normalize_u8': [_p, _p, _i, _i, _l, _p, _p, _p],
    'vsx_im2col': [_p, _p, _p, _p, _p, _p, _i, _i, _l, _l, _i, _i, _i, _i, _i, _i, _i, _p, _i, _l, _p],
    'vsx_col2im': [_p, _l, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _l, _l, _p],
    'vsx_bn_stats': [_p, _i, _l, _i, _p, _p],
    'vsx_bn_finalize': [_p, _l, _i, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p],
    'vsx_bn_bwd_stats': [_p, _p, _i, _l, _i, _p, _p, _p, _p, _p, _p],
    'vsx_bn_bwd_apply': [_p, _p, _i, _l, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    'vsx_conv3x3': [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    'vsx_conv3x3_wgrad': [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    'vsx_embed_assemble': [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    'vsx_embed_assemble_bwd': [_p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _p],
    'vsx_sr_combine': [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    'vsx_sr_combine_bwd': [_p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _p],
    'vsx_soft_ce': [_p, _l, _p, _l, _i, _i, _f, _f, _p, _p, _l, _p],
    'vsx_scale_by_scalar': [_p, _l, _p, _p],
    'vsx_adamw_chunk_elems': [],
    'vsx_adamw': [_p, _p, _p, _i, C.c_double, C.c_double, C.c_double, C.c_double, _i, _p, _p, _p, _p],
}
_RESTYPES = {'vsx_last_error': C.c_char_p, 'vsx_launch_count': C.c_long}



class AdamWTensor(C.Structure):
    _fields_ = [('param', _p), ('grad', _p), ('exp_avg', _p), ('exp_avg_sq', _p), ('shadow_hi', _p), ('shadow_lo', _p),
                ('numel', _l), ('weight_decay', _f), ('ema_decay', _f), ('ema', _p)]


class RowSegments(C.Structure):
    _fields_ = [('count', _i), ('row_end', _i * 8), ('keep', _i * 8), ('keep2', _i * 8)]


class SampleSegments(C.Structure):
    _fields_ = [('count', _i), ('sample_end', _i * 8), ('heads_keep', _i * 8)]


class Segment(C.Structure):
    _fields_ = [('b0', _i), ('b1', _i), ('embed_keep', _i), ('inner_keep', _i), ('out_keep', _i), ('active', _i)]


class HalfBlock(C.Structure):
    _fields_ = [('kind', _i), ('batch', _i), ('tokens', _i), ('width', _i), ('heads', _i), ('head_dim', _i), ('hidden', _i),
                ('pre_norm', _i), ('residual', _i), ('eps', _f), ('num_segments', _i), ('segments', _p),
                ('x', _p), ('out', _p), ('ln_w', _p), ('ln_b', _p), ('w1', _p), ('w2', _p), ('b1', _p), ('b2', _p),
                ('row_scale', _p), ('scale_off', _i), ('xn', _p), ('mean', _p), ('rstd', _p), ('act1', _p), ('act2', _p), ('lse', _p)]


class HalfBlockGrad(C.Structure):
    _fields_ = [('fwd', HalfBlock), ('g_out', _p), ('g_in', _p), ('df', _p), ('dxn', _p), ('d_act1', _p), ('d_act2', _p),
                ('d_ln_w', _p), ('d_ln_b', _p), ('d_w1', _p), ('d_b1', _p), ('d_w2', _p), ('d_b2', _p),
                ('df_ready', _i), ('next_df', _p), ('next_row_scale', _p), ('next_scale_off', _i), ('next_keep', _i), ('next_d_b2', _p), ('next_segments', _p)]


_lib = None


class VsxError(RuntimeError):
    pass


def lib():
    """Load libvsx.so once.  Raises if it has not been built (python -m vit_search_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VsxError('libvsx.so not found at %s -- build it with `python -m vit_search_b200.build`; '
                           'there is no CPU or PyTorch fallback for the hot path' % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(h, name)          # AttributeError if the library lacks a declared symbol
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _i)
        if h.vsx_abi_version() != ABI_VERSION:
            raise VsxError('libvsx.so ABI version %d does not match the Python binding (%d); rebuild'
                           % (h.vsx_abi_version(), ABI_VERSION))
        _lib = h
    return _lib


def check(rc):
    if rc != 0:
        raise VsxError('libvsx error %d: %s' % (rc, lib().vsx_last_error().decode()))
