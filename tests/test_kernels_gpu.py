"""GPU parity of the individual kernels (through the C ABI) against the CPU oracle / fp64 torch math."""
import numpy as np
import pytest
import torch

from oracle import vit_res_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope='module')
def ops():
    from vit_search_b200 import _lib, ops
    _lib.check(_lib.lib().vsx_device_ok(0))
    return ops


# ---------------------------------------------------------------------------------------------- masked LN
@pytest.fixture(params=[128, 256, 'pair'])
def tile_rows(request, ops):
    """Run every GEMM test with all CTA tile shapes of csrc/gemm_tc.cu: 128 x 128, 256 x 128 (two accumulators sharing one B box per k
    block) and 256 x 256 on a CTA PAIR (tcgen05.mma.cta_group::2, a cluster of two SMs), including tiles whose second half lies beyond
    M or N."""
    from vit_search_b200 import _lib
    if request.param == 'pair':
        _lib.check(_lib.lib().vsx_gemm_force_cta_group(2))
    else:
        _lib.check(_lib.lib().vsx_gemm_force_cta_group(1))
        _lib.check(_lib.lib().vsx_gemm_force_tile_rows(request.param))
    yield request.param
    _lib.lib().vsx_gemm_force_tile_rows(0)
    _lib.lib().vsx_gemm_force_cta_group(0)


@pytest.mark.parametrize('C,keep', [(64, 64), (64, 44), (256, 160), (320, 220), (1024, 704), (1280, 1280)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_masked_ln(ops, C, keep, dtype):
    g = torch.Generator().manual_seed(C + keep)
    B, N = 3, 37
    rows = B * N
    x = torch.randn(B, N, C, generator=g) * O.prefix_mask([keep] * B, C, torch.float32) + 0.3
    x = x * O.prefix_mask([keep] * B, C, torch.float32)
    w = 1 + 0.1 * torch.randn(C, generator=g)
    b = 0.1 * torch.randn(C, generator=g)
    go = torch.randn(B, N, C, generator=g)
    gin = torch.randn(B, N, C, generator=g)
    y_ref = O.masked_layer_norm(x.double(), w.double(), b.double(), [keep] * B)
    m = O.prefix_mask([keep] * B, C, torch.float64)
    go_q = go.to(dtype).double()
    gx_ref, gw_ref, gb_ref = O.masked_layer_norm_backward(go_q * m, x.double(), w.double(), [keep] * B)
    gx_ref = gx_ref * m + gin.double()        # masked channels carry only the incoming gradient

    xd, wd, bd = x.cuda().view(rows, C), w.cuda(), b.cuda()
    y = torch.full((rows, C), float('nan'), device='cuda', dtype=dtype)
    mean = torch.empty(rows, device='cuda')
    rstd = torch.empty(rows, device='cuda')
    ops.masked_ln_fwd(xd, C, wd, bd, y, C, mean, rstd, rows, C, keep, 1e-6)
    tol = 1e-5 if dtype == torch.float32 else 6e-3
    assert rel(y.view(B, N, C), y_ref) < tol
    assert torch.all(y[:, keep:] == 0)

    gout = torch.full((rows, C), float('nan'), device='cuda')
    dgam = torch.zeros(C, device='cuda')
    dbet = torch.zeros(C, device='cuda')
    ops.masked_ln_bwd(go.to(dtype).cuda().view(rows, C), C, xd, C, mean, rstd, wd, gin.cuda().view(rows, C), gout, C,
                      dgam, dbet, rows, C, keep)
    assert rel(gout.view(B, N, C), gx_ref) < 2e-5
    assert rel(dgam, gw_ref) < 2e-5
    assert rel(dbet, gb_ref) < 2e-5


def test_masked_ln_golden(ops):
    """The reference's own outputs (tests/golden/functions.npz) with per-sample keeps -> one segment per sample."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'functions.npz'))
    keep = G['ln_keep'].tolist()
    B, N, C = 6, 5, 48
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, N, C, generator=g) * O.prefix_mask(keep, C, torch.float32)
    wt = 1 + 0.1 * torch.randn(C, generator=g)
    bs = 0.1 * torch.randn(C, generator=g)
    go = torch.randn(B, N, C, generator=g)
    xd = x.cuda().view(B * N, C)
    y = torch.empty(B * N, C, device='cuda')
    mean = torch.empty(B * N, device='cuda')
    rstd = torch.empty(B * N, device='cuda')
    gout = torch.empty(B * N, C, device='cuda')
    dg = torch.zeros(C, device='cuda')
    db = torch.zeros(C, device='cuda')
    god = go.cuda().view(B * N, C)
    for s in range(B):
        ops.masked_ln_fwd(xd, C, wt.cuda(), bs.cuda(), y, C, mean, rstd, N, C, keep[s], 1e-6,
                          x_off=s * N * C, y_off=s * N * C, stat_off=s * N)
        ops.masked_ln_bwd(god, C, xd, C, mean, rstd, wt.cuda(), None, gout, C, dg, db, N, C, keep[s],
                          dy_off=s * N * C, x_off=s * N * C, stat_off=s * N, g_off=s * N * C)
    m = O.prefix_mask(keep, C, torch.float32)
    assert rel(y.view(B, N, C), torch.from_numpy(G['ln_y'])) < 1e-5
    assert rel(gout.view(B, N, C).cpu() * m, torch.from_numpy(G['ln_gx']) * m) < 2e-5   # masked lanes are dead gradients
    assert rel(dg, torch.from_numpy(G['ln_gw'])) < 2e-5
    assert rel(db, torch.from_numpy(G['ln_gb'])) < 2e-5


# ---------------------------------------------------------------------------------------------- GEMM
def _split(t):
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi, lo


def _mk(M, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (256, 128, 128), (300, 200, 96), (1028, 576, 256), (514, 768, 160),
                                   (130, 1000, 1024), (64, 44, 220)])
def test_gemm_store_kmajor(ops, M, N, K, tile_rows):
    Kp = (K + 7) // 8 * 8 + 8            # pitch > K: junk beyond K must be clipped by the TMA descriptor
    A, W = _mk(M, N, Kp, M + N + K)
    bias = torch.randn(N)
    Ab, Wb = A.to(torch.bfloat16), W.to(torch.bfloat16)
    ref = Ab[:, :K].double() @ Wb[:, :K].double().t() + bias.double()
    n_out = (N + 7) // 8 * 8
    for odt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 5e-3)):
        out = torch.full((M, n_out), float('nan'), device='cuda', dtype=odt)
        ops.gemm(Ab.cuda(), Wb.cuda(), Kp, Kp, M, N, K, ops.EPI_STORE, out, n_out, n_out=n_out, bias=bias.cuda())
        torch.cuda.synchronize()
        assert rel(out[:, :N], ref) < tol, (M, N, K, odt)
        assert torch.all(out[:, N:] == 0)


def test_gemm_split3_precision(ops, tile_rows):
    """bf16x3 split mode must recover ~fp32 accuracy (the high-precision parity path)."""
    M, N, K = 384, 256, 512
    A, W = _mk(M, N, K, 5)
    ref = A.double() @ W.double().t()
    (ah, al), (wh, wl) = _split(A), _split(W)
    out = torch.empty(M, N, device='cuda')
    ops.gemm((ah.cuda(), al.cuda()), (wh.cuda(), wl.cuda()), K, K, M, N, K, ops.EPI_STORE, out, N)
    assert rel(out, ref) < 3e-5
    out1 = torch.empty(M, N, device='cuda')
    ops.gemm(ah.cuda(), wh.cuda(), K, K, M, N, K, ops.EPI_STORE, out1, N)
    assert 1e-4 < rel(out1, ref) < 1e-2      # plain bf16 is visibly worse: the 3 terms really are accumulated
    # three bf16 parts per operand, six product terms: fp32-exact (the 'fp32' parity mode of core.py)
    a3 = (ah, al, (A - ah.float() - al.float()).to(torch.bfloat16))
    w3 = (wh, wl, (W - wh.float() - wl.float()).to(torch.bfloat16))
    out6 = torch.empty(M, N, device='cuda')
    ops.gemm(tuple(t.cuda() for t in a3), tuple(t.cuda() for t in w3), K, K, M, N, K, ops.EPI_STORE, out6, N)
    e6 = rel(out6, ref)
    assert e6 < 1e-5      # floor (~5e-6 at K=512) = fp32 accumulation inside the tensor core, not the operand split
    # and the device-side splitter agrees with the host split
    src = A.cuda()
    parts = [torch.empty(M, K, device='cuda', dtype=torch.bfloat16) for _ in range(3)]
    ops.split_bf16(src, K, parts[0], parts[1], K, M, K, lo2=parts[2])
    assert all(torch.equal(p.cpu(), q) for p, q in zip(parts, a3))


@pytest.mark.parametrize('M,N,K', [(256, 128, 128), (300, 136, 200), (1028, 256, 576)])
def test_gemm_dgrad_layout(ops, M, N, K, tile_rows):
    """dX[M,N] = dY[M,K] @ W[K,N]: A K-major, B MN-major (W stored [K rows, N contiguous])."""
    g = torch.Generator().manual_seed(M)
    dY = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(K, N, generator=g) * 0.05).to(torch.bfloat16)
    ref = dY.double() @ W.double()
    out = torch.empty(M, N, device='cuda')
    ops.gemm(dY.cuda(), W.cuda(), K, N, M, N, K, ops.EPI_STORE, out, N, b_layout=ops.MNMAJOR)
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize('R,Nw,Kw,split', [(256, 128, 128, 1), (1000, 200, 136, 3), (4112, 576, 256, 8), (771, 44, 60, 2)])
def test_gemm_wgrad_layout(ops, R, Nw, Kw, split, tile_rows):
    """dW[Nw,Kw] += dY[R,Nw]^T @ X[R,Kw]: both operands MN-major, split-K atomics."""
    g = torch.Generator().manual_seed(R)
    lda, ldb = (Nw + 7) // 8 * 8 + 8, (Kw + 7) // 8 * 8 + 16     # pitch > extent: exercises TMA OOB clipping
    dY = torch.randn(R, lda, generator=g).to(torch.bfloat16)
    X = torch.randn(R, ldb, generator=g).to(torch.bfloat16)
    ref = dY[:, :Nw].double().t() @ X[:, :Kw].double()
    out = torch.ones(Nw, Kw, device='cuda')
    ops.gemm(dY.cuda(), X.cuda(), lda, ldb, Nw, Kw, R, ops.EPI_ATOMIC, out, Kw, a_layout=ops.MNMAJOR,
             b_layout=ops.MNMAJOR, split_k=split)
    assert rel(out - 1, ref) < 2e-5


def test_gemm_epilogues(ops, tile_rows):
    M, N, K, C = 514, 200, 128, 256      # 2 samples x 257 rows
    A, W = _mk(M, N, K, 9)
    Ab, Wb = A.to(torch.bfloat16), W.to(torch.bfloat16)
    bias = torch.randn(N) * 0.1
    acc = Ab.double() @ Wb.double().t() + bias.double()
    # GELU: out = gelu'(pre-activation), out2 = gelu(pre-activation), zero fill up to n_out (the pre-activation itself is not stored)
    pre = torch.full((M, 256), float('nan'), device='cuda', dtype=torch.bfloat16)
    act = torch.full((M, 256), float('nan'), device='cuda', dtype=torch.bfloat16)
    ops.gemm(Ab.cuda(), Wb.cuda(), K, K, M, N, K, ops.EPI_GELU, pre, 256, n_out=256, out2=act, ldo2=256, bias=bias.cuda())
    accg = acc.clone().requires_grad_(True)
    torch.nn.functional.gelu(accg).sum().backward()
    assert rel(pre[:, :N], accg.grad) < 5e-3 and rel(act[:, :N], torch.nn.functional.gelu(acc)) < 6e-3
    assert torch.all(pre[:, N:] == 0) and torch.all(act[:, N:] == 0)
    # the same in the fp32 parity mode (exact erf)
    pre32 = torch.full((M, 256), float('nan'), device='cuda')
    act32 = torch.full((M, 256), float('nan'), device='cuda')
    ops.gemm(Ab.cuda(), Wb.cuda(), K, K, M, N, K, ops.EPI_GELU, pre32, 256, n_out=256, out2=act32, ldo2=256, bias=bias.cuda())
    assert rel(pre32[:, :N], accg.grad) < 1e-5 and rel(act32[:, :N], torch.nn.functional.gelu(acc)) < 1e-5
    assert torch.all(pre32[:, N:] == 0) and torch.all(act32[:, N:] == 0)
    # GELUGRAD: out = acc * aux, aux being the derivative the GELU epilogue stored
    u = torch.randn(M, 256)
    out = torch.full((M, 256), float('nan'), device='cuda')
    ops.gemm(Ab.cuda(), Wb.cuda(), K, K, M, N, K, ops.EPI_GELUGRAD, out, 256, n_out=256, aux=u.cuda(), ld_aux=256)
    assert rel(out[:, :N], (acc - bias.double()) * u[:, :N].double()) < 1e-5 and torch.all(out[:, N:] == 0)
    # RESIDUAL: out = res + [n < keep] * scale[sample] * (acc + bias); columns >= N copy the residual
    res = torch.randn(M, C)
    scale = torch.tensor([1.25, 0.0])
    keep = 160
    out = torch.full((M, C), float('nan'), device='cuda')
    ops.gemm(Ab.cuda(), Wb.cuda(), K, K, M, N, K, ops.EPI_RESIDUAL, out, C, n_out=C, bias=bias.cuda(), aux=res.cuda(),
             ld_aux=C, row_scale=scale.cuda(), rows_per_sample=257, n_keep=keep)
    ref = res.double().clone()
    ref[:, :keep] += scale.double().repeat_interleave(257).view(-1, 1) * acc[:, :keep]
    assert rel(out, ref) < 1e-5


def test_gemm_grouped(ops, tile_rows):
    """vsx_gemm_grouped: several problems of different shapes in ONE launch -- the q/k/v row blocks of a head-masked qkv projection
    (shared A, three weight / output / bias windows) and a set of split-K weight gradients -- against per-problem fp64 references."""
    g = torch.Generator().manual_seed(5)
    M, C, H, D, hk = 700, 160, 4, 64, 3
    HD = H * D
    x = torch.randn(M, C, generator=g).to(torch.bfloat16)
    w = (torch.randn(3 * HD, C, generator=g) * 0.1).to(torch.bfloat16)
    b = torch.randn(3 * HD, generator=g)
    qkv = torch.full((M, 3 * HD), float('nan'), device='cuda', dtype=torch.bfloat16)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    ops.gemm_grouped([((xd, wd, C, C, M, hk * D, 144, ops.EPI_STORE, qkv, 3 * HD),
                       dict(b_off=j * HD * C, out_off=j * HD, bias=bd, bias_off=j * HD)) for j in range(3)])
    ref = x[:, :144].double() @ w[:, :144].double().t() + b.double()
    got = qkv.view(M, 3, H, D)[:, :, :hk].double().cpu()
    assert rel(got, ref.view(M, 3, H, D)[:, :, :hk]) < 6e-3
    assert torch.isnan(qkv.view(M, 3, H, D)[:, :, hk:].float()).all()          # masked heads are not touched
    # weight gradients of different shapes: dW_i[n_i, k_i] += dy_i^T a_i, reductions over R rows split across CTAs
    R = 3000
    shapes = [(192, 160), (64, 96), (320, 40), (24, 200)]
    dys = [torch.randn(R, n, generator=g).to(torch.bfloat16) for n, _ in shapes]
    acs = [torch.randn(R, k, generator=g).to(torch.bfloat16) for _, k in shapes]
    dws = [torch.zeros(n, (k + 3) // 4 * 4, device='cuda') for n, k in shapes]
    dyd, acd = [t.cuda() for t in dys], [t.cuda() for t in acs]
    ops.gemm_grouped([((dyd[i], acd[i], n, k, n, k, R, ops.EPI_ATOMIC, dws[i], dws[i].shape[1]),
                       dict(a_layout=ops.MNMAJOR, b_layout=ops.MNMAJOR, split_k=7)) for i, (n, k) in enumerate(shapes)])
    for i, (n, k) in enumerate(shapes):
        assert rel(dws[i][:, :k], dys[i].double().t() @ acs[i].double()) < 1e-5, shapes[i]


# ---------------------------------------------------------------------------------------------- attention core
@pytest.mark.parametrize('N,H,D,Hk', [(257, 3, 64, 3), (257, 6, 32, 5), (65, 4, 48, 2), (17, 4, 64, 3), (50, 2, 32, 2), (197, 2, 64, 1)])
@pytest.mark.parametrize('mode', ['bf16_tcgen05', 'bf16_auto', 'bf16_mma', 'bf16_fp32math', 'fp32'])
def test_attention_core(ops, N, H, D, Hk, mode):
    """bf16_tcgen05 forces the tcgen05 / TMEM kernel (csrc/attn_tc.cu: head_dim 64, and 32 / 48 through zero-padded 4-D TMA boxes) and
    fails if a shape is not served by it; bf16_auto = what the model gets (the same kernel for every shape here); bf16_mma forces the
    legacy mma.sync kernel."""
    B = 3
    g = torch.Generator().manual_seed(N + H + D)
    dt_ = torch.float32 if mode == 'fp32' else torch.bfloat16
    qkv = (torch.randn(B, N, 3, H, D, generator=g) * 1.5).to(dt_)
    do = torch.randn(B, N, H, D, generator=g).to(dt_)
    q, k, v = [qkv[:, :, i].double().permute(0, 2, 1, 3).requires_grad_(True) for i in range(3)]   # [B,H,N,D]
    s = (q @ k.transpose(-1, -2)) * D ** -0.5
    p = s.softmax(-1)
    o_ref = (p @ v).permute(0, 2, 1, 3)                                       # [B,N,H,D]
    keep = torch.zeros(H, dtype=torch.float64)
    keep[:Hk] = 1
    (o_ref * do.double() * keep.view(1, 1, H, 1)).sum().backward()
    dqkv_ref = torch.stack([t.grad.permute(0, 2, 1, 3) for t in (q, k, v)], dim=2)   # [B,N,3,H,D]
    o_ref = o_ref.detach() * keep.view(1, 1, H, 1)
    lse_ref = torch.logsumexp(s.detach(), -1)                                 # [B,H,N]

    impl = {'bf16_fp32math': ops.ATTN_FP32, 'bf16_mma': ops.ATTN_MMA_SYNC, 'bf16_tcgen05': ops.ATTN_TCGEN05}.get(mode, ops.ATTN_AUTO)
    qd = qkv.cuda().view(B * N, 3 * H * D)
    o = torch.full((B * N, H * D), float('nan'), device='cuda', dtype=dt_)
    lse = torch.zeros(B, H, N, device='cuda')
    ops.attn_fwd(qd, o, lse, B, N, H, D, Hk, D ** -0.5, impl=impl)
    tol = 1e-5 if mode == 'fp32' else 8e-3
    assert rel(o.view(B, N, H, D), o_ref) < tol
    assert rel(lse[:, :Hk], lse_ref[:, :Hk]) < (1e-6 if mode == 'fp32' else 2e-3)
    dq = torch.full((B * N, 3 * H * D), float('nan'), device='cuda', dtype=dt_)
    ops.attn_bwd(qd, o, do.cuda().view(B * N, H * D), lse, dq, B, N, H, D, Hk, D ** -0.5, impl=impl)
    dq = dq.view(B, N, 3, H, D)
    assert torch.all(dq[:, :, :, Hk:] == 0) and torch.all(o.view(B, N, H, D)[:, :, Hk:] == 0)
    for i, nm in enumerate('qkv'):
        assert rel(dq[:, :, i], dqkv_ref[:, :, i]) < (2e-5 if mode == 'fp32' else 1.5e-2), nm


@pytest.mark.parametrize('B,N,H,Hk,D', [(70, 257, 4, 3, 64), (200, 65, 8, 5, 64), (256, 17, 12, 12, 64), (40, 288, 2, 2, 64), (33, 128, 3, 3, 64), (5, 97, 2, 1, 64),
                                        (48, 257, 8, 7, 32), (64, 65, 12, 10, 48), (40, 257, 6, 6, 32), (100, 17, 12, 6, 48), (9, 198, 2, 2, 32)])
def test_attention_tcgen05_persistent(ops, B, N, H, Hk, D):
    """The tcgen05 attention kernels with more (sample, head) pairs than SMs, so every CTA walks several pairs through its
    TMA / TMEM / mbarrier pipelines, against the fp32-math CUDA-core kernel on the same bf16 inputs (outputs, lse, dqkv and
    the fused qkv-bias gradient).  head_dim 32 / 48 are the shapes of sr_tiny_mh, sr_small and the searched Medium network."""
    g = torch.Generator().manual_seed(B + N + H)
    qkv = (torch.randn(B * N, 3 * H * D, generator=g) * 1.2).to(torch.bfloat16).cuda()
    do = torch.randn(B * N, H * D, generator=g).to(torch.bfloat16).cuda()
    res = {}
    for name, impl in (('ref', ops.ATTN_FP32), ('tc', ops.ATTN_TCGEN05)):
        o = torch.full((B * N, H * D), float('nan'), device='cuda', dtype=torch.bfloat16)
        lse = torch.zeros(B, H, N, device='cuda')
        ops.attn_fwd(qkv, o, lse, B, N, H, D, Hk, D ** -0.5, impl=impl)
        dq = torch.full((B * N, 3 * H * D), float('nan'), device='cuda', dtype=torch.bfloat16)
        db = torch.zeros(3 * H * D, device='cuda')
        ops.attn_bwd(qkv, res['ref'][0] if name == 'tc' else o, do, res['ref'][1] if name == 'tc' else lse, dq, B, N, H, D, Hk, D ** -0.5,
                     impl=impl, dbias=db)
        res[name] = (o, lse, dq, db)
    torch.cuda.synchronize()
    o_r, lse_r, dq_r, db_r = res['ref']
    o_t, lse_t, dq_t, db_t = res['tc']
    assert torch.all(o_t.view(B, N, H, D)[:, :, Hk:] == 0) and torch.all(dq_t.view(B, N, 3, H, D)[:, :, :, Hk:] == 0)
    assert rel(o_t, o_r) < 8e-3
    assert rel(lse_t[:, :Hk], lse_r[:, :Hk]) < 1e-5
    for i, nm in enumerate('qkv'):
        assert rel(dq_t.view(B, N, 3, H, D)[:, :, i], dq_r.view(B, N, 3, H, D)[:, :, i]) < 1.5e-2, nm
    assert rel(db_t, db_r) < 5e-3


@pytest.fixture
def odd_modes():
    from vit_search_b200 import _lib

    def force(f, b, s):
        _lib.check(_lib.lib().vsx_attn_odd_token_modes(f, b, s))
    yield force
    force(-1, -1, -1)


@pytest.mark.parametrize('modes', [(0, 0, 0), (1, 1, 1), (1, 2, 2), (1, 3, 3)])
@pytest.mark.parametrize('B,N,H,Hk,D', [(70, 257, 4, 3, 64), (150, 65, 8, 5, 64), (40, 257, 6, 6, 32), (64, 65, 12, 10, 48), (20, 129, 2, 2, 64)])
def test_attention_odd_token_modes(ops, odd_modes, modes, B, N, H, Hk, D):
    """The class token (the LAST of N = 2^k + 1 tokens) on the CUDA-core side warps of the tcgen05 kernels (csrc/attn_tc.cu, "the odd
    token"): every combination -- off, as a query, as a key, both -- against the fp32-math kernel, with more (sample, head) pairs than
    SMs so that the side warps run through their double-buffered side data and barrier phases several times.  The model uses (1, 3, 0)."""
    odd_modes(*modes)
    g = torch.Generator().manual_seed(B + N + H)
    qkv = (torch.randn(B * N, 3 * H * D, generator=g) * 1.2).to(torch.bfloat16).cuda()
    do = torch.randn(B * N, H * D, generator=g).to(torch.bfloat16).cuda()
    res = {}
    for name, impl in (('ref', ops.ATTN_FP32), ('tc', ops.ATTN_TCGEN05)):
        o = torch.full((B * N, H * D), float('nan'), device='cuda', dtype=torch.bfloat16)
        lse = torch.zeros(B, H, N, device='cuda')
        ops.attn_fwd(qkv, o, lse, B, N, H, D, Hk, D ** -0.5, impl=impl)
        dq = torch.full((B * N, 3 * H * D), float('nan'), device='cuda', dtype=torch.bfloat16)
        db = torch.zeros(3 * H * D, device='cuda')
        ops.attn_bwd(qkv, res['ref'][0] if name == 'tc' else o, do, res['ref'][1] if name == 'tc' else lse, dq, B, N, H, D, Hk, D ** -0.5,
                     impl=impl, dbias=db)
        res[name] = (o, lse, dq, db)
    torch.cuda.synchronize()
    (o_r, lse_r, dq_r, db_r), (o_t, lse_t, dq_t, db_t) = res['ref'], res['tc']
    assert torch.all(o_t.view(B, N, H, D)[:, :, Hk:] == 0) and torch.all(dq_t.view(B, N, 3, H, D)[:, :, :, Hk:] == 0)
    assert rel(o_t, o_r) < 8e-3 and rel(lse_t[:, :Hk], lse_r[:, :Hk]) < 1e-5
    # the odd token's own rows separately: they are a 1 / N share of the tensors above
    assert rel(o_t.view(B, N, H * D)[:, -1], o_r.view(B, N, H * D)[:, -1]) < 8e-3
    assert rel(lse_t[:, :Hk, -1], lse_r[:, :Hk, -1]) < 1e-5
    for i, nm in enumerate('qkv'):
        a, b = dq_t.view(B, N, 3, H, D)[:, :, i], dq_r.view(B, N, 3, H, D)[:, :, i]
        assert rel(a, b) < 1.5e-2, nm
        assert rel(a[:, -1], b[:, -1]) < 1.5e-2, nm + ' (odd token)'
    assert rel(db_t, db_r) < 5e-3


# ---------------------------------------------------------------------------------------------- direct 3x3 conv (stem)
@pytest.fixture
def conv_impl():
    from vit_search_b200 import _lib

    def force(impl):
        _lib.check(_lib.lib().vsx_conv3x3_force_impl(impl))
    yield force
    force(0)


@pytest.mark.parametrize('C,impl', [(24, 0), (24, 1), (24, 2), (32, 0)])
def test_conv3x3_direct(ops, conv_impl, C, impl):
    """Tensor-core direct conv (forward with fused BN+ReLU input and batch statistics, data gradient with the BN-backward
    reductions, weight gradient) against torch fp64 math on the same bf16-rounded operands.  impl 1: legacy kernel, 2: TMA kernel."""
    import torch.nn.functional as F
    from vit_search_b200 import core
    conv_impl(impl)
    B, H, W = 2, 16, 32
    g = torch.Generator().manual_seed(C)
    y_in = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16)
    sc, sh = 1 + 0.2 * torch.randn(C, generator=g), 0.3 * torch.randn(C, generator=g)
    wgt = (torch.randn(C, C, 3, 3, generator=g) * 0.1)
    wq = wgt.to(torch.bfloat16).double()
    a = F.relu(y_in.double() * sc.double() + sh.double()).to(torch.bfloat16).double()      # the kernel rounds the activated halo to bf16
    ref = F.conv2d(a.permute(0, 3, 1, 2), wq, padding=1).permute(0, 2, 3, 1)               # [B,H,W,C]
    wparam = torch.nn.Parameter(wgt.cuda())
    out = torch.full((B, H, W, C), float('nan'), device='cuda', dtype=torch.bfloat16)
    sums = torch.zeros(2 * C, device='cuda', dtype=torch.float64)
    ops.call('conv3x3', y_in.cuda(), sc.cuda(), sh.cuda(), core.weights.get(wparam, 'conv3x3_fwd'), None, out, B, H, W, C, 1, None, None,
             None, None, None, sums)
    assert rel(out, ref) < 6e-3
    o64 = out.double().cpu()
    assert rel(sums[:C], o64.sum((0, 1, 2))) < 1e-5 and rel(sums[C:], (o64 * o64).sum((0, 1, 2))) < 1e-5
    # data gradient: dy -> d_a = conv_transpose(dy, W) + add, with (sum dz, sum dz*zhat) of the previous BN
    dy = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16)
    add = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16)
    gam, bet = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    mean, rstd = 0.1 * torch.randn(C, generator=g), 1 + 0.1 * torch.rand(C, generator=g)
    ref_da = F.conv_transpose2d(dy.double().permute(0, 3, 1, 2), wq, padding=1).permute(0, 2, 3, 1) + add.double()
    d_a = torch.empty(B, H, W, C, device='cuda', dtype=torch.bfloat16)
    sums2 = torch.zeros(2 * C, device='cuda', dtype=torch.float64)
    ops.call('conv3x3', dy.cuda(), None, None, core.weights.get(wparam, 'conv3x3_bwd'), add.cuda(), d_a, B, H, W, C, 2, y_in.cuda(),
             gam.cuda(), bet.cuda(), mean.cuda(), rstd.cuda(), sums2)
    assert rel(d_a, ref_da) < 6e-3
    zh = (y_in.float() - mean) * rstd
    dz = d_a.float().cpu() * (gam * zh + bet > 0)
    assert rel(sums2[:C], dz.double().sum((0, 1, 2))) < 1e-4 and rel(sums2[C:], (dz * zh).double().sum((0, 1, 2))) < 1e-4
    # weight gradient: dW[co][ci][ky][kx] = sum dy[p][co] * a[p + tap - 1][ci]
    wref = wq.clone().requires_grad_(True)
    (F.conv2d(a.permute(0, 3, 1, 2), wref, padding=1) * dy.double().permute(0, 3, 1, 2)).sum().backward()
    dw = torch.zeros(C, 9 * C, device='cuda')
    ops.call('conv3x3_wgrad', dy.cuda(), y_in.cuda(), sc.cuda(), sh.cuda(), dw, B, H, W, C)
    assert rel(dw.view(C, 3, 3, C).permute(0, 3, 1, 2), wref.grad) < 2e-3


def test_conv3x3_tma_matches_legacy_at_stem_size(ops, conv_impl):
    """The TMA warp-specialised kernel (csrc/conv3x3_tma.cu) against the legacy direct kernel at the stem's map size, enough tiles
    per CTA to wrap the 4-stage ring several times: both accumulate the same packed reduction in the same order, so the bf16
    outputs must be bit-identical; the fused statistics agree to fp32 summation order.  Every mode: activated input + forward
    statistics, plain, data gradient with / without the residual gradient."""
    from vit_search_b200 import core
    B, H, W, C = 24, 112, 112, 24
    g = torch.Generator(device='cuda').manual_seed(7)
    x = torch.randn(B, H, W, C, device='cuda', generator=g).to(torch.bfloat16)
    y_prev = torch.randn(B, H, W, C, device='cuda', generator=g).to(torch.bfloat16)
    add = torch.randn(B, H, W, C, device='cuda', generator=g).to(torch.bfloat16)
    sc, sh = 1 + 0.2 * torch.randn(C, device='cuda', generator=g), 0.3 * torch.randn(C, device='cuda', generator=g)
    gam, bet = 1 + 0.1 * torch.randn(C, device='cuda', generator=g), 0.1 * torch.randn(C, device='cuda', generator=g)
    mean, rstd = 0.1 * torch.randn(C, device='cuda', generator=g), 1 + 0.1 * torch.rand(C, device='cuda', generator=g)
    wparam = torch.nn.Parameter(torch.randn(C, C, 3, 3, device='cuda', generator=g) * 0.1)
    wf, wb = core.weights.get(wparam, 'conv3x3_fwd'), core.weights.get(wparam, 'conv3x3_bwd')
    modes = [(sc, sh, wf, None, 1, None), (None, None, wf, None, 0, None), (None, None, wb, add, 2, y_prev), (None, None, wb, None, 2, y_prev),
             (sc, sh, wf, add, 0, None)]
    for a_sc, a_sh, w, ad, stats, yp in modes:
        res = []
        for impl in (1, 2):
            conv_impl(impl)
            out = torch.full((B, H, W, C), float('nan'), device='cuda', dtype=torch.bfloat16)
            sums = torch.zeros(2 * C, device='cuda', dtype=torch.float64)
            bn = (gam, bet, mean, rstd) if stats == 2 else (None,) * 4
            ops.call('conv3x3', x, a_sc, a_sh, w, ad, out, B, H, W, C, stats, yp, *bn, sums if stats else None)
            torch.cuda.synchronize()
            res.append((out, sums))
        assert torch.equal(res[0][0], res[1][0]), (stats, ad is not None)
        if stats:
            assert rel(res[1][1], res[0][1]) < 1e-6
    # weight gradient: same products, different fp32 summation order
    dws = []
    for impl in (1, 2):
        conv_impl(impl)
        for a_sc, a_sh in ((sc, sh), (None, None)):
            dw = torch.zeros(C, 9 * C, device='cuda')
            ops.call('conv3x3_wgrad', add, x, a_sc, a_sh, dw, B, H, W, C)
            torch.cuda.synchronize()
            dws.append(dw)
    assert rel(dws[2], dws[0]) < 2e-5 and rel(dws[3], dws[1]) < 2e-5


@pytest.mark.parametrize('shape', [dict(B=3, H=28, W=28, C=24, k=7, s=7, p=0, dual=True),      # conv_proj: patches, BN+ReLU of two maps
                                   dict(B=2, H=16, W=16, C=64, k=3, s=2, p=1, tok=True),        # SR conv on a token tensor (class token skipped)
                                   dict(B=2, H=8, W=8, C=128, k=3, s=2, p=1, tok=True),
                                   dict(B=2, H=12, W=12, C=12, k=3, s=1, p=1)])                 # C % 8 != 0: the generic kernels
def test_im2col_col2im_bf16(ops, shape):
    """im2col / col2im (bf16 channels-last; the 16-byte row kernels where C % 8 == 0) against torch unfold / fold."""
    import torch.nn.functional as F
    B, H, W, C, k, s, p = (shape[n] for n in 'B H W C k s p'.split())
    tok, dual = shape.get('tok', False), shape.get('dual', False)
    g = torch.Generator().manual_seed(H * C + k)
    N = H * W + (1 if tok else 0)
    x = torch.randn(B, N, C, generator=g).to(torch.bfloat16)
    x2 = torch.randn(B, N, C, generator=g).to(torch.bfloat16)
    sc1, sh1 = 1 + 0.2 * torch.randn(C, generator=g), 0.3 * torch.randn(C, generator=g)
    sc2, sh2 = 1 + 0.2 * torch.randn(C, generator=g), 0.3 * torch.randn(C, generator=g)
    off = C if tok else 0
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1

    def to_cols(t):          # [B, H*W, C] fp32 -> [B*Ho*Wo, k*k*C] with tap-major, channel-minor columns
        u = F.unfold(t.view(B, H, W, C).permute(0, 3, 1, 2), k, padding=p, stride=s)
        return u.view(B, C, k * k, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, k * k * C)

    img = x[:, 1:] if tok else x
    if dual:
        a = F.relu(img.float() * sc1 + sh1) + F.relu((x2[:, 1:] if tok else x2).float() * sc2 + sh2)
    else:
        a = img.float()
    ref = to_cols(a).to(torch.bfloat16)
    out = torch.full((B * Ho * Wo, k * k * C), float('nan'), device='cuda', dtype=torch.bfloat16)
    xd, x2d = x.cuda(), x2.cuda()
    if dual:
        ops.call('im2col', (xd, off), sc1.cuda(), sh1.cuda(), (x2d, off), sc2.cuda(), sh2.cuda(), ops.BF16, 0, N * C, C, B, H, W, C, k, s, p, out,
                 ops.BF16, k * k * C)
    else:
        ops.call('im2col', (xd, off), None, None, None, None, None, ops.BF16, 0, N * C, C, B, H, W, C, k, s, p, out, ops.BF16, k * k * C)
    if dual:       # the kernel's fused multiply-add rounds once where torch rounds twice: rare 1-ulp differences after the bf16 rounding
        assert rel(out.cpu(), ref) < 1e-3 and not torch.isnan(out).any()
    else:
        assert torch.equal(out.cpu(), ref)
    # col2im: the adjoint (+ optional add), fp32 accumulation, one rounding
    dcol = torch.randn(B * Ho * Wo, k * k * C, generator=g).to(torch.bfloat16)
    add = torch.randn(B, N, C, generator=g).to(torch.bfloat16)
    fold_in = dcol.float().view(B, Ho * Wo, k * k, C).permute(0, 3, 2, 1).reshape(B, C * k * k, Ho * Wo)
    ref_d = F.fold(fold_in, (H, W), k, padding=p, stride=s).permute(0, 2, 3, 1).reshape(B, H * W, C)
    for use_add in (False, True):
        din = torch.zeros(B, N, C, device='cuda', dtype=torch.bfloat16)
        addd = add.cuda()
        ops.call('col2im', dcol.cuda(), k * k * C, (addd, off) if use_add else None, ops.BF16, B, H, W, C, k, s, p, (din, off), N * C, C)
        want = ref_d + ((add[:, 1:] if tok else add).float() if use_add else 0)
        got = din.cpu()[:, 1:] if tok else din.cpu()
        assert rel(got, want) < 4e-3
        if tok:
            assert torch.all(din[:, 0] == 0)          # the class-token row is not touched


@pytest.mark.parametrize('C,keep,cast_keep', [(256, 256, 256), (512, 448, 384), (1024, 1024, 640), (64, 40, 64)])
def test_ln_bwd_with_fused_cast(ops, C, keep, cast_keep):
    """vsx_masked_ln_bwd_cast == vsx_masked_ln_bwd followed by vsx_scale_mask_cast on its output (the fused kernel where the bulk-copy
    variant applies, the two-launch fallback otherwise: keep = 40 is not a multiple of 8)."""
    B, N = 3, 65
    rows = B * N
    g = torch.Generator(device='cuda').manual_seed(C + keep)
    dy = torch.randn(rows, C, device='cuda', generator=g).to(torch.bfloat16)
    x = torch.randn(rows, C, device='cuda', generator=g)
    g_in = torch.randn(rows, C, device='cuda', generator=g)
    gamma = 1 + 0.1 * torch.randn(C, device='cuda', generator=g)
    scale = torch.rand(B + 2, device='cuda', generator=g) + 0.5
    mean = x[:, :keep].mean(1).contiguous()
    rstd = (1.0 / torch.sqrt(x[:, :keep].var(1, unbiased=False) + 1e-6)).contiguous()
    ref_g = torch.empty(rows, C, device='cuda')
    dgam, dbet = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    ops.masked_ln_bwd(dy, C, x, C, mean, rstd, gamma, g_in, ref_g, C, dgam, dbet, rows, C, keep)
    ref_cast = torch.empty(rows, C, device='cuda', dtype=torch.bfloat16)
    ref_cs = torch.zeros(C, device='cuda')
    ops.scale_mask_cast(ref_g, C, scale, N, cast_keep, ref_cast, C, rows, C, scale_off=1, colsum=ref_cs)
    got_g = torch.empty(rows, C, device='cuda')
    got_cast = torch.full((rows, C), float('nan'), device='cuda', dtype=torch.bfloat16)
    dgam2, dbet2, got_cs = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    ops.call('masked_ln_bwd_cast', dy, ops.BF16, C, x, C, mean, rstd, gamma, g_in, got_g, C, dgam2, dbet2, rows, C, keep, got_cast, C, (scale, 1), N,
             cast_keep, got_cs)
    assert torch.equal(got_g, ref_g) and torch.equal(got_cast, ref_cast)
    assert rel(dgam2, dgam) < 1e-5 and rel(dbet2, dbet) < 1e-5 and rel(got_cs, ref_cs) < 1e-5
    assert torch.all(got_cast[:, cast_keep:] == 0)


@pytest.mark.parametrize('B,H,W', [(2, 32, 48), (3, 224, 224), (1, 6, 10)])
def test_conv1_direct(ops, B, H, W):
    """First stem convolution (3 -> 24, 3x3, stride 2, pad 1) straight from the fp32 image: output map, fused batch statistics and the
    weight gradient against torch fp64 math on the same bf16-rounded operands (ragged tails: 15 and 3 * 5 pixels are not multiples of 32)."""
    import torch.nn.functional as F
    from vit_search_b200 import core
    g = torch.Generator().manual_seed(H + W)
    x = torch.randn(B, 3, H, W, generator=g)
    wgt = torch.randn(24, 3, 3, 3, generator=g) * 0.2
    wparam = torch.nn.Parameter(wgt.cuda())
    wc = core.weights.get(wparam, 'ohwi')
    xq, wq = x.to(torch.bfloat16).double(), wgt.to(torch.bfloat16).double()
    ref = F.conv2d(xq, wq, stride=2, padding=1).permute(0, 2, 3, 1)                # [B, H/2, W/2, 24]
    y = torch.full((B, H // 2, W // 2, 24), float('nan'), device='cuda', dtype=torch.bfloat16)
    sums = torch.zeros(48, device='cuda', dtype=torch.float64)
    ops.call('conv1_fwd', x.cuda(), wc, core.ld_of(wc), y, B, H, W, sums)
    assert rel(y, ref) < 4e-3 and not torch.isnan(y.float()).any()
    y64 = y.double().cpu()
    assert rel(sums[:24], y64.sum((0, 1, 2))) < 1e-5 and rel(sums[24:], (y64 * y64).sum((0, 1, 2))) < 1e-5
    y2 = torch.empty_like(y)
    ops.call('conv1_fwd', x.cuda(), wc, core.ld_of(wc), y2, B, H, W, None)          # eval mode: no statistics
    assert torch.equal(y2, y)
    dy = torch.randn(B, H // 2, W // 2, 24, generator=g).to(torch.bfloat16)
    wref = wq.clone().requires_grad_(True)
    (F.conv2d(xq, wref, stride=2, padding=1) * dy.double().permute(0, 3, 1, 2)).sum().backward()
    dw = torch.zeros(24, 28, device='cuda')
    ops.call('conv1_wgrad', x.cuda(), dy.cuda(), dw, 28, B, H, W)
    assert rel(dw[:, :27].reshape(24, 3, 3, 3).permute(0, 3, 1, 2), wref.grad) < 2e-3
    assert torch.all(dw[:, 27] == 0)


def test_gemm_grouped_wide(ops, tile_rows):
    """Up to 16 single-term problems in one launch: the q / k / v row blocks of FOUR segments with different kept heads / embedding widths
    (a multi-architecture batch), and the GELU' data gradients of four segments adding their bias-gradient column sums into one vector."""
    g = torch.Generator().manual_seed(11)
    C, H, D = 160, 4, 64
    HD = H * D
    segs = [(300, 3, 144), (257, 4, 160), (514, 2, 128), (130, 1, 96)]          # rows, kept heads, kept embedding width
    M = sum(r for r, _, _ in segs)
    x = torch.randn(M, C, generator=g).to(torch.bfloat16)
    w = (torch.randn(3 * HD, C, generator=g) * 0.1).to(torch.bfloat16)
    b = torch.randn(3 * HD, generator=g)
    qkv = torch.full((M, 3 * HD), float('nan'), device='cuda', dtype=torch.bfloat16)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    probs, r0 = [], 0
    for rows, hk, ek in segs:
        for j in range(3):
            probs.append(((xd, wd, C, C, rows, hk * D, ek, ops.EPI_STORE, qkv, 3 * HD),
                          dict(a_off=r0 * C, b_off=j * HD * C, out_off=r0 * 3 * HD + j * HD, bias=bd, bias_off=j * HD)))
        r0 += rows
    assert len(probs) == 12
    ops.gemm_grouped(probs)
    r0 = 0
    for rows, hk, ek in segs:
        ref = (x[r0:r0 + rows, :ek].double() @ w[:, :ek].double().t() + b.double()).view(rows, 3, H, D)[:, :, :hk]
        got = qkv[r0:r0 + rows].view(rows, 3, H, D)[:, :, :hk].double().cpu()
        assert rel(got, ref) < 6e-3, (rows, hk, ek)
        assert torch.isnan(qkv[r0:r0 + rows].view(rows, 3, H, D)[:, :, hk:].float()).all()
        r0 += rows
    # GELU' data gradients of four segments: du = (df W2) * aux (aux = the stored derivative), column sums of all segments into ONE vector
    F = 384
    df = torch.randn(M, C, generator=g).to(torch.bfloat16)
    w2 = (torch.randn(C, F, generator=g) * 0.1).to(torch.bfloat16)
    u = torch.randn(M, F, generator=g).to(torch.bfloat16)
    du = torch.zeros(M, F, device='cuda', dtype=torch.bfloat16)
    db = torch.zeros(F, device='cuda')
    dfd, w2d, ud = df.cuda(), w2.cuda(), u.cuda()
    probs, r0, keeps = [], 0, [(160, 384), (144, 256), (128, 320), (96, 192)]
    for (rows, _, _), (ck, ik) in zip(segs, keeps):
        probs.append(((dfd, w2d, C, F, rows, ik, ck, ops.EPI_GELUGRAD, du, F),
                      dict(a_off=r0 * C, out_off=r0 * F, n_out=ik, aux=ud, ld_aux=F, aux_off=r0 * F, b_layout=ops.MNMAJOR, colsum=db)))
        r0 += rows
    ops.gemm_grouped(probs)
    ref_db = torch.zeros(F, dtype=torch.float64)
    r0 = 0
    for (rows, _, _), (ck, ik) in zip(segs, keeps):
        ref = (df[r0:r0 + rows, :ck].double() @ w2[:ck, :ik].double()) * u[r0:r0 + rows, :ik].double()
        assert rel(du[r0:r0 + rows, :ik], ref) < 8e-3
        ref_db[:ik] += ref.sum(0)
        r0 += rows
    assert rel(db, ref_db) < 5e-3


@pytest.mark.parametrize('N,H,D', [(257, 4, 64), (65, 12, 48), (17, 6, 32)])
def test_attention_segments_one_launch(ops, N, H, D):
    """vsx_attn_fwd_segs / vsx_attn_bwd_segs: a batch whose consecutive sample ranges keep different numbers of heads (one range dropped
    entirely) in ONE launch of the tcgen05 kernels, against per-segment launches of the fp32-math kernel."""
    import ctypes as C
    from vit_search_b200 import _lib
    ranges = [(20, H), (13, max(1, H // 2)), (7, 0), (30, H - 1)]
    B = sum(n for n, _ in ranges)
    g = torch.Generator().manual_seed(N + H)
    qkv = (torch.randn(B * N, 3 * H * D, generator=g) * 1.2).to(torch.bfloat16).cuda()
    do = torch.randn(B * N, H * D, generator=g).to(torch.bfloat16).cuda()
    sg = _lib.SampleSegments()
    sg.count = len(ranges)
    e = 0
    for i, (n, hk) in enumerate(ranges):
        e += n
        sg.sample_end[i], sg.heads_keep[i] = e, hk
    o = torch.full((B * N, H * D), float('nan'), device='cuda', dtype=torch.bfloat16)
    lse = torch.zeros(B, H, N, device='cuda')
    dq = torch.full((B * N, 3 * H * D), float('nan'), device='cuda', dtype=torch.bfloat16)
    db = torch.zeros(3 * H * D, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().vsx_attn_fwd_segs(qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), _lib.BF16, B, N, H, D, C.byref(sg), D ** -0.5, ops.ATTN_TCGEN05, st))
    _lib.check(_lib.lib().vsx_attn_bwd_segs(qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(), _lib.BF16, B, N, H, D, C.byref(sg),
                                            D ** -0.5, ops.ATTN_TCGEN05, db.data_ptr(), st))
    torch.cuda.synchronize()
    db_ref = torch.zeros(3 * H * D, device='cuda')
    b0 = 0
    for n, hk in ranges:
        rows = slice(b0 * N, (b0 + n) * N)
        if hk == 0:
            assert torch.isnan(o[rows].float()).all() and torch.isnan(dq[rows].float()).all()       # dropped samples are not touched
            b0 += n
            continue
        o_r = torch.full((n * N, H * D), float('nan'), device='cuda', dtype=torch.bfloat16)
        lse_r = torch.zeros(n, H, N, device='cuda')
        dq_r = torch.full((n * N, 3 * H * D), float('nan'), device='cuda', dtype=torch.bfloat16)
        ops.attn_fwd(qkv[rows], o_r, lse_r, n, N, H, D, hk, D ** -0.5, impl=ops.ATTN_FP32)
        ops.attn_bwd(qkv[rows], o_r, do[rows], lse_r, dq_r, n, N, H, D, hk, D ** -0.5, impl=ops.ATTN_FP32, dbias=db_ref)
        assert rel(o[rows], o_r) < 8e-3 and rel(lse[b0:b0 + n, :hk], lse_r[:, :hk]) < 1e-5
        assert torch.all(o[rows].view(n * N, H, D)[:, hk:] == 0) and torch.all(dq[rows].view(n * N, 3, H, D)[:, :, hk:] == 0)
        assert rel(dq[rows], dq_r) < 1.5e-2
        b0 += n
    assert rel(db, db_ref) < 8e-3
