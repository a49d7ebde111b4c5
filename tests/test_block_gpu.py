"""GPU parity of one transformer Block (both halves, forward + backward) against the CPU oracle."""
import pytest
import torch

from oracle import vit_res_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


CASES = [
    # C, H, D, F, N, per-sample keeps: embed, attn, layer, mlp, drop-path scales (attn, mlp)
    dict(C=64, H=2, D=32, F=128, N=257, embed=[64, 64, 44, 44], attn=[64, 64, 32, 32], layer=None, mlp=[128, 128, 96, 64]),
    dict(C=128, H=2, D=48, F=256, N=65, embed=[128, 112, 100, 80], attn=[96, 48, 96, 48], layer=[128, 0, 128, 128],
         mlp=[256, 192, 128, 256], layer_in=[128, 128, 0, 128], dp=([1.25, 0., 1.25, 1.25], [0., 1.25, 1.25, 1.25])),
    dict(C=256, H=4, D=64, F=512, N=17, embed=None, attn=None, layer=None, mlp=None),
    dict(C=256, H=4, D=64, F=512, N=17, embed=[256, 200, 160, 160, 224, 224], attn=[256, 192, 128, 128, 64, 64], layer=None,
         mlp=[512, 384, 256, 256, 512, 512]),
]


@pytest.mark.parametrize('prec', ['fp32', 'bf16', 'bf16_python'])     # bf16: native half-block calls; bf16_python: core.py orchestration
@pytest.mark.parametrize('ci', range(len(CASES)))
def test_block_parity(ci, prec):
    from vit_search_b200 import core
    from vit_search_b200.nets import Block
    c = CASES[ci]
    C, H, D, F, N = c['C'], c['H'], c['D'], c['F'], c['N']
    B = len(c['embed']) if c['embed'] is not None else 3
    shapes = {'norm1.weight': (C,), 'norm1.bias': (C,), 'attn.qkv.weight': (3 * H * D, C), 'attn.qkv.bias': (3 * H * D,),
              'attn.proj.weight': (C, H * D), 'attn.proj.bias': (C,), 'norm2.weight': (C,), 'norm2.bias': (C,),
              'mlp.fc1.weight': (F, C), 'mlp.fc1.bias': (F,), 'mlp.fc2.weight': (C, F), 'mlp.fc2.bias': (C,)}
    w = O.keyed_fill(shapes, seed=ci)
    for k in w:                                   # larger weights than init so the branches matter
        if k.endswith('weight') and w[k].ndim == 2:
            w[k] = w[k] * 3
    g = torch.Generator().manual_seed(100 + ci)
    x = torch.randn(B, N, C, generator=g)
    if c['embed'] is not None:
        x = x * O.prefix_mask(c['embed'], C, torch.float32)
    gout = torch.randn(B, N, C, generator=g)
    keeps = {k: c[k] for k in ('attn', 'layer', 'mlp') if c.get(k) is not None}
    dp = c.get('dp')

    # oracle in fp64
    p = {'b.' + k: v.double().requires_grad_(True) for k, v in w.items()}
    xo = x.double().requires_grad_(True)
    dpk = None
    if dp is not None:
        dpk = ([1.0 if v > 0 else 0.0 for v in dp[0]], [1.0 if v > 0 else 0.0 for v in dp[1]])
    yo, _, cur_o = O.block(xo, p, 'b.', H, D, c['embed'], c.get('layer_in'), keeps, dpk, 0.2 if dp is not None else 0.0)
    yo.backward(gout.double())

    blk = Block(C, H, D, F).cuda()
    blk.load_state_dict(w)
    xd = x.cuda().requires_grad_(True)
    dpt = None if dp is None else torch.tensor(dp, dtype=torch.float32).cuda().contiguous()
    native0 = core.USE_NATIVE_HALF
    core.USE_NATIVE_HALF = prec != 'bf16_python'
    prec = 'bf16' if prec == 'bf16_python' else prec
    try:
        with core.precision(prec):
            y, cur = blk.forward_keeps(xd, c['embed'], c.get('layer_in'), keeps, dpt, 0)
            y.backward(gout.cuda())
        torch.cuda.synchronize()
    finally:
        core.USE_NATIVE_HALF = native0
    assert cur == cur_o
    tol = 2e-5 if prec == 'fp32' else 2e-2      # fp32 parity path / bf16 training path, rel-L2 vs the fp64 oracle
    errs = {'y': rel(y, yo)}
    m = O.prefix_mask(c['embed'], C, torch.float64) if c['embed'] is not None else 1.0
    errs['gx'] = rel(xd.grad.double().cpu() * m, xo.grad * m)       # gradients on masked input lanes are dead
    for k, prm in blk.named_parameters():
        errs[k] = rel(prm.grad, p['b.' + k].grad)
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, (prec, bad, errs)


@pytest.mark.parametrize('prec', ['fp32', 'bf16', 'bf16_python'])
def test_standalone_mlp_and_attention_modules(prec):
    """Mlp.forward / Attention.forward used on their own (reference nets/supernet_blocks.py:37-52, :100-120: no LayerNorm in front, no
    residual) against plain torch fp64 math, forward and backward -- the pre_norm=False / residual=False route of the half-block calls."""
    import torch.nn.functional as Fn
    from vit_search_b200 import core
    from vit_search_b200.nets import Attention, Mlp
    B, N, C, H, D, F = 3, 65, 128, 2, 64, 256
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, N, C, generator=g)
    gout = torch.randn(B, N, C, generator=g)
    native0 = core.USE_NATIVE_HALF
    core.USE_NATIVE_HALF = prec != 'bf16_python'
    p = 'bf16' if prec == 'bf16_python' else prec
    tol = 2e-5 if p == 'fp32' else 2e-2
    try:
        for mod in (Mlp(C, F).cuda(), Attention(C, H, D).cuda()):
            for prm in mod.parameters():
                with torch.no_grad():
                    prm.copy_(torch.randn(prm.shape, generator=g) * (0.2 if prm.ndim == 2 else 0.1))
            xd = x.cuda().requires_grad_(True)
            with core.precision(p):
                y = mod(xd)
                y.backward(gout.cuda())
            torch.cuda.synchronize()
            w = {k: v.detach().double().cpu().requires_grad_(True) for k, v in mod.named_parameters()}
            xo = x.double().requires_grad_(True)
            if isinstance(mod, Mlp):
                yo = Fn.linear(Fn.gelu(Fn.linear(xo, w['fc1.weight'], w['fc1.bias'])), w['fc2.weight'], w['fc2.bias'])
            else:
                qkv = Fn.linear(xo, w['qkv.weight'], w['qkv.bias']).view(B, N, 3, H, D).permute(2, 0, 3, 1, 4)
                att = ((qkv[0] @ qkv[1].transpose(-1, -2)) * D ** -0.5).softmax(-1)
                yo = Fn.linear((att @ qkv[2]).transpose(1, 2).reshape(B, N, H * D), w['proj.weight'], w['proj.bias'])
            yo.backward(gout.double())
            errs = {'y': rel(y, yo), 'gx': rel(xd.grad, xo.grad)}
            for k, prm in mod.named_parameters():
                errs[k] = rel(prm.grad, w[k].grad)
            bad = {k: v for k, v in errs.items() if not v < tol}
            assert not bad, (type(mod).__name__, prec, bad)
    finally:
        core.USE_NATIVE_HALF = native0


@pytest.mark.parametrize('with_embed', [True, False])
@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_block_reference_signature_with_bool_masks(with_embed, prec):
    """Block.forward(x, embed_mask, layer_mask) exactly as the reference calls it (nets/supernet_blocks.py:209-255): UNTAGGED [B,1,C]
    bool tensors in (keep counts are read back from the device), the block draws its own attn / layer / mlp masks from its
    ChannelDrops, and returns (x, embed_mask, current_layer_mask) as bool tensors.  `with_embed=False` is the case in which the
    reference masks the attention branch with the block's OWN layer mask only and the MLP branch with own & incoming (:220-251)."""
    import numpy as np
    from vit_search_b200 import core
    from vit_search_b200.nets import Block, ChannelDrop
    C, H, D, F, N, B = 128, 2, 64, 256, 65, 8
    attn_ch, mlp_ch, layer_ch = [128, 64], [256, 192, 128], [128, 128, 0, 0]
    blk = Block(C, H, D, F, num_chs_to_keep_attn=np.array(attn_ch), num_chs_to_keep_mlp=np.array(mlp_ch),
                num_chs_to_keep_block=np.array(layer_ch), num_warmup_epochs=0, example_per_arch=2).cuda()
    shapes = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
    w = O.keyed_fill(shapes, seed=5)
    for k in w:
        if k.endswith('weight') and w[k].ndim == 2:
            w[k] = w[k] * 3
    blk.load_state_dict(w)
    for mod in blk.modules():
        if isinstance(mod, ChannelDrop):
            mod.set_epoch(0)
    blk.train()
    embed = [128, 128, 112, 112, 100, 100, 80, 80]
    layer_in = [128, 0, 128, 128, 0, 128, 128, 128]
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, N, C, generator=g)
    if with_embed:
        x = x * O.prefix_mask(embed, C, torch.float32)
    gout = torch.randn(B, N, C, generator=g)
    embed_mask = O.prefix_mask(embed, C).cuda() if with_embed else None          # plain bool tensors, no keep-count tag
    layer_mask = O.prefix_mask(layer_in, C).cuda()
    xd = x.cuda().requires_grad_(True)
    with core.precision(prec):
        torch.manual_seed(77)
        y, em, cur = blk(xd, embed_mask, layer_mask)
        y.backward(gout.cuda())
    torch.cuda.synchronize()
    # the same draws through the oracle's restatement of ChannelDrop (order: attn, layer, mlp -- SURVEY.md A3)
    torch.manual_seed(77)
    keeps = {}
    for key, ch in (('attn', attn_ch), ('layer', layer_ch), ('mlp', mlp_ch)):
        keeps[key] = O.draw_keep(O.keep_table(ch, B, 2, False, 0, 0), B, 2, False)
    p = {'b.' + k: v.double().requires_grad_(True) for k, v in w.items()}
    xo = x.double().requires_grad_(True)
    yo, _, cur_o = O.block(xo, p, 'b.', H, D, embed if with_embed else None, layer_in, keeps)
    yo.backward(gout.double())
    assert em is embed_mask
    assert cur.dtype == torch.bool and tuple(cur.shape) == (B, 1, C)
    assert cur.sum(dim=(1, 2)).tolist() == cur_o
    assert len(set(zip(keeps['attn'], keeps['layer'], keeps['mlp']))) > 1          # the draw really mixed architectures
    tol = 2e-5 if prec == 'fp32' else 2e-2
    m = O.prefix_mask(embed, C, torch.float64) if with_embed else 1.0
    errs = {'y': rel(y, yo), 'gx': rel(xd.grad.double().cpu() * m, xo.grad * m)}
    for k, prm in blk.named_parameters():
        errs[k] = rel(prm.grad, p['b.' + k].grad)
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, (prec, with_embed, bad)
