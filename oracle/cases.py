"""Parity cases shared by oracle/make_golden.py and tests/ (test infrastructure)."""
import numpy as np

# A deliberately small ViT-Res (image 224, patch 14 are fixed by the reference, so N = 257/65/17 as in
# the real nets) that still exercises: conv stem, head dims 32/48/64, a skippable block, a BypassBlock
# (exists=0), both SR blocks, and keep counts that are not multiples of 8.
SMALL_DEF = ((4, 64),
             (1, (64, 2, 32), (64, 128), 1), (1, (64, 2, 32), (64, 128), 1),
             (3, 64, 128),
             (1, (128, 2, 48), (128, 256), 1), (1, (128, 2, 48), (128, 256), 0), (1, (128, 2, 48), (128, 256), 1),
             (3, 128, 256),
             (1, (256, 4, 64), (256, 512), 1), (1, (256, 4, 64), (256, 512), 1),
             (2, 256, 1000))


def _blk(attn, mlp, layer=None):
    return {'attn': np.array(attn), 'mlp': np.array(mlp), 'layer': None if layer is None else np.array(layer)}


SMALL_SPACE = [np.array([64, 56, 44, 40]),
               _blk([64, 32], [128, 96, 64]), _blk([64, 32], [128, 96, 64], [64, 64, 0, 0]),
               np.array([128, 112, 100, 80]),
               _blk([96, 48], [256, 192, 128]), _blk([96, 48], [256, 192, 128]), _blk([96, 48], [256, 192, 128], [128, 128, 0, 0]),
               np.array([256, 224, 200, 160]),
               _blk([256, 192, 128], [512, 384, 256]), _blk([256, 192, 128], [512, 384, 256], [256, 0]),
               None]

VIT_RES_TINY = ((4, 192),) + ((1, (192, 3, 64), (192, 768), 1),) * 4 + ((3, 192, 384),) + \
    ((1, (384, 6, 64), (384, 1536), 1),) * 4 + ((3, 384, 768),) + \
    ((1, (768, 12, 64), (768, 3072), 1),) * 4 + ((2, 768, 1000),)

CASES = {
    # name: net, supernet?, batch, example_per_arch, warm-up epochs, epoch, RNG seed before forward
    'small_multi':   dict(net='small', supernet=True, batch=8, epa=2, warmup=0, epoch=0, seed=11),
    'small_multi2':  dict(net='small', supernet=True, batch=8, epa=4, warmup=0, epoch=0, seed=12, xseed=77),
    'small_single':  dict(net='small', supernet=True, batch=8, epa=2, warmup=0, epoch=3, seed=30007, single=True),
    'small_hybrid':  dict(net='small', supernet=True, batch=8, epa=2, warmup=0, epoch=1, seed=10002, hybrid=True),
    'small_warmup':  dict(net='small', supernet=True, batch=8, epa=2, warmup=5, epoch=2, seed=13),
    'small_eval':    dict(net='small', supernet=True, batch=4, epa=2, warmup=0, epoch=0, seed=14, train=False),
    'small_dense':   dict(net='small', supernet=False, batch=4, epa=None, warmup=0, epoch=0, seed=15),
    # BASELINE.json configs[0]: ViT-Res-Tiny reference net, forward + loss on CPU, batch 2
    'vit_res_tiny_b2': dict(net='vit_res_tiny', supernet=False, batch=2, epa=None, warmup=0, epoch=0, seed=16),
}


# Candidate evaluation (BASELINE configs[4]): a super-network definition in which every block exists, and dense
# sub-network definitions in the form search_utils/gen_utils.py produces (same entry count, prefix extents, exists flags).
EVO_SUPER_DEF = tuple((1,) + d[1:3] + (1,) if d[0] == 1 else d for d in SMALL_DEF)


def _sub(c1, c2, c3, blocks):
    """blocks: per transformer block (heads, mlp_features, exists)."""
    out, width, j = [(4, c1)], c1, 0
    for d in EVO_SUPER_DEF[1:]:
        if d[0] == 1:
            h, f, e = blocks[j]
            j += 1
            out.append((1, (width, h, d[1][2]), (width, f), e))
        elif d[0] == 3:
            nxt = c2 if d[2] == 128 else c3
            out.append((3, width, nxt))
            width = nxt
        else:
            out.append((2, width, 1000))
    return tuple(out)


EVO_CANDIDATES = {
    'largest': EVO_SUPER_DEF,
    'narrow':  _sub(56, 112, 224, [(2, 96, 1), (1, 64, 1), (2, 256, 1), (1, 192, 1), (2, 128, 0), (3, 384, 1), (4, 512, 1)]),
    'odd':     _sub(44, 100, 200, [(1, 128, 1), (2, 128, 0), (1, 128, 1), (2, 256, 0), (2, 256, 0), (2, 256, 1), (3, 384, 0)]),
    'mixed':   _sub(64, 80, 256, [(2, 64, 1), (2, 96, 1), (2, 192, 1), (1, 256, 1), (1, 128, 1), (4, 256, 1), (2, 512, 1)]),
}


# Patch-16 super-network with a distillation token (nets/vision_transformer_supernet.py): N = 196 + 2 tokens,
# head dims 32 and 64 (both attention kernels), a droppable block and a BypassBlock.
VIT16_DEF = ((0, 64),
             (1, (64, 2, 32), (64, 128), 1), (1, (64, 1, 64), (64, 192), 1), (1, (64, 2, 32), (64, 128), 0), (1, (64, 2, 64), (64, 128), 1),
             (2, 64, 1000))
VIT16_SPACE = [np.array([64, 56, 40]),
               _blk([64, 32], [128, 96, 64]), _blk([64], [192, 128], [64, 64, 0]), _blk([64, 32], [128, 64]), _blk([128, 64], [128, 96]),
               None]
VIT16_CASES = {
    'vit16_multi': dict(supernet=True, batch=8, epa=2, warmup=0, epoch=0, seed=21),
    'vit16_single': dict(supernet=True, batch=4, epa=2, warmup=0, epoch=2, seed=22, single=True),
    'vit16_eval': dict(supernet=True, batch=4, epa=2, warmup=0, epoch=0, seed=23, train=False),
    'vit16_dense': dict(supernet=False, batch=4, epa=None, warmup=0, epoch=0, seed=24),
}


# BASELINE-size parity cases (oracle/make_golden_baseline.py): the real search spaces of BASELINE.json configs[1] / [2] and of the
# published Tiny recipe, at a batch the CPU reference finishes in seconds.  `single` = one architecture per step (engine.py:121-122
# seeds the draw per iteration), `multi` = example_per_arch 2 -> four architectures in a batch of 8.
BASELINE_CASES = {
    'sr_tiny_single':    dict(space='sr_tiny', supernet=True, batch=8, epa=2, warmup=0, epoch=2, seed=20003, single=True),
    'sr_tiny_multi':     dict(space='sr_tiny', supernet=True, batch=8, epa=2, warmup=0, epoch=0, seed=41),
    'sr_tiny_mh_single': dict(space='sr_tiny_mh', supernet=True, batch=8, epa=2, warmup=0, epoch=1, seed=10005, single=True),
    'sr_tiny_mh_multi':  dict(space='sr_tiny_mh', supernet=True, batch=8, epa=2, warmup=0, epoch=0, seed=42),
    'sr_small_single':   dict(space='sr_small', supernet=True, batch=8, epa=2, warmup=0, epoch=0, seed=7, single=True),
    'sr_small_multi':    dict(space='sr_small', supernet=True, batch=8, epa=2, warmup=0, epoch=0, seed=43),
}


def baseline_net(space):
    """(network_def, num_channels_to_keep) of a named search space (tables restated in vit_search_b200/supernet_config)."""
    from vit_search_b200 import supernet_config as sc
    return sc.network_def(space), sc.num_channels_to_keep(space)


def probe_vectors(shape):
    """Seeded +-1 vectors (l [rows], r [cols]) for the gradient projections stored in the BASELINE-size goldens."""
    import torch
    g = torch.Generator().manual_seed(1000003 * shape[0] + shape[1])
    lv = torch.randint(0, 2, (shape[0],), generator=g).float() * 2 - 1
    rv = torch.randint(0, 2, (shape[1],), generator=g).float() * 2 - 1
    return lv, rv


def element_keys(network_def, full):
    """Parameters whose gradients are stored with element-level evidence: the four Linear weights of the first transformer block of
    every stage, both SR convolutions / token Linears and both heads.  Returns (stored in full, stored as sample + projections)."""
    first, stage_start = [], True
    bi = 0
    for d in network_def[1:]:
        if d[0] == 1:
            if stage_start:
                first.append(bi)
                stage_start = False
            bi += 1
        elif d[0] == 3:
            bi += 1
            stage_start = True
    names = []
    for j, b in enumerate(first):
        names.append([('blocks.%d.%s.weight' % (b, n)) for n in ('attn.qkv', 'attn.proj', 'mlp.fc1', 'mlp.fc2')])
    full = set(names[0]) if full else set()
    full_set = full
    sampled = set(k for grp in names for k in grp) - full_set
    sampled |= {'cls_head.weight', 'patch_head.weight'}
    for i, d in enumerate(network_def[1:]):
        if d[0] == 3:
            sampled |= {'blocks.%d.patch_reduce.weight' % i, 'blocks.%d.token_transform.weight' % i}
    return full, sampled
