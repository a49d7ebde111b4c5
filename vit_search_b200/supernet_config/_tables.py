"""Search-space tables for the ViT-Res super-networks (input DATA of the hot path).

Same content and the same access pattern as the reference's ``supernet_config`` package
(``getattr(supernet_config, args.search_space).num_channels_to_keep``, main.py:344-345): a list with
one entry per ``network_def`` item -- an ``np.ndarray`` of embedding widths for the patch-embed / SR
entries, a ``{'attn','mlp','layer'}`` dict for transformer blocks, ``None`` for the head.
Values transcribed from supernet_config/{sr_tiny,sr_tiny_mh,sr_tiny_666,sr_small,sr_small_mh}.py;
here they are generated from one compact description per space instead of repeated literals.

Each stage: (C, heads, head_dim, hidden, pattern, embed choices, attn choices, mlp choices,
number of all-or-nothing 'layer' zeros out of 4).  Pattern letters: B = always-on block,
S = block that can be skipped.
"""
import numpy as np


def _ladder(top, step, n):
    return [top - i * step for i in range(n)]


SPACES = {
    # supernet_config/sr_tiny.py:34-71 (7/7/4 blocks, 64-wide heads)
    'sr_tiny': dict(embed_type=4, stages=[
        (256, 4, 64, 768, 'BSBSBSB', [256, 224, 192, 176, 160], [256, 192, 128], _ladder(768, 128, 4), 1),
        (512, 8, 64, 1536, 'BSBSBSB', [512, 448, 384, 352, 320], [512, 384, 256], _ladder(1536, 256, 4), 1),
        (1024, 12, 64, 3072, 'BBBB', [1024, 896, 768, 704, 640], [768, 640, 512], _ladder(3072, 512, 4), 1)]),
    # supernet_config/sr_tiny_mh.py:34-66 (6/6/6 blocks, 32/48/64-wide heads) -- scripts/vit-sr-nas/super_net/tiny.sh:19-20
    'sr_tiny_mh': dict(embed_type=4, stages=[
        (256, 6, 32, 768, 'BSBSBS', [256, 224, 192, 176, 160], _ladder(192, 32, 4), _ladder(768, 64, 7), 2),
        (512, 12, 48, 1536, 'BSBSBS', [512, 448, 384, 352, 320], _ladder(576, 96, 4), _ladder(1536, 128, 7), 2),
        (1024, 12, 64, 3072, 'BSBSBS', [1024, 896, 768, 704, 640], _ladder(768, 128, 4), _ladder(3072, 256, 7), 2)]),
    # supernet_config/sr_tiny_666.py
    'sr_tiny_666': dict(embed_type=4, stages=[
        (256, 4, 64, 768, 'BSBSBS', [256, 224, 192, 176, 160], [256, 192, 128], _ladder(768, 64, 7), 2),
        (512, 8, 64, 1536, 'BSBSBS', [512, 448, 384, 352, 320], [512, 384, 256], _ladder(1536, 128, 7), 2),
        (1024, 12, 64, 3072, 'BSBSBS', [1024, 896, 768, 704, 640], _ladder(768, 128, 4), _ladder(3072, 256, 7), 2)]),
    # supernet_config/sr_small.py:37-72 -- scripts/.../super_net/no_distill/small_flexible-conv-patch.sh:19
    'sr_small': dict(embed_type=5, mid=32, stages=[
        (320, 8, 32, 960, 'BSBSBSB', [320, 280, 240, 220, 200], _ladder(256, 32, 4), _ladder(960, 80, 7), 2),
        (640, 12, 48, 1920, 'BSBSBSB', [640, 560, 480, 440, 400], _ladder(576, 96, 4), _ladder(1920, 160, 7), 2),
        (1280, 12, 64, 3840, 'BSBSBSB', [1280, 1120, 960, 880, 800], _ladder(768, 128, 4), _ladder(3840, 320, 7), 2)]),
    # supernet_config/sr_small_mh.py:37-72 -- scripts/.../super_net/small.sh:19
    'sr_small_mh': dict(embed_type=4, stages=[
        (320, 8, 32, 960, 'BSBSBSB', [320, 280, 240, 220, 200], _ladder(256, 32, 4), _ladder(960, 80, 7), 2),
        (640, 16, 48, 1920, 'BSBSBSB', [640, 560, 480, 440, 400], _ladder(768, 96, 4), _ladder(1920, 160, 7), 2),
        (1280, 16, 64, 3840, 'BSBSBSB', [1280, 1120, 960, 880, 800], _ladder(1024, 128, 4), _ladder(3840, 320, 7), 2)]),
}


def num_channels_to_keep(space):
    out = []
    for (c, _h, _d, _f, pattern, embed, attn, mlp, nzero) in SPACES[space]['stages']:
        out.append(np.array(embed))
        for letter in pattern:
            layer = np.array([c] * (4 - nzero) + [0] * nzero) if letter == 'S' else None
            out.append({'attn': np.array(attn), 'mlp': np.array(mlp), 'layer': layer})
    out.append(None)
    return out


def network_def(space, num_classes=1000):
    """The largest network of a space (what the super_net launch scripts pass as --network-def)."""
    sp = SPACES[space]
    stages = sp['stages']
    c0 = stages[0][0]
    nd = [(5, c0, sp['mid']) if sp['embed_type'] == 5 else (sp['embed_type'], c0)]
    for si, (c, h, d, f, pattern, *_rest) in enumerate(stages):
        if si > 0:
            nd.append((3, stages[si - 1][0], c))
        nd.extend((1, (c, h, d), (c, f), 1) for _ in pattern)
    nd.append((2, stages[-1][0], num_classes))
    return tuple(nd)


# Dense networks named by BASELINE.json configs 1 and 4.
VIT_RES_TINY = ((4, 192),) + ((1, (192, 3, 64), (192, 768), 1),) * 4 + ((3, 192, 384),) + \
    ((1, (384, 6, 64), (384, 1536), 1),) * 4 + ((3, 384, 768),) + \
    ((1, (768, 12, 64), (768, 3072), 1),) * 4 + ((2, 768, 1000),)   # scripts/vit-sr-nas/reference_net/tiny.sh:18

# BASELINE configs[3]: the searched ViT-ResNAS-Medium network, 4.6 G MACs (scripts/vit-sr-nas/searched_net/medium_mac@4.6G.sh:18)
VIT_RESNAS_MEDIUM = (
    (4, 240), (1, (240, 7, 32), (240, 960), 1), (1, (240, 6, 32), (240, 960), 1), (1, (240, 7, 32), (240, 800), 1),
    (1, (240, 8, 32), (240, 960), 1), (1, (240, 7, 32), (240, 880), 1), (1, (240, 8, 32), (240, 880), 1), (1, (240, 6, 32), (240, 800), 1),
    (3, 240, 640), (1, (640, 10, 48), (640, 1120), 1), (1, (640, 14, 48), (640, 1760), 1), (1, (640, 14, 48), (640, 1920), 1),
    (1, (640, 16, 48), (640, 1760), 1), (1, (640, 14, 48), (640, 1440), 1), (1, (640, 16, 48), (640, 1760), 1), (1, (640, 16, 48), (640, 1920), 1),
    (3, 640, 880), (1, (880, 16, 64), (880, 3200), 1), (1, (880, 10, 64), (880, 3840), 1), (1, (880, 16, 64), (880, 3840), 1),
    (1, (880, 12, 64), (880, 3200), 1), (1, (880, 16, 64), (880, 3520), 1), (1, (880, 14, 64), (880, 3520), 1), (2, 880, 1000))
