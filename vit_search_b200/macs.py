"""Algorithmic multiply-accumulate counts for a ViT-Res ``network_def``.

Same counting convention as the reference's ``ComputationEstimator`` with ``return_mac=True``
(network_utils/compute_flop_mac.py:227-307): only the dense contractions count -- stem convs, qkv /
QK^T / PV / proj, fc1 / fc2, the SR conv and token Linear, one classifier head.  LayerNorm, softmax,
GELU, biases, position embeddings and the patch head are free.  ``bench.py`` uses it for the roofline's
algorithmic FLOPs (train step = 6 x MAC, forward = 2 x MAC); the reference's printed known answers
(1.7944e9 for ViT-Res-Tiny, compute_flop_mac.py:411-429) are checked in tests/test_macs.py.
"""


def attention_macs(c, heads, head_dim, n):
    hd = heads * head_dim
    return n * c * 3 * hd + 2 * n * n * hd + n * hd * c          # compute_flop_mac.py:53-74


def mlp_macs(c, hidden, n):
    return 2 * n * c * hidden                                      # compute_flop_mac.py:77-93


def stem_macs(c, num_patches, embed_type, mid=24, patch=14, in_chans=3):
    if embed_type == 0:
        return c * in_chans * patch * patch * num_patches
    r = 112 * 112                                                  # compute_flop_mac.py:131-143
    half = patch // 2
    return in_chans * mid * 9 * r + 2 * mid * mid * 9 * r + c * mid * half * half * num_patches


def sr_macs(grid, c_in, c_out, tokens=1):
    g = grid // 2
    return g * g * c_out * 9 * c_in + tokens * c_in * c_out        # compute_flop_mac.py:169-194


def network_macs(network_def, resolution=224, patch=14, tokens=1):
    grid = resolution // patch
    n = grid * grid + tokens
    d0 = network_def[0]
    c = d0[1]
    total = stem_macs(c, grid * grid, d0[0], d0[2] if d0[0] == 5 else 24, patch)
    for d in network_def[1:]:
        if d[0] == 1 and d[3]:
            total += attention_macs(d[1][0], d[1][1], d[1][2], n) + mlp_macs(d[2][0], d[2][1], n)
        elif d[0] == 3:
            total += sr_macs(grid, d[1], d[2], tokens)
            grid //= 2
            n = grid * grid + tokens
            c = d[2]
        elif d[0] == 2:
            total += tokens * c * d[2]
    return total


def effective_network_def(network_def, keeps, sample):
    """The dense sub-network that sample `sample` of a supernet batch actually evaluates, given the
    per-entry keep dicts of one step (see nets.vit_sr_supernet.sample_keeps).  Masked-away heads,
    hidden channels, embedding channels and dropped blocks earn no credit."""
    out = []
    c = None
    for d, k in zip(network_def, keeps):
        if d[0] in (0, 4, 5):
            c = k['embed'][sample] if k else d[1]
            out.append((d[0], c) + tuple(d[2:]))
        elif d[0] == 1:
            on = bool(d[3]) and (not k or k.get('layer') is None or k['layer'][sample] > 0)
            heads = (k['attn'][sample] // d[1][2]) if k and 'attn' in k else d[1][1]
            hidden = k['mlp'][sample] if k and 'mlp' in k else d[2][1]
            out.append((1, (c, heads, d[1][2]), (c, hidden), 1 if on else 0))
        elif d[0] == 3:
            c2 = k['embed'][sample] if k else d[2]
            out.append((3, c, c2))
            c = c2
        else:
            out.append((2, c, d[2]))
    return tuple(out)
