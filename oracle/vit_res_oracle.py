"""CPU ORACLE (test infrastructure, NOT product code).

A functional restatement, in plain torch on the CPU, of the reference's ViT-Res
super-network training hot path.  Every function cites the reference lines
(relative to /root/reference) whose arithmetic it restates.  Nothing under
``vit_search_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs do, and only as the checker / the CPU baseline.

Pinning: the reference ships no golden vectors for this path, so
the oracle is pinned against the reference ITSELF: ``oracle/make_golden.py``
imports ``/root/reference/nets/*.py`` through the timm shim in
``oracle/ref_shim.py``, runs both on identical seeded inputs / weights / RNG
state, asserts agreement and writes ``tests/golden/*.npz``.  The CPU test-suite
re-checks this oracle against those committed vectors.

All state lives in a flat ``{name: tensor}`` dict with the reference's
``state_dict`` keys, so weights move freely between reference, oracle and the
CUDA modules.  dtype follows the parameters (fp32 or fp64).
"""
import math
import zlib

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-6  # nets/masked_layer_norm.py:100
BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, nets/patch_conv.py:28
BN_MOMENTUM = 0.1

T_EMBED, T_TRANS, T_HEAD, T_SR, T_CONV_EMBED, T_FLEX_CONV_EMBED = 0, 1, 2, 3, 4, 5  # nets/vit_sr_supernet.py:26-32


# --------------------------------------------------------------------------------------
# sub-architecture sampling (ChannelDrop)
# --------------------------------------------------------------------------------------
def warmup_num_choices(n_choices, epoch, num_warmup_epochs):
    """Number of (largest) width choices live at `epoch`.  nets/channel_drop.py:131-141."""
    if num_warmup_epochs == 0:
        return n_choices
    n = min(1 + math.floor(epoch * (n_choices - 1) / num_warmup_epochs), n_choices)
    return max(n, 1)


def keep_table(choices, batch, example_per_arch, single_arch, epoch, num_warmup_epochs):
    """Per-row keep counts of the prefix-mask table.  nets/channel_drop.py:33 (descending sort),
    :143-157 (rows cycle through the live choices)."""
    ch = sorted((int(c) for c in choices), reverse=True)
    nconf = warmup_num_choices(len(ch), epoch, num_warmup_epochs)
    ncycles = 1 if single_arch else math.ceil((batch // example_per_arch) / nconf)
    return [ch[r % nconf] for r in range(nconf * ncycles)]


def draw_keep(table, batch, example_per_arch, single_arch):
    """One ChannelDrop.forward_mask draw -> per-sample keep counts (list of int, len=batch).
    Consumes exactly one ``torch.randperm(len(table))`` from the global CPU generator.
    nets/channel_drop.py:93-111: rows permuted; multi-arch: first G=B//epa rows tiled epa times
    (sample i -> row perm[i % G]); single-arch: row perm[0] for every sample."""
    perm = torch.randperm(len(table)).tolist()
    if single_arch:
        return [table[perm[0]]] * batch
    g = batch // example_per_arch
    return [table[perm[i % g]] for i in range(batch)]


def prefix_mask(keep, width, dtype=torch.bool):
    """[B,1,C] prefix mask from per-sample keep counts.  nets/channel_drop.py:153-157."""
    k = torch.as_tensor(keep, dtype=torch.long).view(-1, 1, 1)
    return (torch.arange(width).view(1, 1, -1) < k).to(dtype)


# --------------------------------------------------------------------------------------
# masked layer norm
# --------------------------------------------------------------------------------------
def masked_layer_norm(x, weight, bias, keep, eps=LN_EPS):
    """nets/masked_layer_norm.py:23-50 followed by the re-mask at :124.
    x[B,N,C] is zero beyond each sample's prefix; statistics are plain means over C rescaled by
    1/p with p = keep/C.  keep=None -> F.layer_norm (:119-122)."""
    if keep is None:
        return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)
    c = x.shape[-1]
    m = prefix_mask(keep, c, x.dtype)
    inv_p = 1.0 / m.mean(dim=2, keepdim=True)
    mu = x.mean(dim=2, keepdim=True) * inv_p
    ex2 = (x * x).mean(dim=2, keepdim=True) * inv_p
    inv_std = 1.0 / (ex2 - mu * mu + eps).sqrt()
    z = (x - mu) * inv_std
    return (weight.view(1, 1, c) * z + bias.view(1, 1, c)) * m


def masked_layer_norm_backward(grad_out, x, weight, keep, eps=LN_EPS):
    """Explicit restatement of MaskedLayerNormFunc.backward, nets/masked_layer_norm.py:55-88.
    grad_out is the gradient w.r.t. the *unmasked* y (i.e. already multiplied by the mask).
    Returns (gx, g_gamma, g_beta)."""
    c = x.shape[-1]
    m = prefix_mask(keep, c, x.dtype)
    inv_p = 1.0 / m.mean(dim=2, keepdim=True)
    mu = x.mean(dim=2, keepdim=True) * inv_p
    ex2 = (x * x).mean(dim=2, keepdim=True) * inv_p
    inv_std = 1.0 / (ex2 - mu * mu + eps).sqrt()
    z = (x - mu) * inv_std
    dz = grad_out * weight.view(1, 1, c)
    mean_dz = dz.mean(dim=2, keepdim=True)
    mean_zdz = (z * dz).mean(dim=2, keepdim=True)
    gx = (dz - (mean_dz + z * mean_zdz) * inv_p) * inv_std
    return gx, (grad_out * z).sum(dim=(0, 1)), grad_out.sum(dim=(0, 1))


# --------------------------------------------------------------------------------------
# attention / mlp / block
# --------------------------------------------------------------------------------------
def attention(x, p, pre, num_heads, head_dim, keep_hd=None):
    """nets/supernet_blocks.py:100-120.  qkv features ordered (3,H,D); scale = D^-0.5 (:85);
    the head ChannelDrop zeroes a prefix-complement of the concatenated head outputs (:111-112)
    before proj."""
    b, n, _ = x.shape
    qkv = F.linear(x, p[pre + 'qkv.weight'], p[pre + 'qkv.bias'])
    qkv = qkv.reshape(b, n, 3, num_heads, head_dim).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = (q @ k.transpose(-2, -1)) * (head_dim ** -0.5)
    a = a.softmax(dim=-1)
    o = (a @ v).transpose(1, 2).reshape(b, n, num_heads * head_dim)
    if keep_hd is not None:
        o = o * prefix_mask(keep_hd, num_heads * head_dim, o.dtype)
    return F.linear(o, p[pre + 'proj.weight'], p[pre + 'proj.bias'])


def mlp(x, p, pre, keep_hidden=None):
    """nets/supernet_blocks.py:37-52: fc1 -> exact (erf) GELU -> hidden prefix mask -> fc2."""
    h = F.gelu(F.linear(x, p[pre + 'fc1.weight'], p[pre + 'fc1.bias']))
    if keep_hidden is not None:
        h = h * prefix_mask(keep_hidden, h.shape[-1], h.dtype)
    return F.linear(h, p[pre + 'fc2.weight'], p[pre + 'fc2.bias'])


def drop_path_scale(f, dp_keep, drop_prob):
    """nets/drop.py:11-26 with the Bernoulli draw made explicit: dp_keep[b] in {0,1}."""
    if dp_keep is None or drop_prob == 0.0:
        return f
    s = torch.as_tensor(dp_keep, dtype=f.dtype).view(-1, 1, 1) / (1.0 - drop_prob)
    return f * s


def block(x, p, pre, num_heads, head_dim, embed_keep, layer_keep_in, keeps, dp=None, drop_prob=0.0):
    """nets/supernet_blocks.py:209-255.  `keeps` = dict(attn=..., layer=..., mlp=...) of per-sample keep
    lists (or None).  Masks are prefix masks, so `a & b` is min(keep_a, keep_b).
    Returns (x, embed_keep, current_layer_keep)."""
    c = x.shape[-1]
    f = masked_layer_norm(x, p[pre + 'norm1.weight'], p[pre + 'norm1.bias'], embed_keep)
    f = attention(f, p, pre + 'attn.', num_heads, head_dim, keeps.get('attn'))
    f = drop_path_scale(f, None if dp is None else dp[0], drop_prob)
    cur = None
    if keeps.get('layer') is not None:                      # :220-223
        cur = list(keeps['layer'])
        f = f * prefix_mask(cur, c, f.dtype)
        if layer_keep_in is not None:
            cur = [min(a, b) for a, b in zip(cur, layer_keep_in)]
    if embed_keep is not None:                              # :238-243
        cur = list(embed_keep) if cur is None else [min(a, b) for a, b in zip(cur, embed_keep)]
        f = f * prefix_mask(cur, c, f.dtype)
    x = x + f
    f = masked_layer_norm(x, p[pre + 'norm2.weight'], p[pre + 'norm2.bias'], embed_keep)
    f = mlp(f, p, pre + 'mlp.', keeps.get('mlp'))
    f = drop_path_scale(f, None if dp is None else dp[1], drop_prob)
    if cur is not None:                                     # :250-251
        f = f * prefix_mask(cur, c, f.dtype)
    return x + f, embed_keep, cur


# --------------------------------------------------------------------------------------
# patch-conv stem and spatial-reduction embedding
# --------------------------------------------------------------------------------------
def conv_bn_relu(x, p, pre, stride, training, new_stats):
    """nets/patch_conv.py:23-36: conv3x3(pad 1, no bias) -> BatchNorm2d -> ReLU.  In training the
    normalisation uses biased batch statistics and the running buffers move with momentum 0.1
    towards (mean, unbiased var)."""
    y = F.conv2d(x, p[pre + 'conv.weight'], None, stride=stride, padding=1)
    if training:
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        if new_stats is not None:
            n = y.numel() // y.shape[1]
            new_stats[pre + 'bn.running_mean'] = (1 - BN_MOMENTUM) * p[pre + 'bn.running_mean'] + BN_MOMENTUM * mean.detach()
            new_stats[pre + 'bn.running_var'] = (1 - BN_MOMENTUM) * p[pre + 'bn.running_var'] + BN_MOMENTUM * var.detach() * n / (n - 1)
            new_stats[pre + 'bn.num_batches_tracked'] = p[pre + 'bn.num_batches_tracked'] + 1
    else:
        mean, var = p[pre + 'bn.running_mean'], p[pre + 'bn.running_var']
    y = (y - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS)
    y = y * p[pre + 'bn.weight'].view(1, -1, 1, 1) + p[pre + 'bn.bias'].view(1, -1, 1, 1)
    return F.relu(y)


def patch_conv_embed(x, p, pre, training, new_stats=None):
    """nets/patch_conv.py:63-74: conv1(s2) -> [conv2 -> conv3] + residual -> conv_proj 7x7 s7 -> [B,256,C]."""
    y = conv_bn_relu(x, p, pre + 'conv1.', 2, training, new_stats)
    r = y
    y = conv_bn_relu(y, p, pre + 'conv2.', 1, training, new_stats)
    y = conv_bn_relu(y, p, pre + 'conv3.', 1, training, new_stats)
    y = y + r
    k = p[pre + 'conv_proj.weight'].shape[-1]
    y = F.conv2d(y, p[pre + 'conv_proj.weight'], p[pre + 'conv_proj.bias'], stride=k)
    return y.flatten(2).transpose(1, 2)


def plain_patch_embed(x, p, pre):
    """timm 0.3.2 PatchEmbed (embed type 0): Conv2d(k=s=patch) -> flatten(2).transpose(1,2)."""
    w = p[pre + 'proj.weight']
    return F.conv2d(x, w, p[pre + 'proj.bias'], stride=w.shape[-1]).flatten(2).transpose(1, 2)


def sr_embed(x, p, pre, grid, num_tokens, embed_keep, new_keep):
    """SpatialReductionPatchEmbedding.forward, nets/vit_sr_supernet.py:114-172."""
    b, _, c = x.shape
    tok_res, patch_res = x[:, :num_tokens], x[:, num_tokens:]
    xn = masked_layer_norm(x, p[pre + 'norm.weight'], p[pre + 'norm.bias'], embed_keep)
    patch_res = patch_res.transpose(1, 2).reshape(b, c, grid, grid)
    patch_res = F.avg_pool2d(patch_res, 2, 2).flatten(2).transpose(1, 2)          # :131-136
    pe = xn[:, num_tokens:].transpose(1, 2).reshape(b, c, grid, grid)
    pe = F.conv2d(pe, p[pre + 'patch_reduce.weight'], p[pre + 'patch_reduce.bias'], stride=2, padding=1)
    pe = pe.flatten(2).transpose(1, 2) + p[pre + 'pos_embed']                      # :139-144
    tok = F.linear(xn[:, :num_tokens], p[pre + 'token_transform.weight'], p[pre + 'token_transform.bias'])
    res = torch.cat((tok_res, patch_res), dim=1)
    c_out = tok.shape[-1]
    res = torch.cat((res, res.new_zeros(b, res.shape[1], c_out - c)), dim=2)       # :155-158
    y = torch.cat((tok, pe), dim=1) + res
    if new_keep is not None:
        y = y * prefix_mask(new_keep, c_out, y.dtype)                              # :165-166
    return y


# --------------------------------------------------------------------------------------
# whole model
# --------------------------------------------------------------------------------------
class Sampler:
    """Holds the lazily built ChannelDrop tables of one model and reproduces the reference's
    per-forward draw order: embed drop; per block attn, layer, mlp; SR drop at the
    end of the SR block."""

    def __init__(self, network_def, num_channels_to_keep, example_per_arch, num_warmup_epochs,
                 single_arch=False, hybrid_arch=False):
        self.network_def = network_def
        self.space = num_channels_to_keep
        self.epa = example_per_arch
        self.warm = num_warmup_epochs
        self.single = single_arch
        self.hybrid = hybrid_arch
        self.epoch = None
        self.tables = {}

    def set_epoch(self, epoch):                              # nets/channel_drop.py:160-162
        self.epoch = epoch
        self.tables = {}

    def _draw(self, key, choices, batch, single):
        if key not in self.tables:                           # lazy build, nets/channel_drop.py:77-79
            self.tables[key] = keep_table(choices, batch, self.epa, single, self.epoch, self.warm)
        return draw_keep(self.tables[key], batch, self.epa, single)

    def sample(self, batch):
        """-> list (one entry per network_def item) of keep-dicts, drawing in forward order."""
        out = []
        embed_single = self.single or self.hybrid            # nets/vit_sr_supernet.py:257-260,317-324
        for i, d in enumerate(self.network_def):
            sp = self.space[i]
            if i == 0:
                out.append({'embed': self._draw((i, 'embed'), sp, batch, embed_single)})
            elif d[0] == T_TRANS:
                k = {}
                if d[3]:
                    k['attn'] = self._draw((i, 'attn'), sp['attn'], batch, self.single)
                    if sp.get('layer') is not None:
                        k['layer'] = self._draw((i, 'layer'), sp['layer'], batch, self.single)
                    k['mlp'] = self._draw((i, 'mlp'), sp['mlp'], batch, self.single)
                out.append(k)
            elif d[0] == T_SR:
                out.append({'embed': self._draw((i, 'embed'), sp, batch, embed_single)})
            else:
                out.append({})
        return out


def block_names(network_def):
    """Index of each network_def entry inside model.blocks (BypassBlocks occupy a slot).
    nets/vit_sr_supernet.py:271-328."""
    names, j = {}, 0
    for i, d in enumerate(network_def):
        if d[0] in (T_TRANS, T_SR):
            names[i] = j
            j += 1
    return names


def forward(p, network_def, x, keeps=None, training=True, patch_output=True, num_tokens=1,
            drop_path_rate=0.0, dp_keeps=None, new_stats=None, eval_full_mask=False):
    """FlexibleDistillVisionTransformerSR.forward, nets/vit_sr_supernet.py:396-462.

    keeps: output of Sampler.sample (None for a non-supernet model).  In eval mode the reference's
    ChannelDrop emits an all-true mask (nets/channel_drop.py:84-88): pass eval_full_mask=True to take
    the masked-LN code path with full keeps, as the supernet does at eval time.
    dp_keeps: optional list (per transformer block) of (attn_keep[B], mlp_keep[B]) 0/1 draws.
    Returns cls_pred | (cls_pred, patch_pred) | (cls_pred, dst_pred) like the reference."""
    b = x.shape[0]
    d0 = network_def[0]
    if d0[0] in (T_CONV_EMBED, T_FLEX_CONV_EMBED):
        h = patch_conv_embed(x, p, 'patch_embed.', training, new_stats)
    else:
        h = plain_patch_embed(x, p, 'patch_embed.')
    h = torch.cat((p['tokens'].expand(b, -1, -1), h), dim=1) + p['pos_embed']      # :399-401
    grid = int(round(math.sqrt(h.shape[1] - num_tokens)))
    embed_keep = layer_keep = None
    width = d0[1]
    if keeps is not None:
        embed_keep = list(keeps[0]['embed'])
    elif eval_full_mask:
        embed_keep = [width] * b
    if embed_keep is not None:
        h = h * prefix_mask(embed_keep, width, h.dtype)                           # :406-407
    names = block_names(network_def)
    depth = sum(1 for d in network_def if d[0] == T_TRANS)
    dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]             # :267
    t = 0
    for i, d in enumerate(network_def):
        if d[0] == T_TRANS:
            if d[3]:
                k = keeps[i] if keeps is not None else {}
                if keeps is None and eval_full_mask:
                    k = {}
                h, embed_keep, layer_keep = block(
                    h, p, 'blocks.%d.' % names[i], d[1][1], d[1][2], embed_keep, layer_keep, k,
                    None if dp_keeps is None else dp_keeps[t], dpr[t])
            else:
                layer_keep = None                                                  # BypassBlock :54-56
            t += 1
        elif d[0] == T_SR:
            new_keep = None
            if keeps is not None:
                new_keep = list(keeps[i]['embed'])
            elif eval_full_mask:
                new_keep = [d[2]] * b
            h = sr_embed(h, p, 'blocks.%d.' % names[i], grid, num_tokens, embed_keep, new_keep)
            embed_keep, layer_keep, grid, width = new_keep, None, grid // 2, d[2]
    if training and patch_output:                                                  # :420-424
        h = masked_layer_norm(h, p['norm.weight'], p['norm.bias'], embed_keep)
        tok, patches = h[:, :num_tokens], h[:, num_tokens:]
    else:                                                                          # :426-428
        tok = masked_layer_norm(h[:, :num_tokens], p['norm.weight'], p['norm.bias'], embed_keep)
        patches = None
    cls_pred = F.linear(tok[:, 0], p['cls_head.weight'], p['cls_head.bias'])
    if patch_output:
        if training:
            return cls_pred, F.linear(patches, p['patch_head.weight'], p['patch_head.bias'])
        return cls_pred
    if num_tokens == 2:
        return cls_pred, F.linear(tok[:, 1], p['dst_head.weight'], p['dst_head.bias'])
    return cls_pred


def soft_target_ce(logits, target):
    """timm 0.3.2 SoftTargetCrossEntropy (source absent from /root/reference; call site main.py:392-394):
    mean over all leading dims of sum_c(-t * log_softmax(x))."""
    return torch.sum(-target * F.log_softmax(logits, dim=-1), dim=-1).mean()


def train_loss(p, network_def, x, targets, patch_targets, keeps, **kw):
    """Loss of engine.train_one_epoch's patch-mixup branch, engine.py:152-157 ('seq' patch output)."""
    cls_pred, patch_pred = forward(p, network_def, x, keeps, training=True, patch_output=True, **kw)
    return soft_target_ce(cls_pred, targets) + soft_target_ce(patch_pred, patch_targets), cls_pred, patch_pred


def adamw_step(params, grads, state, lr, weight_decay, step, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.AdamW as configured by timm create_optimizer (main.py:385): decoupled decay on
    parameters with ndim > 1 that are not biases and not in no_weight_decay() ('tokens',
    nets/vit_sr_supernet.py:379-385)."""
    b1, b2 = betas
    for name, w in params.items():
        g = grads[name]
        wd = 0.0 if (w.ndim <= 1 or name.endswith('.bias') or name == 'tokens') else weight_decay
        m, v = state.setdefault(name, (torch.zeros_like(w), torch.zeros_like(w)))
        w.mul_(1 - lr * wd)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** step)).add_(eps)
        w.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))


# --------------------------------------------------------------------------------------
# parameter construction helpers (shapes follow the reference constructors)
# --------------------------------------------------------------------------------------
def param_shapes(network_def, num_tokens=1, patch_output=True, img_size=224, patch_size=14, num_classes=1000):
    """state_dict key -> shape, in the reference's registration order."""
    s = {}
    d0 = network_def[0]
    c = d0[1]
    npatch = (img_size // patch_size) ** 2
    s['tokens'] = (1, num_tokens, c)
    s['pos_embed'] = (1, npatch + num_tokens, c)
    if d0[0] in (T_CONV_EMBED, T_FLEX_CONV_EMBED):
        mid = d0[2] if d0[0] == T_FLEX_CONV_EMBED else 24
        cin = 3
        for n in ('conv1', 'conv2', 'conv3'):
            s['patch_embed.%s.conv.weight' % n] = (mid, cin, 3, 3)
            s['patch_embed.%s.bn.weight' % n] = (mid,)
            s['patch_embed.%s.bn.bias' % n] = (mid,)
            s['patch_embed.%s.bn.running_mean' % n] = (mid,)
            s['patch_embed.%s.bn.running_var' % n] = (mid,)
            s['patch_embed.%s.bn.num_batches_tracked' % n] = ()
            cin = mid
        s['patch_embed.conv_proj.weight'] = (c, mid, patch_size // 2, patch_size // 2)
        s['patch_embed.conv_proj.bias'] = (c,)
    else:
        s['patch_embed.proj.weight'] = (c, 3, patch_size, patch_size)
        s['patch_embed.proj.bias'] = (c,)
    names = block_names(network_def)
    grid = img_size // patch_size
    for i, d in enumerate(network_def):
        pre = 'blocks.%d.' % names[i] if i in names else None
        if d[0] == T_TRANS and d[3]:
            _, hh, dd = d[1]
            f = d[2][1]
            s[pre + 'norm1.weight'] = (c,)
            s[pre + 'norm1.bias'] = (c,)
            s[pre + 'attn.qkv.weight'] = (3 * hh * dd, c)
            s[pre + 'attn.qkv.bias'] = (3 * hh * dd,)
            s[pre + 'attn.proj.weight'] = (c, hh * dd)
            s[pre + 'attn.proj.bias'] = (c,)
            s[pre + 'norm2.weight'] = (c,)
            s[pre + 'norm2.bias'] = (c,)
            s[pre + 'mlp.fc1.weight'] = (f, c)
            s[pre + 'mlp.fc1.bias'] = (f,)
            s[pre + 'mlp.fc2.weight'] = (c, f)
            s[pre + 'mlp.fc2.bias'] = (c,)
        elif d[0] == T_SR:
            c2 = d[2]
            grid //= 2
            s[pre + 'pos_embed'] = (1, grid * grid, c2)
            s[pre + 'norm.weight'] = (c,)
            s[pre + 'norm.bias'] = (c,)
            s[pre + 'patch_reduce.weight'] = (c2, c, 3, 3)
            s[pre + 'patch_reduce.bias'] = (c2,)
            s[pre + 'token_transform.weight'] = (c2, c)
            s[pre + 'token_transform.bias'] = (c2,)
            c = c2
    s['norm.weight'] = (c,)
    s['norm.bias'] = (c,)
    s['cls_head.weight'] = (num_classes, c)
    s['cls_head.bias'] = (num_classes,)
    if num_tokens == 2:
        s['dst_head.weight'] = (num_classes, c)
        s['dst_head.bias'] = (num_classes,)
    if patch_output:
        s['patch_head.weight'] = (num_classes, c)
        s['patch_head.bias'] = (num_classes,)
    return s


def keyed_fill(shapes, seed=0, dtype=torch.float32, running_stats=False):
    """Deterministic, key-addressed synthetic weights (order independent, so the reference, the oracle
    and the CUDA modules can all be filled identically via their state_dict keys).  Scales imitate the
    reference init (trunc-normal 0.02 Linear weights, LN gamma ~ 1) but biases/betas are non-zero so
    every bias path is exercised."""
    out = {}
    for k, shp in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + seed) & 0x7FFFFFFF)
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros((), dtype=torch.long)
            continue
        r = torch.randn(tuple(shp), generator=g, dtype=torch.float64)
        if k.endswith('running_mean'):
            t = 0.1 * r if running_stats else torch.zeros_like(r)           # running_stats: a trained net's buffers (eval-mode cases)
        elif k.endswith('running_var'):
            t = 1.0 + 0.3 * r.abs() if running_stats else torch.ones_like(r)
        elif k.endswith('norm.weight') or k.endswith('norm1.weight') or k.endswith('norm2.weight') or k.endswith('bn.weight'):
            t = 1.0 + 0.1 * r
        elif k.endswith('conv.weight'):
            t = r * (1.0 / math.sqrt(shp[1] * shp[2] * shp[3]))
        elif k.endswith('conv_proj.weight') or k.endswith('patch_reduce.weight') or k.endswith('proj.weight') and len(shp) == 4:
            t = r * (1.0 / math.sqrt(shp[1] * shp[2] * shp[3]))
        elif len(shp) >= 2:
            t = r * 0.02
        else:
            t = r * 0.02
        out[k] = t.to(dtype)
    return out


def synthetic_batch(batch, seed=1234, num_classes=1000, num_patch_targets=16, dtype=torch.float32):
    """seeded N(0,1) images; smoothed one-hot soft targets, repeated per patch."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, 224, 224, generator=g)
    y = torch.randint(0, num_classes, (batch,), generator=g)
    t = torch.full((batch, num_classes), 0.1 / num_classes)
    t[torch.arange(batch), y] += 0.9
    return x.to(dtype), t.to(dtype), t.unsqueeze(1).repeat(1, num_patch_targets, 1).to(dtype)


def ema_update(ema_state, model_state, decay):
    """timm 0.3.2 ModelEmaV2._update (source absent from /root/reference; call sites main.py:357-363, engine.py:179-180): every
    state_dict entry, parameters and buffers alike, follows ema = decay * ema + (1 - decay) * model.  "parity unpinned": the
    reference ships no test for it; this restates the published formula."""
    for k, m in model_state.items():
        e = ema_state[k]
        ema_state[k] = (decay * e.double() + (1.0 - decay) * m.double()).to(e.dtype)
    return ema_state


# ------------------------------------------------------------------------------------------------ SwitchTokenMix (token_mixup.py)
def token_mix_draws(batch, patch_len):
    """The random draws of one SwitchTokenMix.__call__ (token_mixup.py:146-162), consumed in the reference's order from the global
    torch CPU generator and numpy's global RandomState: first half of the batch -> _patch_mixup_fn (:112-128: randperm, then
    _gen_random_bbox :75-99), second half -> _image_mixup_fn (:131-137: randperm, then beta(0.8, 0.8))."""
    import numpy as np
    n1 = batch // 2
    n2 = batch - n1
    perm1 = torch.randperm(n1)
    lam = np.random.beta(1., 1.)
    area = int(patch_len * patch_len * lam)
    max_length = min(patch_len, area)

    def my_randint(low, high, size=None):                    # token_mixup.py:32-35
        if low == high:
            high = low + 1
        return np.random.randint(low, high, size=size)
    cut_h = my_randint(1, max(1, max_length - 1))
    cut_w = area // cut_h
    if cut_w > patch_len:
        cut_w = patch_len
        cut_h = area // cut_w
    yl = my_randint(0, max(0, patch_len - cut_h), size=2)
    xl = my_randint(0, max(0, patch_len - cut_w), size=2)
    y0, x0 = int(yl[1]), int(xl[1])                          # :90-91: both boxes are forced to the second draw
    lam1 = 1 - (cut_h * cut_w + 0.0) / (patch_len * patch_len)
    perm2 = torch.randperm(n2)
    lam2 = np.random.beta(0.8, 0.8)
    return dict(perm1=perm1, box=(y0, y0 + int(cut_h), x0, x0 + int(cut_w)), lam1=float(lam1), perm2=perm2, lam2=float(lam2))


def switch_token_mix(samples, labels, draws, patch_len, num_classes=1000, smoothing=0.1):
    """SwitchTokenMix.__call__ for given draws: first half patch-level mix (a box of patches pasted from a permuted image, per-patch
    targets switched with it, image-level target mixed with lam = 1 - box area), second half image-level mixup.  Returns NEW tensors
    (the reference overwrites `samples` in place with the same values).  fp32 arithmetic in the reference's operation order."""
    B, C, H, W = samples.shape
    n1 = B // 2
    ps = H // patch_len
    off = smoothing / num_classes
    on = 1. - smoothing + off
    y = torch.full((B, num_classes), off, dtype=torch.float32)
    y.scatter_(1, labels.long().view(-1, 1), on)
    out = samples.clone()
    targets = torch.zeros(B, num_classes)
    ptargets = torch.zeros(B, patch_len * patch_len, num_classes)
    # patch half (:112-128)
    p1 = draws['perm1']
    y0, y1, x0, x1 = draws['box']
    s1 = samples[:n1]
    o1 = s1.clone()
    o1[:, :, ps * y0:ps * y1, ps * x0:ps * x1] = s1[p1][:, :, ps * y0:ps * y1, ps * x0:ps * x1]
    out[:n1] = o1
    yt = y[:n1]
    pt = yt.reshape(n1, 1, 1, -1).repeat(1, patch_len, patch_len, 1)
    pt[:, y0:y1, x0:x1, :] = pt[p1][:, y0:y1, x0:x1, :]
    ptargets[:n1] = pt.flatten(1, 2)
    targets[:n1] = yt * draws['lam1'] + yt[p1] * (1. - draws['lam1'])
    # image half (:131-143)
    p2 = draws['perm2']
    s2 = samples[n1:]
    lam = draws['lam2']
    out[n1:] = s2 * lam + s2[p2] * (1. - lam)
    y2 = y[n1:]
    t2 = y2 * lam + y2[p2] * (1. - lam)
    targets[n1:] = t2
    ptargets[n1:] = t2.reshape(B - n1, 1, -1).repeat(1, patch_len * patch_len, 1)
    return out, targets, ptargets


# --------------------------------------------------------------------------------------
# evolutionary-search candidate evaluation (BASELINE configs[4])
# --------------------------------------------------------------------------------------
def sub_state_dict(source, sub_shapes):
    """nets/net_utils.py:11-57 (get_qkv_subnet_state_dict + get_sub_state_dict): the dense sub-network's weights are prefix slices of
    the super-network's -- every dimension of every tensor is cut to the sub-network's extent; q, k and v rows are cut separately
    inside the stacked qkv tensor (:22-26); 4-D conv kernels keep their spatial extent (:48-51); 0-d entries (BatchNorm's
    num_batches_tracked) keep the freshly built sub-network's own value (:52-53), which is 0."""
    out = {}
    for k, shp in sub_shapes.items():
        src = source[k]
        shp = tuple(shp)
        if 'qkv' in k:
            n_sub, n_src = shp[0] // 3, src.shape[0] // 3
            t = torch.cat([src[0:n_sub], src[n_src:n_src + n_sub], src[2 * n_src:2 * n_src + n_sub]], dim=0)
            out[k] = t[:, :shp[1]] if len(shp) == 2 else t
        elif len(shp) == 0:
            out[k] = torch.zeros((), dtype=torch.long)
        elif len(shp) == 4:
            assert src.shape[2:] == shp[2:]
            out[k] = src[:shp[0], :shp[1]]
        else:
            out[k] = src[tuple(slice(0, n) for n in shp)]
    return out


def candidate_logits(p_super, sub_network_def, x, patch_output=True):
    """evo_search.py:256-273 + engine.py:201-212 for one batch: build the dense sub-network `sub_network_def`, load the prefix slices of
    the super-network state dict, eval-mode forward (BatchNorm running statistics, no ChannelDrop, no drop-path) -> class logits."""
    p_sub = sub_state_dict(p_super, param_shapes(sub_network_def, patch_output=patch_output))
    return forward(p_sub, sub_network_def, x, None, training=False, patch_output=patch_output)


def eval_metrics(logits, labels):
    """One batch of engine.evaluate (engine.py:195,222-228): mean hard-label cross entropy and timm's accuracy(topk=(1, 5)) in
    percent.  A sample counts for top-k when fewer than k logits are strictly larger than its target logit (no ties in practice)."""
    z = logits.double()
    loss = (torch.logsumexp(z, dim=1) - z.gather(1, labels.view(-1, 1)).squeeze(1)).mean()
    rank = (z > z.gather(1, labels.view(-1, 1))).sum(dim=1)
    return dict(loss=loss.item(), acc1=100.0 * (rank < 1).double().mean().item(), acc5=100.0 * (rank < 5).double().mean().item())


def evaluate_meters(batches):
    """The averaging of engine.evaluate over a loader (engine.py:230-233, utils.MetricLogger): `loss` is the mean of the per-batch mean
    losses (each update has n=1), acc1 / acc5 are weighted by batch size.  batches: list of (metrics dict, batch_size)."""
    nb = len(batches)
    n = sum(b for _, b in batches)
    return dict(loss=sum(m['loss'] for m, _ in batches) / nb, acc1=sum(m['acc1'] * b for m, b in batches) / n,
                acc5=sum(m['acc5'] * b for m, b in batches) / n)
