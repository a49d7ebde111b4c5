"""CPU: the oracle restatement against the REFERENCE's outputs (tests/golden/*.npz, written by oracle/make_golden.py
from the unmodified /root/reference modules).  This is what pins the oracle; it runs without /root/reference."""
import os

import numpy as np
import pytest
import torch

from oracle import vit_res_oracle as O
from oracle.cases import CASES, SMALL_DEF, SMALL_SPACE, VIT_RES_TINY

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
TOL = 2e-5


def rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_matches_reference_golden(name):
    case = CASES[name]
    G = np.load(os.path.join(GOLD, name + '.npz'))
    nd = VIT_RES_TINY if case['net'] == 'vit_res_tiny' else SMALL_DEF
    w = O.keyed_fill(O.param_shapes(nd), seed=case.get('wseed', 0))
    p = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in w.items()}
    B = case['batch']
    x, t, pt = O.synthetic_batch(B, seed=case.get('xseed', 1234))
    if not case.get('train', True):
        with torch.no_grad():
            cls = O.forward(p, nd, x, None, training=False, eval_full_mask=case['supernet'])
        assert rel(cls.numpy(), G['cls']) < TOL
        return
    keeps = None
    if case['supernet']:
        smp = O.Sampler(nd, SMALL_SPACE, case['epa'], case['warmup'], case.get('single', False), case.get('hybrid', False))
        smp.set_epoch(case['epoch'])
        torch.manual_seed(case['seed'])
        keeps = smp.sample(B)
        flat = [k[n] for k in keeps for n in ('embed', 'attn', 'layer', 'mlp') if n in k]
        assert flat == G['keeps'].tolist()
    stats = {}
    loss, cls, patch = O.train_loss(p, nd, x, t, pt, keeps, new_stats=stats)
    loss.backward()
    assert rel(cls.detach().numpy(), G['cls']) < TOL and rel(patch.detach().numpy(), G['patch']) < TOL
    assert abs(loss.item() - float(G['loss'])) < 1e-5
    for k in G.files:
        if k.startswith('g:'):
            g = p[k[2:]].grad
            ref = G[k]
            assert (rel(g.numpy(), ref) < TOL) if np.linalg.norm(ref) > 0 else (g.norm().item() < 1e-12), k
        elif k.startswith('gn:'):
            gn = float(G[k])
            n = p[k[3:]].grad.double().norm().item()
            assert abs(n - gn) <= TOL * max(gn, 1e-12) + 1e-12, k
        elif k.startswith('s:'):
            assert rel(stats[k[2:]].float().numpy(), G[k].astype(np.float32)) < TOL, k


def test_function_level_vectors():
    G = np.load(os.path.join(GOLD, 'functions.npz'))
    g = torch.Generator().manual_seed(7)
    B, N, C = 6, 5, 48
    keep = G['ln_keep'].tolist()
    x = torch.randn(B, N, C, generator=g) * O.prefix_mask(keep, C, torch.float32)
    wt = 1 + 0.1 * torch.randn(C, generator=g)
    bs = 0.1 * torch.randn(C, generator=g)
    go = torch.randn(B, N, C, generator=g)
    assert rel(O.masked_layer_norm(x, wt, bs, keep).numpy(), G['ln_y']) < TOL
    gx, gw, gb = O.masked_layer_norm_backward(go * O.prefix_mask(keep, C, torch.float32), x, wt, keep)
    assert rel(gx.numpy(), G['ln_gx']) < TOL and rel(gw.numpy(), G['ln_gw']) < TOL and rel(gb.numpy(), G['ln_gb']) < TOL
    # masked LN == plain LN on the kept slice
    y = O.masked_layer_norm(x, wt, bs, keep)
    for b, k in enumerate(keep):
        ref = torch.nn.functional.layer_norm(x[b, :, :k], (k,), wt[:k], bs[:k], 1e-6)
        assert rel(y[b, :, :k].numpy(), ref.numpy()) < 1e-5 and float(y[b, :, k:].abs().max() if k < C else 0) == 0
    # ChannelDrop tables / draws across warm-up epochs
    i = 0
    for epoch in (0, 2, 5, 9):
        for single in (False, True):
            table = O.keep_table([96, 64, 128, 32, 80], 12, 2, single, epoch, 5)
            torch.manual_seed(100 + epoch)
            assert O.draw_keep(table, 12, 2, single) == G['cd_draws'][i].tolist()
            i += 1
    f = torch.randn(8, 3, 4, generator=g)
    assert rel(O.drop_path_scale(f, G['dp_keep'].tolist(), 0.25).numpy(), G['dp_y']) < 1e-6


def test_oracle_fp64_noise_floor():
    """fp32 oracle vs its own fp64 evaluation: the noise floor the GPU tolerances sit above."""
    case = CASES['small_multi']
    w32 = O.keyed_fill(O.param_shapes(SMALL_DEF))
    w64 = {k: (v.double() if v.is_floating_point() else v) for k, v in w32.items()}
    x, t, pt = O.synthetic_batch(4)
    smp = O.Sampler(SMALL_DEF, SMALL_SPACE, 2, 0)
    smp.set_epoch(0)
    torch.manual_seed(1)
    keeps = smp.sample(4)
    with torch.no_grad():
        a = O.forward(w32, SMALL_DEF, x, keeps)[0]
        b = O.forward(w64, SMALL_DEF, x.double(), keeps)[0]
    assert rel(a.numpy(), b.numpy()) < 1e-5


def test_switch_token_mix_oracle_vs_reference_golden():
    """oracle.switch_token_mix / token_mix_draws against the REFERENCE's SwitchTokenMix outputs stored by oracle/make_golden_mixup.py
    (bit exact), including the RNG protocol: re-seeding and re-drawing must reproduce the stored permutations, box and lambdas."""
    import numpy as np
    import os
    import torch
    from oracle import vit_res_oracle as O
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'token_mix.npz'))
    ncase = len([k for k in z.files if k.endswith('_meta')])
    assert ncase >= 4
    for c in range(ncase):
        y0, y1, x0, x1, pl, seed = [int(v) for v in z['c%d_meta' % c]]
        samples, labels = torch.from_numpy(z['c%d_samples' % c]), torch.from_numpy(z['c%d_labels' % c])
        torch.manual_seed(seed)
        np.random.seed(seed)
        d = O.token_mix_draws(samples.shape[0], pl)
        assert d['box'] == (y0, y1, x0, x1)
        assert torch.equal(d['perm1'], torch.from_numpy(z['c%d_perm1' % c])) and torch.equal(d['perm2'], torch.from_numpy(z['c%d_perm2' % c]))
        assert [d['lam1'], d['lam2']] == list(z['c%d_lams' % c])
        out, t, pt = O.switch_token_mix(samples, labels, d, pl)
        assert torch.equal(out, torch.from_numpy(z['c%d_out' % c]))
        assert torch.equal(t, torch.from_numpy(z['c%d_targets' % c]))
        assert torch.equal(pt, torch.from_numpy(z['c%d_ptargets' % c]))


def test_candidate_evaluation_oracle_vs_reference_golden():
    """oracle.candidate_logits (prefix-sliced dense sub-network, eval forward) and eval_metrics against the REFERENCE's outputs stored by
    oracle/make_golden_evo.py (evo_search.py:256-273 + engine.py:195-228)."""
    import numpy as np
    import os
    import torch
    from oracle import vit_res_oracle as O
    from oracle.cases import EVO_SUPER_DEF, EVO_CANDIDATES
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'evo_eval.npz'))
    sup = O.keyed_fill(O.param_shapes(EVO_SUPER_DEF), seed=int(z['w_seed']), running_stats=True)
    x, _, _ = O.synthetic_batch(8, seed=int(z['x_seed']))
    for name, sub_def in EVO_CANDIDATES.items():
        ref = torch.from_numpy(z[name + '_logits'])
        got = O.candidate_logits(sup, sub_def, x)
        assert ((got - ref).norm() / ref.norm()).item() < 2e-6, name
        m = O.eval_metrics(ref, torch.from_numpy(z[name + '_labels']))
        loss, acc1, acc5 = z[name + '_metrics']
        assert abs(m['loss'] - loss) < 1e-5 and m['acc1'] == acc1 and m['acc5'] == acc5
    # the q / k / v row blocks are cut separately (nets/net_utils.py:22-26)
    sub = O.sub_state_dict(sup, O.param_shapes(EVO_CANDIDATES['narrow']))
    w, ws = sup['blocks.1.attn.qkv.weight'], sub['blocks.1.attn.qkv.weight']
    assert ws.shape == (96, 56) and torch.equal(ws[32:64], w[64:96, :56]) and torch.equal(ws[64:96], w[128:160, :56])


def test_vit16_oracle_vs_reference_golden():
    """Patch-16 super-network with a distillation token: oracle forward (+ sub-architecture draws) against the reference's outputs."""
    import numpy as np
    import os
    import torch
    from oracle import vit_res_oracle as O
    from oracle.cases import VIT16_CASES, VIT16_DEF, VIT16_SPACE
    shapes = O.param_shapes(VIT16_DEF, num_tokens=2, patch_output=False, patch_size=16)
    p = O.keyed_fill(shapes, seed=5)
    for name, case in VIT16_CASES.items():
        G = np.load(os.path.join(os.path.dirname(__file__), 'golden', name + '.npz'))
        x, _, _ = O.synthetic_batch(case['batch'], seed=99)
        train = case.get('train', True)
        keeps = None
        if case['supernet'] and train:
            smp = O.Sampler(VIT16_DEF, VIT16_SPACE, case['epa'], case['warmup'], case.get('single', False))
            smp.set_epoch(case['epoch'])
            torch.manual_seed(case['seed'])
            keeps = smp.sample(case['batch'])
            flat = [k[n] for k in keeps for n in ('embed', 'attn', 'layer', 'mlp') if n in k]
            assert flat == G['keeps'].tolist()
        with torch.no_grad():
            cls, dst = O.forward(p, VIT16_DEF, x, keeps, training=train, patch_output=False, num_tokens=2,
                                 eval_full_mask=case['supernet'] and not train)
        for got, want in ((cls, G['cls']), (dst, G['dst'])):
            want = torch.from_numpy(want)
            assert ((got - want).norm() / want.norm()).item() < 2e-6, name


def test_switch_token_mix_properties():
    """Size-independent properties of the augmentation (any batch / seed): soft targets are distributions; in the patch half the image-level
    target equals the mean of the per-patch targets (lam = 1 - box area); in the image half every patch target equals the image target;
    pixels outside the box are untouched and pixels inside come from the permuted partner."""
    import numpy as np
    import torch
    from oracle import vit_res_oracle as O
    for B, seed in ((10, 3), (7, 11), (32, 5)):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(B, 3, 56, 56, generator=g)
        y = torch.randint(0, 1000, (B,), generator=g)
        torch.manual_seed(seed)
        np.random.seed(seed)
        d = O.token_mix_draws(B, 4)
        out, t, pt = O.switch_token_mix(x, y, d, 4)
        n1 = B // 2
        assert torch.allclose(t.sum(-1), torch.ones(B), atol=1e-5) and torch.allclose(pt.sum(-1), torch.ones(B, 16), atol=1e-5)
        assert torch.allclose(pt[:n1].mean(1), t[:n1], atol=1e-6)
        assert torch.equal(pt[n1:], t[n1:].unsqueeze(1).expand(-1, 16, -1))
        y0, y1, x0, x1 = d['box']
        ps = 56 // 4
        inside = torch.zeros(56, 56, dtype=torch.bool)
        inside[ps * y0:ps * y1, ps * x0:ps * x1] = True
        assert torch.equal(out[:n1][..., ~inside], x[:n1][..., ~inside])
        assert torch.equal(out[:n1][..., inside], x[:n1][d['perm1']][..., inside])


def test_oracle_adamw_pinned_to_torch_optim():
    """O.adamw_step (the checker of the fused optimizer kernel) against the installed torch.optim.AdamW with timm's grouping, over
    several steps so that both moments and the bias corrections matter (not just lr * sign(g) of a first step)."""
    torch.manual_seed(0)
    shapes = {'blocks.0.attn.qkv.weight': (24, 16), 'blocks.0.attn.qkv.bias': (24,), 'blocks.0.norm1.weight': (16,), 'tokens': (1, 1, 16),
              'pos_embed': (1, 5, 16)}
    w0 = {k: torch.randn(s, dtype=torch.float64) for k, s in shapes.items()}
    mine = {k: v.clone() for k, v in w0.items()}
    theirs = {k: torch.nn.Parameter(v.clone()) for k, v in w0.items()}
    nd = lambda k, v: v.ndim <= 1 or k.endswith('.bias') or k == 'tokens'      # noqa: E731
    opt = torch.optim.AdamW([{'params': [v for k, v in theirs.items() if nd(k, v)], 'weight_decay': 0.},
                             {'params': [v for k, v in theirs.items() if not nd(k, v)], 'weight_decay': 0.05}], lr=3e-3, betas=(0.9, 0.999), eps=1e-8)
    state = {}
    for step in range(1, 8):
        grads = {k: torch.randn(s, dtype=torch.float64) * (0.1 if step % 2 else 3.0) for k, s in shapes.items()}
        for k in theirs:
            theirs[k].grad = grads[k].clone()
        opt.step()
        O.adamw_step(mine, grads, state, lr=3e-3, weight_decay=0.05, step=step)
        for k in shapes:
            assert rel(mine[k].numpy(), theirs[k].detach().numpy()) < 1e-12, (step, k)
    for k in shapes:
        st = opt.state[theirs[k]]
        assert rel(state[k][0].numpy(), st['exp_avg'].numpy()) < 1e-12 and rel(state[k][1].numpy(), st['exp_avg_sq'].numpy()) < 1e-12


def test_oracle_matches_baseline_size_golden():
    """The oracle at the real widths (sr_tiny, BASELINE configs[1]) against the reference's outputs, incl. gradient elements."""
    from oracle.cases import BASELINE_CASES, baseline_net, probe_vectors
    name = 'sr_tiny_multi'
    case = BASELINE_CASES[name]
    G = np.load(os.path.join(GOLD, name + '.npz'))
    nd, space = baseline_net(case['space'])
    w = O.keyed_fill(O.param_shapes(nd), seed=0)
    p = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in w.items()}
    B = case['batch']
    x, t, pt = O.synthetic_batch(B, seed=1234)
    smp = O.Sampler(nd, space, case['epa'], 0, case.get('single', False), False)
    smp.set_epoch(case['epoch'])
    torch.manual_seed(case['seed'])
    keeps = smp.sample(B)
    flat = [k[n] for k in keeps for n in ('embed', 'attn', 'layer', 'mlp') if n in k]
    assert flat == G['keeps'].tolist()
    loss, cls, patch = O.train_loss(p, nd, x, t, pt, keeps)
    loss.backward()
    assert rel(cls.detach().numpy(), G['cls']) < TOL and rel(patch.detach().numpy(), G['patch']) < TOL
    n = 0
    for k in G.files:
        if k.startswith('gs:'):
            g = p[k[3:]].grad
            g2 = g.reshape(g.shape[0], -1).double()
            lv, rv = probe_vectors(tuple(g2.shape))
            gn = float(G['gn:' + k[3:]])
            assert (g2[::7, ::11] - torch.from_numpy(G[k]).double()).norm().item() < TOL * gn, k
            assert (lv.double() @ g2 - torch.from_numpy(G['gl:' + k[3:]]).double()).norm().item() < TOL * gn, k
            assert (g2 @ rv.double() - torch.from_numpy(G['gr:' + k[3:]]).double()).norm().item() < TOL * gn, k
            n += 1
    assert n >= 14
