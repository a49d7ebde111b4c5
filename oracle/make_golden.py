"""Pin the oracle against the reference itself and write tests/golden/*.npz  (test infrastructure).

Run in the build container (needs /root/reference):  ``python -m oracle.make_golden``

For every case the UNMODIFIED reference modules (through oracle/ref_shim.py) and the oracle
restatement are run on identical keyed weights, seeded inputs and CPU RNG state; the script asserts
agreement (fp32 tolerance below) and stores the REFERENCE's outputs as the golden vectors.  Inputs and
weights are not stored -- they are regenerated from seeds by ``oracle.vit_res_oracle.keyed_fill`` /
``synthetic_batch``.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, vit_res_oracle as O   # noqa: E402
from oracle.cases import SMALL_DEF, SMALL_SPACE, CASES, VIT_RES_TINY  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
TOL = 2e-5


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build_reference(R, network_def, space, case):
    V = R['vit_sr_supernet']
    kw = dict(network_def=network_def, num_classes=1000, drop_rate=0., drop_path_rate=0.)
    if space is not None:
        m = V.flexible_vit_sr_patch14_224_patch_output_supernet(
            num_channels_to_keep=space, example_per_arch=case['epa'], num_warmup_epochs=case['warmup'],
            single_arch=case.get('single', False), hybrid_arch=case.get('hybrid', False), **kw)
    else:
        m = V.flexible_vit_sr_patch14_224_patch_output(**kw)
    return m


def run_case(R, name, case):
    nd = VIT_RES_TINY if case['net'] == 'vit_res_tiny' else SMALL_DEF
    space = SMALL_SPACE if case['supernet'] else None
    B = case['batch']
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        torch.manual_seed(0)
        ref = build_reference(R, nd, space, case)
    if case['supernet']:
        ref.set_epoch(case['epoch'])            # rewiring happens here; weights are overwritten below
    shapes = O.param_shapes(nd)
    sd = ref.state_dict()
    assert list(sd.keys()) == list(shapes.keys()), 'state_dict key order differs from oracle.param_shapes'
    for k in sd:
        assert tuple(sd[k].shape) == tuple(shapes[k]), (k, sd[k].shape, shapes[k])
    w = O.keyed_fill(shapes, seed=case.get('wseed', 0))
    ref.load_state_dict(w)
    x, t, pt = O.synthetic_batch(B, seed=case.get('xseed', 1234))
    train = case.get('train', True)
    ref.train(train)

    # ---- reference ----
    torch.manual_seed(case['seed'])
    if train:
        cls_r, patch_r = ref(x, patch_output_type='seq')
        loss_r = O.soft_target_ce(cls_r, t) + O.soft_target_ce(patch_r, pt)
        loss_r.backward()
        grads_r = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
        stats_r = {k: v.detach().clone() for k, v in ref.state_dict().items() if 'running' in k or 'tracked' in k}
    else:
        with torch.no_grad():
            cls_r = ref(x)
    masks_r = None
    if case['supernet'] and train:
        # replay the draws to record the reference's masks (the RNG is reseeded identically)
        torch.manual_seed(case['seed'])
        rec = []
        CD = R['channel_drop'].ChannelDrop
        orig = CD.forward_mask          # some call sites use .forward() directly, so hooks would miss them

        def spy(self, inp):
            m = orig(self, inp)
            rec.append(m.sum(dim=(1, 2)).tolist())
            return m
        CD.forward_mask = spy
        try:
            with torch.no_grad():
                ref(x, patch_output_type='seq')
        finally:
            CD.forward_mask = orig
        masks_r = rec

    # ---- oracle ----
    p = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in w.items()}
    keeps = None
    if case['supernet'] and train:
        smp = O.Sampler(nd, space, case['epa'], case['warmup'], case.get('single', False), case.get('hybrid', False))
        smp.set_epoch(case['epoch'])
        torch.manual_seed(case['seed'])
        keeps = smp.sample(B)
        flat = [k[n] for k in keeps for n in ('embed', 'attn', 'layer', 'mlp') if n in k]
        assert flat == masks_r, 'oracle mask draws differ from the reference'
    new_stats = {}
    if train:
        loss_o, cls_o, patch_o = O.train_loss(p, nd, x, t, pt, keeps, new_stats=new_stats)
        loss_o.backward()
        e = {'cls': rel(cls_o, cls_r), 'patch': rel(patch_o, patch_r), 'loss': abs(loss_o.item() - loss_r.item())}
        for k, g in grads_r.items():
            e['g:' + k] = rel(p[k].grad, g) if g.norm() > 0 else p[k].grad.norm().item()
        for k, v in stats_r.items():
            e['s:' + k] = rel(new_stats[k].float(), v.float())
    else:
        with torch.no_grad():
            cls_o = O.forward(p, nd, x, None, training=False, eval_full_mask=case['supernet'])
        e = {'cls': rel(cls_o, cls_r)}
    worst = max(e, key=e.get)
    print('%-28s worst %-40s %.2e   (cls %.2e)' % (name, worst, e[worst], e['cls']))
    assert e[worst] < TOL, (name, worst, e[worst])

    out = {'cls': cls_r.detach().numpy()}
    if train:
        out['patch'] = patch_r.detach().numpy()
        out['loss'] = np.array(loss_r.item())
        for k, g in grads_r.items():
            # full tensors for the small ones, norms for everything (keeps the fixtures small)
            out['gn:' + k] = np.array(g.double().norm().item())
            if g.numel() <= 4096:
                out['g:' + k] = g.numpy()
        for k, v in stats_r.items():
            out['s:' + k] = v.numpy()
    if masks_r is not None:
        out['keeps'] = np.array(masks_r, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)


def function_level(R):
    """Known-answer vectors for the pieces the CUDA kernels restate one by one."""
    out = {}
    g = torch.Generator().manual_seed(7)
    # masked LN forward / the reference's hand-written backward
    B, N, C = 6, 5, 48
    keep = [48, 40, 32, 24, 48, 16]
    x = torch.randn(B, N, C, generator=g) * O.prefix_mask(keep, C, torch.float32)
    wt = 1 + 0.1 * torch.randn(C, generator=g)
    bs = 0.1 * torch.randn(C, generator=g)
    go = torch.randn(B, N, C, generator=g)
    M = R['masked_layer_norm']
    ln = M.MaskedLayerNorm(C)
    ln.weight.data.copy_(wt)
    ln.bias.data.copy_(bs)
    xr = x.clone().requires_grad_(True)
    mask = O.prefix_mask(keep, C)
    y = ln(xr, mask)
    y.backward(go)
    y_o = O.masked_layer_norm(x, wt, bs, keep)
    gx_o, gg_o, gb_o = O.masked_layer_norm_backward(go * mask, x, wt, keep)
    assert rel(y_o, y) < TOL and rel(gx_o, xr.grad) < TOL and rel(gg_o, ln.weight.grad) < TOL and rel(gb_o, ln.bias.grad) < TOL
    out.update(ln_keep=np.array(keep), ln_y=y.detach().numpy(), ln_gx=xr.grad.numpy(),
               ln_gw=ln.weight.grad.numpy(), ln_gb=ln.bias.grad.numpy())
    # ChannelDrop tables across warm-up epochs and draw protocol
    CD = R['channel_drop'].ChannelDrop
    tabs = []
    for epoch in (0, 2, 5, 9):
        for single in (False, True):
            cd = CD(np.array([96, 64, 128, 32, 80]), num_warmup_epochs=5, example_per_arch=2, single_arch=single)
            cd.set_epoch(epoch)
            cd.train()
            torch.manual_seed(100 + epoch)
            _, m = cd(torch.zeros(12, 3, 128))
            table_ref = cd.mask.sum(dim=(1, 2)).tolist()
            table_o = O.keep_table([96, 64, 128, 32, 80], 12, 2, single, epoch, 5)
            assert table_ref == table_o, (epoch, single, table_ref, table_o)
            torch.manual_seed(100 + epoch)
            assert O.draw_keep(table_o, 12, 2, single) == m.sum(dim=(1, 2)).tolist()
            tabs.append(m.sum(dim=(1, 2)).tolist())
    out['cd_draws'] = np.array(tabs)
    # drop-path arithmetic (nets/drop.py:11-26) with the Bernoulli draw made explicit
    D = R['drop']
    f = torch.randn(8, 3, 4, generator=g)
    torch.manual_seed(5)
    y = D.drop_path(f, 0.25, True)
    torch.manual_seed(5)
    keep_draw = (0.75 + torch.rand((8, 1, 1))).floor().view(-1).tolist()
    assert rel(O.drop_path_scale(f, keep_draw, 0.25), y) < 1e-6
    out.update(dp_keep=np.array(keep_draw), dp_y=y.numpy())
    # MAC counter known answers (network_utils/compute_flop_mac.py __main__ prints)
    est = R['compute_flop_mac'].ComputationEstimator(distill=False, input_resolution=224, patch_size=14)
    with contextlib.redirect_stdout(io.StringIO()):
        out['mac_vit_res_tiny'] = np.array(est(VIT_RES_TINY))
        out['mac_small_def'] = np.array(est(SMALL_DEF))
    np.savez_compressed(os.path.join(GOLD, 'functions.npz'), **out)
    print('function-level vectors ok; ViT-Res-Tiny MACs =', int(out['mac_vit_res_tiny']))


def main():
    os.makedirs(GOLD, exist_ok=True)
    R = ref_shim.load()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    function_level(R)
    for name, case in CASES.items():
        run_case(R, name, case)
    print('golden vectors written to', GOLD)


if __name__ == '__main__':
    main()
