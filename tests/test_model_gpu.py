"""Whole-model GPU parity: the CUDA modules against the REFERENCE's own outputs (tests/golden, written by
oracle/make_golden.py from /root/reference) and against the CPU oracle, forward + backward + one optimizer step.

Tolerances (stated per SURVEY.md §8c / BASELINE.json north star "logits within 1e-3 rel"):
  fp32 parity path (fp32 storage, 6-term split-bf16 tensor-core GEMMs): logits rel-L2 <= 2e-4 asserted (5e-6 observed);
      parameter gradients rel-L2 <= 2e-4 (2e-5 observed), except the conv-stem conv/BN parameters: 5e-3.
  bf16 training path: logits rel-L2 <= 3e-2, loss abs <= 3e-2 (the reference's own bf16 autocast sits 8.4e-3 from fp64),
      parameter gradients <= 8e-2, stem conv/BN parameters <= 0.3.
Why the stem is looser: its gradients pass through three BatchNorm+ReLU layers whose batch reductions cancel almost
completely (|sum dz| ~ sqrt(P) |dz| over P = B*112*112 pixels), so ONE ReLU unit flipping sign because a pre-activation moved
by 1e-7 relative changes d(beta) by ~1/sqrt(P) ~ 3e-3.  The reference's own fp32 run differs from its fp64 run by that much
on these tensors; with fp32-exact GEMMs our stem gradients match the fp64 oracle to 2e-5 (test_model_full_gradients_vs_oracle).
"""
import os

import numpy as np
import pytest
import torch

from oracle import vit_res_oracle as O
from oracle.cases import CASES, SMALL_DEF, SMALL_SPACE, VIT_RES_TINY

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def rel(a, b):
    a = torch.as_tensor(np.asarray(a)).double() if not isinstance(a, torch.Tensor) else a.double().cpu()
    b = torch.as_tensor(np.asarray(b)).double() if not isinstance(b, torch.Tensor) else b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build(case, drop_path_rate=0.0):
    from vit_search_b200.nets import create_model
    nd = VIT_RES_TINY if case['net'] == 'vit_res_tiny' else SMALL_DEF
    kw = dict(network_def=nd, num_classes=1000, drop_rate=0., drop_path_rate=drop_path_rate)
    if case['supernet']:
        m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', num_channels_to_keep=SMALL_SPACE,
                         example_per_arch=case['epa'], num_warmup_epochs=case['warmup'], single_arch=case.get('single', False),
                         hybrid_arch=case.get('hybrid', False), **kw)
    else:
        m = create_model('flexible_vit_sr_patch14_224_patch_output', **kw)
    m = m.cuda()
    if case['supernet']:
        m.set_epoch(case['epoch'])
    m.load_state_dict(O.keyed_fill(O.param_shapes(nd), seed=case.get('wseed', 0)))
    return m, nd


@pytest.mark.parametrize('name', [n for n in CASES])
@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_model_vs_reference_golden(name, prec):
    from vit_search_b200 import core
    from vit_search_b200.engine import SoftTargetCrossEntropy
    case = CASES[name]
    G = np.load(os.path.join(GOLD, name + '.npz'))
    m, nd = build(case)
    B = case['batch']
    x, t, pt = O.synthetic_batch(B, seed=case.get('xseed', 1234))
    x, t, pt = x.cuda(), t.cuda(), pt.cuda()
    train = case.get('train', True)
    m.train(train)
    crit = SoftTargetCrossEntropy()
    tol_logit, tol_grad, tol_stem = (2e-4, 2e-4, 5e-3) if prec == 'fp32' else (3e-2, 8e-2, 0.3)
    with core.precision(prec):
        torch.manual_seed(case['seed'])
        if not train:
            with torch.no_grad():
                cls = m(x)
            assert rel(cls, G['cls']) < tol_logit
            return
        cls, patch = m(x, patch_output_type='seq')
        loss = crit(cls, t) + crit(patch, pt)
        loss.backward()
    torch.cuda.synchronize()
    if case['supernet']:
        flat = [k[n] for k in m.last_keeps for n in ('embed', 'attn', 'layer', 'mlp') if n in k]
        assert flat == G['keeps'].tolist(), 'sub-architecture draws differ from the reference'
    errs = {'cls': rel(cls, G['cls']), 'patch': rel(patch, G['patch'])}
    assert errs['cls'] < tol_logit and errs['patch'] < tol_logit, errs
    assert abs(loss.item() - float(G['loss'])) < (1e-4 if prec == 'fp32' else 3e-2)
    gerr = {}
    for k, p in m.named_parameters():
        gn = float(G['gn:' + k])
        if 'g:' + k in G.files:
            ref = torch.from_numpy(G['g:' + k])
            gerr[k] = rel(p.grad, ref) if gn > 0 else p.grad.norm().item()
        else:
            gerr[k] = abs(p.grad.double().norm().item() - gn) / max(gn, 1e-30) if gn > 0 else p.grad.norm().item()
    stem = lambda k: k.startswith('patch_embed.conv') and 'conv_proj' not in k    # noqa: E731
    bad = {k: v for k, v in gerr.items() if not v < (tol_stem if stem(k) else tol_grad)}
    assert not bad, (prec, bad)
    if prec == 'fp32':
        for k, v in m.state_dict().items():
            if 'running' in k or 'tracked' in k:
                assert rel(v.float(), G['s:' + k].astype(np.float32)) < 1e-4, k


def test_model_full_gradients_vs_oracle():
    """Every gradient element (not just norms) of a multi-architecture step against the fp64 oracle."""
    from vit_search_b200 import core
    from vit_search_b200.engine import SoftTargetCrossEntropy
    case = CASES['small_multi']
    m, nd = build(case)
    B = case['batch']
    x, t, pt = O.synthetic_batch(B, seed=1234)
    m.train()
    with core.precision('fp32'):
        torch.manual_seed(case['seed'])
        cls, patch = m(x.cuda(), patch_output_type='seq')
        crit = SoftTargetCrossEntropy()
        (crit(cls, t.cuda()) + crit(patch, pt.cuda())).backward()
    w = O.keyed_fill(O.param_shapes(nd), dtype=torch.float64)
    p = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in w.items()}
    loss_o, cls_o, patch_o = O.train_loss(p, nd, x.double(), t.double(), pt.double(), m.last_keeps)
    loss_o.backward()
    assert rel(cls, cls_o) < 1e-4 and rel(patch, patch_o) < 1e-4
    bad = {}
    for k, prm in m.named_parameters():
        e = rel(prm.grad, p[k].grad)
        if not e < 1e-4:
            bad[k] = e
    assert not bad, bad


def test_drop_path_and_train_step():
    """Drop-path scales + one fused AdamW step against the oracle's restatement of torch.optim.AdamW / timm grouping."""
    from vit_search_b200 import core
    from vit_search_b200.engine import TrainStep, FusedAdamW
    case = CASES['small_single']
    m, nd = build(case, drop_path_rate=0.3)
    B = case['batch']
    x, t, pt = O.synthetic_batch(B, seed=5)
    m.train()
    opt = FusedAdamW(m, lr=1e-3, weight_decay=0.05)
    step = TrainStep(m, opt, arch_sample='single')
    w0 = {k: v.detach().clone().double().cpu() for k, v in m.state_dict().items()}
    # capture the drop-path table the model draws
    import vit_search_b200.nets.vit_sr_supernet as V
    drawn = {}
    orig_rand = torch.rand

    def spy(*a, **kw):
        r = orig_rand(*a, **kw)
        if kw.get('device') is not None and len(a) == 1 and len(a[0]) == 3:
            drawn['u'] = r.clone()
        return r
    torch.rand = spy
    try:
        with core.precision('fp32'):
            loss = step(x.cuda(), t.cuda(), pt.cuda(), epoch=3)
    finally:
        torch.rand = orig_rand
    torch.cuda.synchronize()
    depth = drawn['u'].shape[0]
    rates = torch.linspace(0, 0.3, depth)
    dpk = [((1 - rates[i] + drawn['u'][i, 0].cpu()).floor().tolist(), (1 - rates[i] + drawn['u'][i, 1].cpu()).floor().tolist())
           for i in range(depth)]
    p = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in w0.items()}
    loss_o, _, _ = O.train_loss(p, nd, x.double(), t.double(), pt.double(), m.last_keeps, drop_path_rate=0.3, dp_keeps=dpk)
    loss_o.backward()
    assert abs(loss.item() - loss_o.item()) < 1e-4
    params = {k: p[k].detach().clone() for k, _ in m.named_parameters()}
    grads = {k: p[k].grad for k in params}
    O.adamw_step(params, grads, {}, lr=1e-3, weight_decay=0.05, step=1)
    sd = m.state_dict()
    # AdamW's first step moves every weight by ~lr*sign(g): compare the UPDATE, not the weight
    bad = {}
    for k in params:
        du = (sd[k].double().cpu() - w0[k]).flatten()
        do = (params[k] - w0[k]).flatten()
        sel = grads[k].flatten().abs() > 1e-7      # sign(g) is ill-conditioned for (near-)zero gradients
        if sel.any():
            e = ((du - do)[sel].norm() / do[sel].norm().clamp_min(1e-30)).item()
            if not e < 2e-2:
                bad[k] = e
    assert not bad, bad


def test_fused_adamw_operand_shadows():
    """The fused optimizer kernel also rewrites the bf16 GEMM-operand copies of the Linear weights: after every step the copy the
    next forward will read must equal bf16(master weight), and training with shadows must match training with per-step re-casts."""
    from vit_search_b200 import core
    from vit_search_b200.engine import TrainStep, FusedAdamW
    case = CASES['small_single']
    B = case['batch']
    x, t, pt = O.synthetic_batch(B, seed=9)
    losses = {}
    for mode in ('shadow', 'recast'):
        torch.manual_seed(0)
        m, nd = build(case)
        m.train()
        opt = FusedAdamW(m, lr=1e-3, weight_decay=0.05)
        step = TrainStep(m, opt, arch_sample='single')
        out = []
        with core.precision('bf16'):
            for it in range(3):
                if mode == 'recast':
                    core.weights.clear()          # forget every cached / adopted operand copy: the forward re-casts from the masters
                out.append(step(x.cuda(), t.cuda(), pt.cuda(), epoch=1).item())
                if mode == 'shadow':
                    n_shadow = 0
                    for name, p in m.named_parameters():
                        if p.ndim == 2 and p.shape[1] % 8 == 0:
                            got = core.weights.get(p)
                            assert got.data_ptr() == opt.shadow[name].data_ptr(), name      # served from the optimizer's shadow ...
                            assert torch.equal(got, p.detach().to(torch.bfloat16)), name    # ... which equals bf16(master)
                            n_shadow += 1
                    assert n_shadow > 10
        losses[mode] = out
    # split-K / bias-gradient reductions use atomics, so two runs agree to rounding, not bit for bit
    assert all(abs(a - b) < 2e-3 * abs(b) for a, b in zip(losses['shadow'], losses['recast'])), losses


def test_model_ema_fused_into_optimizer():
    """ModelEma (timm ModelEmaV2 semantics) with the parameter averages updated inside the fused AdamW kernel and the buffers by
    ModelEma.update: after every step the EMA state must equal the oracle's recursion over the model's own post-step states."""
    from vit_search_b200 import core
    from vit_search_b200.engine import TrainStep, FusedAdamW, ModelEma
    case = CASES['small_single']
    B = case['batch']
    x, t, pt = O.synthetic_batch(B, seed=11)
    m, nd = build(case)
    m.train()
    decay = 0.9
    ema = ModelEma(m, decay=decay)
    step = TrainStep(m, FusedAdamW(m, lr=1e-3, weight_decay=0.05), arch_sample='single', model_ema=ema)
    ref = {k: v.detach().clone().cpu() for k, v in m.state_dict().items()}
    with core.precision('bf16'):
        for it in range(3):
            step(x.cuda(), t.cuda(), pt.cuda(), epoch=2)
            torch.cuda.synchronize()
            ref = O.ema_update(ref, {k: v.detach().cpu() for k, v in m.state_dict().items()}, decay)
    assert ema.fused
    got = ema.module.state_dict()
    assert set(got) == set(ref)
    for k in ref:
        if ref[k].is_floating_point():
            assert rel(got[k], ref[k]) < 1e-6, k
        else:
            assert torch.equal(got[k].cpu(), ref[k]), k
    # the average really moved away from the initial weights and differs from the live model
    name = next(n for n, p in m.named_parameters() if p.ndim == 2)
    assert rel(got[name], m.state_dict()[name]) > 1e-6


def test_device_feeder_pipeline():
    """engine.DeviceFeeder: batches submitted from pinned host memory come out in order with the right contents while a later batch is
    already in flight on the side stream, and a slot is not overwritten before its consumer has been released."""
    from vit_search_b200.engine import DeviceFeeder
    dev = torch.device('cuda', 0)
    feeder = DeviceFeeder(dev)
    host = [(torch.full((64, 3, 32, 32), float(i)).pin_memory(), torch.full((64, 10), float(-i)).pin_memory()) for i in range(5)]
    feeder.submit(*host[0])
    sums = []
    for i in range(5):
        if i + 1 < 5:
            feeder.submit(*host[i + 1])
        a, b = feeder.next()
        # a long-ish consumer on the compute stream: the next submit must wait for release() before it reuses this slot
        acc = a.float().sum() + b.float().sum()
        for _ in range(20):
            acc = acc + (a * 0).sum()
        sums.append(acc)
        feeder.release()
    torch.cuda.synchronize()
    for i, sacc in enumerate(sums):
        assert sacc.item() == float(i) * 64 * 3 * 32 * 32 - float(i) * 64 * 10, i
