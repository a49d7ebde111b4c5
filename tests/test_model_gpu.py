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


# ------------------------------------------------------------------------------------------------ BASELINE-size networks
def _build_baseline(case):
    from vit_search_b200.nets import create_model
    from oracle.cases import baseline_net
    nd, space = baseline_net(case['space'])
    m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_classes=1000, drop_rate=0., drop_path_rate=0.,
                     num_channels_to_keep=space, example_per_arch=case['epa'], num_warmup_epochs=0, single_arch=case.get('single', False)).cuda()
    m.set_epoch(case['epoch'])
    m.load_state_dict(O.keyed_fill(O.param_shapes(nd), seed=0))
    return m, nd


def _baseline_cases():
    from oracle.cases import BASELINE_CASES
    return BASELINE_CASES


@pytest.mark.parametrize('name', list(_baseline_cases()))
@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_baseline_size_vs_reference_golden(name, prec):
    """The real search spaces (sr_tiny = BASELINE configs[1], sr_tiny_mh = the published Tiny recipe, sr_small = configs[2]) at B = 8,
    one and four architectures per step, against the REFERENCE's outputs (oracle/make_golden_baseline.py): logits, loss, mask draws,
    the norm of every parameter gradient and, for the first block of every stage, the SR blocks and the heads, gradient ELEMENTS
    (full tensors, strided samples and +-1 projections that involve every element).

    Tolerances: fp32 parity path logits 2e-4, gradients 2e-4 (conv stem 5e-3).  bf16 training path: logits within 1.5 x the
    reference's OWN bf16-autocast error against fp64 (stored in the golden, ~6.5e-3) and <= 2e-2; gradients 8e-2 in
    norm and as elements (stem 0.3)."""
    from vit_search_b200 import core
    from vit_search_b200.engine import SoftTargetCrossEntropy
    from oracle.cases import probe_vectors
    case = _baseline_cases()[name]
    G = np.load(os.path.join(GOLD, name + '.npz'))
    m, nd = _build_baseline(case)
    B = case['batch']
    x, t, pt = O.synthetic_batch(B, seed=case.get('xseed', 1234))
    x, t, pt = x.cuda(), t.cuda(), pt.cuda()
    m.train()
    crit = SoftTargetCrossEntropy()
    with core.precision(prec):
        torch.manual_seed(case['seed'])
        cls, patch = m(x, patch_output_type='seq')
        loss = crit(cls, t) + crit(patch, pt)
        loss.backward()
    torch.cuda.synchronize()
    flat = [k[n] for k in m.last_keeps for n in ('embed', 'attn', 'layer', 'mlp') if n in k]
    assert flat == G['keeps'].tolist(), 'sub-architecture draws differ from the reference'
    errs = {'cls': rel(cls, G['cls']), 'patch': rel(patch, G['patch'])}
    if prec == 'fp32':
        tol_cls = tol_patch = 2e-4
        tol_grad, tol_stem, tol_loss = 2e-4, 5e-3, 1e-4
    else:
        tol_cls = min(2e-2, 1.5 * float(G['bf16_ref_err_cls']))
        tol_patch = min(2e-2, 1.5 * float(G['bf16_ref_err_patch']))
        tol_grad, tol_stem, tol_loss = 8e-2, 0.3, 3e-2
    print('%s %s logits err cls %.2e patch %.2e (reference bf16 autocast: %.2e / %.2e)' %
          (name, prec, errs['cls'], errs['patch'], float(G['bf16_ref_err_cls']), float(G['bf16_ref_err_patch'])))
    assert errs['cls'] < tol_cls and errs['patch'] < tol_patch, (errs, tol_cls, tol_patch)
    assert abs(loss.item() - float(G['loss'])) < tol_loss
    stem = lambda k: k.startswith('patch_embed.conv') and 'conv_proj' not in k    # noqa: E731
    bad, n_elem = {}, 0
    for k, p in m.named_parameters():
        gn = float(G['gn:' + k])
        tol = tol_stem if stem(k) else tol_grad
        g = p.grad
        if gn == 0:
            if not g.norm().item() == 0:
                bad[k] = ('nonzero', g.norm().item())
            continue
        e = abs(g.double().norm().item() - gn) / gn
        if not e < tol:
            bad[k] = ('norm', e)
        if 'g:' + k in G.files:
            e = rel(g, G['g:' + k])
            n_elem += 1
            if not e < tol:
                bad[k] = ('elements', e)
        elif 'gs:' + k in G.files:
            g2 = g.reshape(g.shape[0], -1).double().cpu()
            lv, rv = probe_vectors(tuple(g2.shape))
            # sample: relative to the RMS of the whole tensor x sqrt(sample size) (a sample may be mostly masked-out zeros)
            s_ref = torch.from_numpy(G['gs:' + k]).double()
            e_s = ((g2[::7, ::11] - s_ref).norm() / max(s_ref.norm().item(), 1e-30)).item()
            # projections: their error is a sum of g.numel() / len independent element errors -> compare against the same scale
            e_l = ((lv.double() @ g2 - torch.from_numpy(G['gl:' + k]).double()).norm() / (gn * 1.0)).item()
            e_r = ((g2 @ rv.double() - torch.from_numpy(G['gr:' + k]).double()).norm() / (gn * 1.0)).item()
            n_elem += 1
            if not (e_s < tol and e_l < tol and e_r < tol):
                bad[k] = ('sample/left/right', e_s, e_l, e_r)
    assert n_elem >= 14
    assert not bad, (prec, bad)


def test_fused_adamw_multi_step_vs_oracle_moments():
    """Five fused AdamW steps on fixed synthetic gradients (so that nothing but the optimizer is compared): parameters, both moments
    and the bf16 operand shadows against O.adamw_step, which tests/test_oracle_cpu.py pins to torch.optim.AdamW.  After five steps
    with gradients of alternating scale the update depends on the moment history and both bias corrections."""
    from vit_search_b200 import core
    from vit_search_b200.engine import FusedAdamW
    case = CASES['small_single']
    m, nd = build(case)
    opt = FusedAdamW(m, lr=2e-3, weight_decay=0.05)
    names = [n for n, _ in m.named_parameters()]
    params = {n: p.detach().double().cpu().clone() for n, p in m.named_parameters()}
    state = {}
    g = torch.Generator().manual_seed(3)
    with core.precision('bf16'):
        for step in range(1, 6):
            grads = {n: torch.randn(params[n].shape, generator=g) * (1e-3 if step % 2 else 0.5) for n in names}
            for n, p in m.named_parameters():
                p.grad = grads[n].cuda()
            opt.step()
            O.adamw_step(params, {n: grads[n].double() for n in names}, state, lr=2e-3, weight_decay=0.05, step=step)
    torch.cuda.synchronize()
    bad = {}
    for n, p in m.named_parameters():
        e = [rel(p.detach(), params[n]), rel(opt.state[n][0], state[n][0]), rel(opt.state[n][1], state[n][1])]
        if not max(e) < 2e-6:
            bad[n] = e
        if n in opt.shadow:
            assert torch.equal(opt.shadow[n], p.detach().to(torch.bfloat16)), n
    assert not bad, bad
    # checkpoint round trip on the device: a fresh optimizer that loads the state continues identically
    sd = opt.state_dict()
    opt2 = FusedAdamW(m, lr=2e-3, weight_decay=0.05)
    opt2.load_state_dict(sd)
    assert opt2.step_count == 5
    n0 = names[3]
    assert torch.equal(opt2.state[n0][0], opt.state[n0][0]) and opt2.state[n0][0].data_ptr() != opt.state[n0][0].data_ptr()


def test_finite_loss_guard_skips_the_update_on_the_device():
    """engine.py:168-173 aborts on a non-finite loss after a host read-back every step.  Here the optimizer kernel checks the loss on the
    device: a step with an inf / nan loss changes nothing and raises a sticky counter that TrainStep.check_finite() turns into the
    reference's abort at the logging interval."""
    from vit_search_b200 import core
    from vit_search_b200.engine import TrainStep, FusedAdamW
    case = CASES['small_single']
    m, nd = build(case)
    m.train()
    B = case['batch']
    x, t, pt = O.synthetic_batch(B, seed=5)
    step = TrainStep(m, FusedAdamW(m, lr=1e-3), arch_sample='single')
    with core.precision('bf16'):
        step(x.cuda(), t.cuda(), pt.cuda(), epoch=0)
        torch.cuda.synchronize()
        step.check_finite()                                   # finite so far
        w1 = {k: v.detach().clone() for k, v in m.named_parameters()}
        m1 = {k: v[0].clone() for k, v in step.optimizer.state.items()}
        tb = t.clone()
        tb[0, 0] = float('inf')           # (a NaN pixel would not do: the stem's ReLU is fmaxf, which maps NaN to 0)
        loss = step(x.cuda(), tb.cuda(), pt.cuda(), epoch=0)
        torch.cuda.synchronize()
        assert not torch.isfinite(loss).item()
        for k, v in m.named_parameters():
            assert torch.equal(v.detach(), w1[k]), k           # nothing moved
        for k, v in step.optimizer.state.items():
            assert torch.equal(v[0], m1[k]), k
        with pytest.raises(FloatingPointError, match='not finite'):
            step.check_finite()


def test_device_feeder_uint8_normalize():
    """DeviceFeeder(normalize=(mean, std)): a uint8 image batch uploaded as 1 byte per pixel comes out as the fp32 batch torchvision's
    ToTensor + Normalize would have produced on the host; the other tensors of the batch pass through unchanged."""
    from vit_search_b200.engine import DeviceFeeder
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    dev = torch.device('cuda', 0)
    feeder = DeviceFeeder(dev, normalize=(mean, std))
    g = torch.Generator().manual_seed(0)
    for it in range(3):
        u8 = torch.randint(0, 256, (5, 3, 224, 224), generator=g, dtype=torch.uint8).pin_memory()
        tgt = torch.randn(5, 10, generator=g).pin_memory()
        feeder.submit(u8, tgt)
        x, t = feeder.next()
        ref = (u8.float() / 255.0 - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
        assert x.dtype == torch.float32 and (x.cpu() - ref).abs().max().item() < 1e-6
        assert torch.equal(t.cpu(), tgt)
        feeder.release()
