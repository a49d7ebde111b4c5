"""GPU parity of the patch-16 super-network with a distillation token against the REFERENCE's outputs
(tests/golden/vit16_*.npz, written by oracle/make_golden_vit16.py from /root/reference/nets/vision_transformer_supernet.py).
Tolerances as in test_model_gpu.py: fp32 parity path 2e-4, bf16 training path logits 3e-2 / gradients 8e-2."""
import os

import numpy as np
import pytest
import torch

from oracle import vit_res_oracle as O
from oracle.cases import VIT16_CASES, VIT16_DEF, VIT16_SPACE

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def rel(a, b):
    a = torch.as_tensor(np.asarray(a)).double() if not isinstance(a, torch.Tensor) else a.double().cpu()
    b = torch.as_tensor(np.asarray(b)).double() if not isinstance(b, torch.Tensor) else b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build(case):
    from vit_search_b200.nets import create_model
    if case['supernet']:
        m = create_model('flexible_vit_patch16_224_supernet', network_def=VIT16_DEF, num_classes=1000, num_channels_to_keep=VIT16_SPACE,
                         example_per_arch=case['epa'], num_warmup_epochs=case['warmup'], single_arch=case.get('single', False)).cuda()
        m.set_epoch(case['epoch'])
    else:
        m = create_model('flexible_vit_patch16_224', network_def=VIT16_DEF, num_classes=1000).cuda()
    m.load_state_dict(O.keyed_fill(O.param_shapes(VIT16_DEF, num_tokens=2, patch_output=False, patch_size=16), seed=5))
    return m


@pytest.mark.parametrize('name', list(VIT16_CASES))
@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_vit16_vs_reference_golden(name, prec):
    from vit_search_b200 import core
    from vit_search_b200.engine import SoftTargetCrossEntropy
    case = VIT16_CASES[name]
    G = np.load(os.path.join(GOLD, name + '.npz'))
    m = build(case)
    B = case['batch']
    x, t, _ = O.synthetic_batch(B, seed=99)
    t2 = t.roll(1, dims=0)
    x, t, t2 = x.cuda(), t.cuda(), t2.cuda()
    train = case.get('train', True)
    m.train(train)
    tol_logit, tol_grad = (2e-4, 2e-4) if prec == 'fp32' else (3e-2, 8e-2)
    with core.precision(prec):
        torch.manual_seed(case['seed'])
        if not train:
            with torch.no_grad():
                cls, dst = m(x)
            assert rel(cls, G['cls']) < tol_logit and rel(dst, G['dst']) < tol_logit
            return
        cls, dst = m(x)
        crit = SoftTargetCrossEntropy()
        loss = crit(cls, t) + crit(dst, t2)
        loss.backward()
    torch.cuda.synchronize()
    if case['supernet']:
        flat = [k[n] for k in m.last_keeps for n in ('embed', 'attn', 'layer', 'mlp') if n in k]
        assert flat == G['keeps'].tolist(), 'sub-architecture draws differ from the reference'
    assert rel(cls, G['cls']) < tol_logit and rel(dst, G['dst']) < tol_logit
    assert abs(loss.item() - float(G['loss'])) < (1e-4 if prec == 'fp32' else 3e-2)
    bad = {}
    for k, p in m.named_parameters():
        gn = float(G['gn:' + k])
        if p.grad is None:
            e = gn
        elif 'g:' + k in G.files:
            e = rel(p.grad, torch.from_numpy(G['g:' + k])) if gn > 0 else p.grad.norm().item()
        else:
            e = abs(p.grad.double().norm().item() - gn) / max(gn, 1e-30) if gn > 0 else p.grad.norm().item()
        if not e < tol_grad:
            bad[k] = e
    assert not bad, (prec, bad)
