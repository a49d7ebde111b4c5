"""GPU parity of candidate evaluation on resident super-network weights: the super-network with prefix extents
must give the logits the REFERENCE gives for a freshly built dense sub-network loaded with sliced weights
(tests/golden/evo_eval.npz, oracle/make_golden_evo.py), and the device-side meters must match engine.evaluate's numbers."""
import os

import numpy as np
import pytest
import torch

from oracle import vit_res_oracle as O
from oracle.cases import EVO_SUPER_DEF, EVO_CANDIDATES

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'evo_eval.npz')


def _supernet(prec):
    from vit_search_b200 import core
    from vit_search_b200.nets import create_model
    core.set_precision(prec)
    m = create_model('flexible_vit_sr_patch14_224_patch_output', network_def=EVO_SUPER_DEF, num_classes=1000)
    z = np.load(GOLD)
    m.load_state_dict(O.keyed_fill(O.param_shapes(EVO_SUPER_DEF), seed=int(z['w_seed']), running_stats=True))
    return m.cuda().eval(), z


@pytest.mark.parametrize('prec,tol', [('fp32', 2e-4), ('bf16', 3e-2)])
def test_candidate_logits_vs_reference_golden(prec, tol):
    from vit_search_b200 import core
    from vit_search_b200.evo_eval import CandidateEvaluator
    try:
        m, z = _supernet(prec)
        x, _, _ = O.synthetic_batch(8, seed=int(z['x_seed']))
        ev = CandidateEvaluator(m, 'cuda')
        xd = x.cuda()
        for name, sub_def in EVO_CANDIDATES.items():
            got = ev.logits(sub_def, xd).float().cpu()
            ref = torch.from_numpy(z[name + '_logits'])
            err = ((got - ref).norm() / ref.norm()).item()
            assert err < tol, (name, prec, err)
        # back to the network's own behaviour afterwards; and the largest candidate IS the network
        own = m(xd).float().cpu()
        ref = torch.from_numpy(z['largest_logits'])
        assert ((own - ref).norm() / ref.norm()).item() < tol
    finally:
        core.set_precision('bf16')


def test_eval_meters_vs_reference_golden():
    from vit_search_b200.evo_eval import EvalMeters
    z = np.load(GOLD)
    meters = EvalMeters('cuda')
    per_batch = []
    for name in EVO_CANDIDATES:
        logits, labels = torch.from_numpy(z[name + '_logits']), torch.from_numpy(z[name + '_labels'])
        meters.reset()
        meters.update(logits.cuda(), labels.cuda())
        r = meters.result()
        loss, acc1, acc5 = z[name + '_metrics']
        assert abs(r['loss'] - loss) < 2e-5 and r['acc1'] == acc1 and r['acc5'] == acc5, (name, r, z[name + '_metrics'])
        per_batch.append((dict(loss=float(loss), acc1=float(acc1), acc5=float(acc5)), logits.shape[0]))
    # several ragged batches in one pass: the loader-level averaging of engine.evaluate
    meters.reset()
    batches = []
    g = torch.Generator().manual_seed(3)
    for rows in (8, 5, 1, 256):
        lg = torch.randn(rows, 1000, generator=g) * 3
        lb = torch.randint(0, 1000, (rows,), generator=g)
        lb[: rows // 2] = lg[: rows // 2].argmax(1)
        meters.update(lg.cuda(), lb.cuda())
        batches.append((O.eval_metrics(lg, lb), rows))
    want = O.evaluate_meters(batches)
    got = meters.result()
    assert abs(got['loss'] - want['loss']) < 2e-5 and abs(got['acc1'] - want['acc1']) < 1e-9 and abs(got['acc5'] - want['acc5']) < 1e-9


def test_candidate_score_over_loader_and_rejects_bad_defs():
    from vit_search_b200.evo_eval import CandidateEvaluator
    m, z = _supernet('bf16')
    ev = CandidateEvaluator(m, 'cuda')
    x, _, _ = O.synthetic_batch(8, seed=int(z['x_seed']))
    labels = torch.from_numpy(z['narrow_labels'])
    loader = [(x[:5], labels[:5]), (x[5:], labels[5:])]
    got = ev.score(EVO_CANDIDATES['narrow'], loader)
    ref_logits = torch.from_numpy(z['narrow_logits'])
    want = O.evaluate_meters([(O.eval_metrics(ref_logits[:5], labels[:5]), 5), (O.eval_metrics(ref_logits[5:], labels[5:]), 3)])
    assert abs(got['loss'] - want['loss']) < 5e-2
    assert m.active_subnet is None
    bad = list(EVO_CANDIDATES['narrow'])
    bad[1] = (1, (56, 3, 32), (56, 96), 1)              # more heads than the super-network has
    with pytest.raises(ValueError):
        ev.logits(tuple(bad), x.cuda())
    m.train()
    m.set_active_subnet(EVO_CANDIDATES['narrow'])
    with pytest.raises(RuntimeError):
        m(x.cuda())
    m.set_active_subnet(None)
