"""GPU parity of SwitchTokenMix (csrc/token_mix.cu through vit_search_b200.token_mixup) -- bit exact against the reference's own outputs
(tests/golden/token_mix.npz, written by oracle/make_golden_mixup.py from /root/reference) and against the oracle at train-step size."""
import os

import numpy as np
import pytest
import torch

from oracle import vit_res_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'token_mix.npz')


def test_token_mix_vs_reference_golden():
    from vit_search_b200.token_mixup import SwitchTokenMix
    z = np.load(GOLD)
    ncase = len([k for k in z.files if k.endswith('_meta')])
    for c in range(ncase):
        pl, seed = int(z['c%d_meta' % c][4]), int(z['c%d_meta' % c][5])
        samples, labels = torch.from_numpy(z['c%d_samples' % c]), torch.from_numpy(z['c%d_labels' % c])
        mix = SwitchTokenMix(pl, num_classes=1000, smoothing=0.1)
        torch.manual_seed(seed)          # the class draws with the reference's RNG protocol: same seeds, same augmentation
        np.random.seed(seed)
        keep = samples.clone().cuda()
        out, t, pt, kind = mix(keep, labels.cuda())
        torch.cuda.synchronize()
        assert kind == 'seq'
        assert torch.equal(out.cpu(), torch.from_numpy(z['c%d_out' % c])), c
        assert torch.equal(t.cpu(), torch.from_numpy(z['c%d_targets' % c])), c
        assert torch.equal(pt.cpu(), torch.from_numpy(z['c%d_ptargets' % c])), c
        assert torch.equal(keep.cpu(), samples)            # the input batch is left untouched


@pytest.mark.parametrize('B', [256, 33])
def test_token_mix_train_step_size_vs_oracle(B):
    from vit_search_b200.token_mixup import SwitchTokenMix
    g = torch.Generator().manual_seed(B)
    samples = torch.randn(B, 3, 224, 224, generator=g)
    labels = torch.randint(0, 1000, (B,), generator=g)
    mix = SwitchTokenMix(4)
    for seed in (0, 5):
        torch.manual_seed(seed)
        np.random.seed(seed)
        d = mix.draw(B)
        torch.manual_seed(seed)
        np.random.seed(seed)
        d_o = O.token_mix_draws(B, 4)
        assert d['box'] == d_o['box'] and d['lam1'] == d_o['lam1'] and d['lam2'] == d_o['lam2'] and torch.equal(d['perm1'], d_o['perm1'])
        out, t, pt, _ = mix(samples.cuda(), labels.cuda(), draws=d)
        ro, rt, rp = O.switch_token_mix(samples, labels, d_o, 4)
        assert torch.equal(out.cpu(), ro) and torch.equal(t.cpu(), rt) and torch.equal(pt.cpu(), rp)
        # soft targets are distributions
        assert torch.allclose(t.sum(-1).cpu(), torch.ones(B), atol=1e-5) and torch.allclose(pt.sum(-1).cpu(), torch.ones(B, 16), atol=1e-5)
