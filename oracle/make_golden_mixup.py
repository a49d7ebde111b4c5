"""Pin oracle.switch_token_mix against the reference's own SwitchTokenMix (token_mixup.py, executed from /root/reference with the
`.cuda()` / device='cuda' calls neutralised) and write tests/golden/token_mix.npz.  Test infrastructure; needs /root/reference."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vit_res_oracle as O  # noqa: E402

REF = os.environ.get('VIT_SEARCH_REFERENCE', '/root/reference')


def load_reference():
    spec = importlib.util.spec_from_file_location('ref_token_mixup', os.path.join(REF, 'token_mixup.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class cpu_device:
    """Run the reference on the CPU: Tensor.cuda() is the identity and factory calls ignore device='cuda'."""

    def __enter__(self):
        self.cuda = torch.Tensor.cuda
        self.full, self.zeros = torch.full, torch.zeros
        torch.Tensor.cuda = lambda t, *a, **k: t

        def strip(fn):
            def wrapped(*a, **k):
                k.pop('device', None)
                return fn(*a, **k)
            return wrapped
        torch.full, torch.zeros = strip(self.full), strip(self.zeros)

    def __exit__(self, *exc):
        torch.Tensor.cuda = self.cuda
        torch.full, torch.zeros = self.full, self.zeros


def main():
    ref = load_reference()
    out = {}
    for case, (B, HW, pl, seed) in enumerate([(6, 56, 4, 0), (8, 28, 4, 1), (5, 56, 4, 2), (16, 56, 4, 3), (8, 56, 4, 7)]):
        g = torch.Generator().manual_seed(100 + seed)
        samples = torch.randn(B, 3, HW, HW, generator=g)
        labels = torch.randint(0, 1000, (B,), generator=g)
        torch.manual_seed(seed)
        np.random.seed(seed)
        draws = O.token_mix_draws(B, pl)
        mine = O.switch_token_mix(samples, labels, draws, pl)
        torch.manual_seed(seed)
        np.random.seed(seed)
        with cpu_device():
            mix = ref.SwitchTokenMix(pl, num_classes=1000, smoothing=0.1)
            r_s, r_t, r_p, kind = mix(samples.clone(), labels.clone())
        assert kind == 'seq'
        for a, b, name in ((mine[0], r_s, 'samples'), (mine[1], r_t, 'targets'), (mine[2], r_p, 'patch_targets')):
            assert torch.equal(a, b), (case, name, (a - b).abs().max().item())
        out['c%d_samples' % case] = samples.numpy()
        out['c%d_labels' % case] = labels.numpy()
        out['c%d_perm1' % case] = draws['perm1'].numpy()
        out['c%d_perm2' % case] = draws['perm2'].numpy()
        out['c%d_meta' % case] = np.array(list(draws['box']) + [pl, seed], dtype=np.int64)
        out['c%d_lams' % case] = np.array([draws['lam1'], draws['lam2']], dtype=np.float64)
        out['c%d_out' % case] = r_s.numpy()
        out['c%d_targets' % case] = r_t.numpy()
        out['c%d_ptargets' % case] = r_p.numpy().astype(np.float16 if False else np.float32)
        print('case %d: oracle == reference (bit exact); box %s lam1 %.4f lam2 %.4f' % (case, draws['box'], draws['lam1'], draws['lam2']))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'token_mix.npz'), **out)


if __name__ == '__main__':
    main()
