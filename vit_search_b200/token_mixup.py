"""SwitchTokenMix (reference token_mixup.py:39-162; built at main.py:316-322, called at engine.py:109-110) on the GPU.

Same constructor and call signature as the reference class.  The random draws follow the reference's protocol exactly -- the global
torch CPU generator for the two permutations and numpy's global RandomState for the box and the lambdas, in the same order -- so the
same seeds give the same augmentation; the tensors are produced by csrc/token_mix.cu (vsx_token_mix) in two launches.  Unlike the
reference, `samples` is not overwritten: the mixed batch is returned as a new tensor.
"""
import numpy as np
import torch

from . import core, ops


def _my_randint(low, high, size=None):                      # token_mixup.py:32-35
    if low == high:
        high = low + 1
    return np.random.randint(low, high, size=size)


class SwitchTokenMix:
    def __init__(self, patch_len, switch_prob=0.5, num_classes=1000, smoothing=0.1):
        self.patch_len = patch_len
        self.switch_prob = switch_prob
        self.num_classes = num_classes
        self.smoothing = smoothing

    def __repr__(self):
        return '(patch_len={}, switch_prob={})'.format(self.patch_len, self.switch_prob)

    def draw(self, batch):
        """One call's random draws, consumed in the reference's order (:112-114 -> :75-99, then :131-134)."""
        pl = self.patch_len
        n1 = batch // 2
        perm1 = torch.randperm(n1)
        lam = np.random.beta(1., 1.)
        area = int(pl * pl * lam)
        max_length = min(pl, area)
        cut_h = _my_randint(1, max(1, max_length - 1))
        cut_w = area // cut_h
        if cut_w > pl:
            cut_w = pl
            cut_h = area // cut_w
        yl = _my_randint(0, max(0, pl - cut_h), size=2)
        xl = _my_randint(0, max(0, pl - cut_w), size=2)
        y0, x0 = int(yl[1]), int(xl[1])
        lam1 = 1 - (cut_h * cut_w + 0.0) / (pl * pl)
        perm2 = torch.randperm(batch - n1)
        lam2 = np.random.beta(0.8, 0.8)
        return dict(perm1=perm1, box=(y0, y0 + int(cut_h), x0, x0 + int(cut_w)), lam1=float(lam1), perm2=perm2, lam2=float(lam2))

    def __call__(self, samples, targets, draws=None):
        core.require_cuda(samples, 'SwitchTokenMix')
        B, C, H, W = samples.shape
        d = self.draw(B) if draws is None else draws
        dev = samples.device
        x = samples.contiguous().float()
        labels = targets.to(device=dev, dtype=torch.int64).contiguous()
        out = torch.empty_like(x)
        P, K = self.patch_len * self.patch_len, self.num_classes
        new_targets = torch.empty(B, K, device=dev)
        patch_targets = torch.empty(B, P, K, device=dev)
        p1 = core.h2d(d['perm1'].to(torch.int32), dev)
        p2 = core.h2d(d['perm2'].to(torch.int32), dev)
        off = self.smoothing / K
        on = 1. - self.smoothing + off
        y0, y1, x0, x1 = d['box']
        ops.call('token_mix', x, out, labels, p1, p2, new_targets, patch_targets, B, C, H, W, self.patch_len, K, y0, y1, x0, x1,
                 float(np.float32(on)), float(np.float32(off)), float(d['lam1']), float(d['lam2']))
        return out, new_targets, patch_targets, 'seq'
