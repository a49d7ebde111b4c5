"""Pin the oracle against the reference's patch-16 super-network with a distillation token (nets/vision_transformer_supernet.py,
executed from /root/reference through oracle/ref_shim.py) and write tests/golden/vit16_*.npz.  Test infrastructure."""
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, vit_res_oracle as O  # noqa: E402
from oracle.cases import VIT16_DEF, VIT16_SPACE, VIT16_CASES  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
TOL = 2e-5


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    R = ref_shim.load()
    spec = importlib.util.spec_from_file_location('nets.vision_transformer_supernet',
                                                  os.path.join(ref_shim.REF, 'nets', 'vision_transformer_supernet.py'))
    V = importlib.util.module_from_spec(spec)
    sys.modules['nets.vision_transformer_supernet'] = V
    spec.loader.exec_module(V)
    nd = VIT16_DEF
    shapes = O.param_shapes(nd, num_tokens=2, patch_output=False, patch_size=16)
    for name, case in VIT16_CASES.items():
        B = case['batch']
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            torch.manual_seed(0)
            if case['supernet']:
                ref = V.flexible_vit_patch16_224_supernet(network_def=nd, num_classes=1000, num_channels_to_keep=VIT16_SPACE,
                                                          example_per_arch=case['epa'], num_warmup_epochs=case['warmup'],
                                                          single_arch=case.get('single', False))
                ref.set_epoch(case['epoch'])
            else:
                ref = V.flexible_vit_patch16_224(network_def=nd, num_classes=1000)
        sd = ref.state_dict()
        assert list(sd.keys()) == list(shapes.keys()), (list(sd.keys()), list(shapes.keys()))
        for k in sd:
            assert tuple(sd[k].shape) == tuple(shapes[k]), (k, sd[k].shape, shapes[k])
        w = O.keyed_fill(shapes, seed=5)
        ref.load_state_dict(w)
        x, t, _ = O.synthetic_batch(B, seed=99)
        t2 = t.roll(1, dims=0)                              # a different soft target for the distillation head
        train = case.get('train', True)
        ref.train(train)
        torch.manual_seed(case['seed'])
        if train:
            cls_r, dst_r = ref(x)
            loss_r = O.soft_target_ce(cls_r, t) + O.soft_target_ce(dst_r, t2)
            loss_r.backward()
            grads_r = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
        else:
            with torch.no_grad():
                cls_r, dst_r = ref(x)
        masks_r = None
        if case['supernet'] and train:
            torch.manual_seed(case['seed'])
            rec = []
            CD = R['channel_drop'].ChannelDrop
            orig = CD.forward_mask

            def spy(self, inp):
                m = orig(self, inp)
                rec.append(m.sum(dim=(1, 2)).tolist())
                return m
            CD.forward_mask = spy
            try:
                with torch.no_grad():
                    ref(x)
            finally:
                CD.forward_mask = orig
            masks_r = rec
        # ---- oracle ----
        p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in w.items()}
        keeps = None
        if case['supernet'] and train:
            smp = O.Sampler(nd, VIT16_SPACE, case['epa'], case['warmup'], case.get('single', False))
            smp.set_epoch(case['epoch'])
            torch.manual_seed(case['seed'])
            keeps = smp.sample(B)
            flat = [k[n] for k in keeps for n in ('embed', 'attn', 'layer', 'mlp') if n in k]
            assert flat == masks_r, 'oracle mask draws differ from the reference'
        if train:
            cls_o, dst_o = O.forward(p, nd, x, keeps, training=True, patch_output=False, num_tokens=2)
            loss_o = O.soft_target_ce(cls_o, t) + O.soft_target_ce(dst_o, t2)
            loss_o.backward()
            e = {'cls': rel(cls_o, cls_r), 'dst': rel(dst_o, dst_r), 'loss': abs(loss_o.item() - loss_r.item())}
            for k, g in grads_r.items():
                e['g:' + k] = rel(p[k].grad, g) if g.norm() > 0 else p[k].grad.norm().item()
        else:
            with torch.no_grad():
                cls_o, dst_o = O.forward(p, nd, x, None, training=False, patch_output=False, num_tokens=2, eval_full_mask=case['supernet'])
            e = {'cls': rel(cls_o, cls_r), 'dst': rel(dst_o, dst_r)}
        worst = max(e, key=e.get)
        print('%-14s worst %-36s %.2e   (cls %.2e dst %.2e)' % (name, worst, e[worst], e['cls'], e['dst']))
        assert e[worst] < TOL, (name, worst, e[worst])
        out = {'cls': cls_r.detach().numpy(), 'dst': dst_r.detach().numpy()}
        if train:
            out['loss'] = np.array(loss_r.item())
            for k, g in grads_r.items():
                out['gn:' + k] = np.array(g.double().norm().item())
                if g.numel() <= 4096:
                    out['g:' + k] = g.numpy()
        if masks_r is not None:
            out['keeps'] = np.array(masks_r, dtype=np.int64)
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)
    print('wrote tests/golden/vit16_*.npz')


if __name__ == '__main__':
    main()
