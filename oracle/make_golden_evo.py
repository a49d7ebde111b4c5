"""Pin oracle.candidate_logits / sub_state_dict / eval_metrics against the reference's evolutionary-search evaluation
(evo_search.py:256-273: dense sub-network from the reference factory + nets/net_utils.get_sub_state_dict + eval forward;
engine.py:195-228: CrossEntropyLoss + timm accuracy) and write tests/golden/evo_eval.npz.  Test infrastructure; needs /root/reference."""
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, vit_res_oracle as O  # noqa: E402
from oracle.cases import EVO_SUPER_DEF, EVO_CANDIDATES  # noqa: E402

BATCH = 8


def timm_accuracy(output, target, topk=(1,)):
    """timm 0.3.2 utils.accuracy (package absent here; published formula): top-k hits in percent."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum(0) * 100. / target.size(0) for k in topk]


def main():
    R = ref_shim.load()
    V = R['vit_sr_supernet']
    spec = importlib.util.spec_from_file_location('nets.net_utils', os.path.join(ref_shim.REF, 'nets', 'net_utils.py'))
    NU = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(NU)

    super_sd = O.keyed_fill(O.param_shapes(EVO_SUPER_DEF), seed=21, running_stats=True)
    x, _, _ = O.synthetic_batch(BATCH, seed=4321)
    out = {'x_seed': np.array(4321), 'w_seed': np.array(21)}
    for name, sub_def in EVO_CANDIDATES.items():
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            torch.manual_seed(0)
            model = V.flexible_vit_sr_patch14_224_patch_output(network_def=sub_def, num_classes=1000)      # evo_search.py:256-262
        sub_sd = NU.get_sub_state_dict(source_dict=super_sd, sub_dict=model.state_dict())
        model.load_state_dict(sub_sd)
        model.eval()
        with torch.no_grad():
            logits_r = model(x)
        # oracle: slicing and forward
        mine_sd = O.sub_state_dict(super_sd, O.param_shapes(sub_def))
        assert list(mine_sd.keys()) == list(sub_sd.keys())
        for k in sub_sd:
            assert torch.equal(mine_sd[k], sub_sd[k]), k
        logits_o = O.candidate_logits(super_sd, sub_def, x)
        err = ((logits_o - logits_r).norm() / logits_r.norm()).item()
        assert err < 2e-6, (name, err)
        # labels with a spread of ranks: argmax, 3rd, 7th, random
        order = logits_r.argsort(dim=1, descending=True)
        g = torch.Generator().manual_seed(5)
        labels = torch.stack([order[b, [0, 2, 6, 0, 4, 17, 1, 0][b]] for b in range(BATCH)])
        labels[5] = torch.randint(0, 1000, (1,), generator=g)[0]
        loss_r = torch.nn.CrossEntropyLoss()(logits_r, labels)                                            # engine.py:195,222
        acc1_r, acc5_r = timm_accuracy(logits_r, labels, topk=(1, 5))                                      # engine.py:223
        m = O.eval_metrics(logits_r, labels)
        assert abs(m['loss'] - loss_r.item()) < 1e-5 and m['acc1'] == acc1_r.item() and m['acc5'] == acc5_r.item(), (m, loss_r, acc1_r, acc5_r)
        print('%-8s oracle vs reference logits rel err %.2e | loss %.5f acc1 %.1f acc5 %.1f' % (name, err, loss_r.item(), acc1_r.item(), acc5_r.item()))
        out[name + '_logits'] = logits_r.numpy()
        out[name + '_labels'] = labels.numpy()
        out[name + '_metrics'] = np.array([loss_r.item(), acc1_r.item(), acc5_r.item()])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'evo_eval.npz'), **out)
    print('wrote tests/golden/evo_eval.npz')


if __name__ == '__main__':
    main()
