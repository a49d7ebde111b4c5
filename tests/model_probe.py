"""Debug helper (not a test; lives under tests/ because it uses the oracle): per-tensor errors of the CUDA model vs the fp64 oracle for
one case (prints everything).  Usage: python tests/model_probe.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vit_res_oracle as O  # noqa: E402
from oracle.cases import CASES, SMALL_DEF, SMALL_SPACE, VIT_RES_TINY  # noqa: E402
from vit_search_b200 import core  # noqa: E402
from vit_search_b200.engine import SoftTargetCrossEntropy  # noqa: E402
from tests.test_model_gpu import build, rel  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'small_multi'
prec = sys.argv[2] if len(sys.argv) > 2 else 'fp32'
case = CASES[name]
m, nd = build(case)
B = case['batch']
x, t, pt = O.synthetic_batch(B, seed=case.get('xseed', 1234))
m.train()
crit = SoftTargetCrossEntropy()
with core.precision(prec):
    torch.manual_seed(case['seed'])
    # stem only
    pe = m.patch_embed(x.cuda())
    torch.manual_seed(case['seed'])
    m.load_state_dict(O.keyed_fill(O.param_shapes(nd), seed=case.get('wseed', 0)))
    cls, patch = m(x.cuda(), patch_output_type='seq')
    loss = crit(cls, t.cuda()) + crit(patch, pt.cuda())
    loss.backward()
w = O.keyed_fill(O.param_shapes(nd), dtype=torch.float64)
p = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in w.items()}
pe_o = O.patch_conv_embed(x.double(), p, 'patch_embed.', True, None)
print('stem out', rel(pe, pe_o))
loss_o, cls_o, patch_o = O.train_loss(p, nd, x.double(), t.double(), pt.double(), m.last_keeps if case['supernet'] else None)
loss_o.backward()
print('cls', rel(cls, cls_o), 'patch', rel(patch, patch_o), 'loss', loss.item(), loss_o.item())
for k, prm in m.named_parameters():
    print('%-40s %.3e   |g| %.3e' % (k, rel(prm.grad, p[k].grad), p[k].grad.norm().item()))
