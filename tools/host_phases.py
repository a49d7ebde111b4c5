"""Host-side phase timing of the bench train step (no profiler): where does the Python thread spend its wall time, and does it block?"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import core, supernet_config as sc  # noqa: E402
from vit_search_b200.engine import FusedAdamW, TrainStep  # noqa: E402
from vit_search_b200.nets import create_model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nd, ks = sc.network_def('sr_tiny'), sc.num_channels_to_keep('sr_tiny')
torch.manual_seed(0)
m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_classes=1000, drop_path_rate=0.2,
                 num_channels_to_keep=ks, example_per_arch=B, num_warmup_epochs=0, single_arch=True).cuda()
m.set_epoch(0)
m.train()
core.set_precision('bf16')
opt = FusedAdamW(m)
step = TrainStep(m, opt, arch_sample='single')
x = torch.randn(B, 3, 224, 224, device='cuda')
t = torch.softmax(torch.randn(B, 1000, device='cuda'), -1)
pt = t.unsqueeze(1).repeat(1, 16, 1).contiguous()
for _ in range(5):
    step(x, t, pt)
torch.cuda.synchronize()
crit = step.criterion
rows = []
T0 = time.perf_counter()
for it in range(8):
    t0 = time.perf_counter()
    torch.manual_seed(it)
    cls_pred, patch_pred = m(x, patch_output_type='seq')
    t1 = time.perf_counter()
    loss = crit(cls_pred, t) + crit(patch_pred, pt)
    opt.zero_grad()
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    opt.step()
    t4 = time.perf_counter()
    rows.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
torch.cuda.synchronize()
T1 = time.perf_counter()
print('B=%d  wall per step %.2f ms' % (B, (T1 - T0) / 8 * 1e3))
print('step | forward  loss   backward  optimizer  (host ms)')
for i, r in enumerate(rows):
    print('%4d | %7.2f %6.2f %9.2f %9.2f   sum %.2f' % ((i,) + tuple(v * 1e3 for v in r) + (sum(r) * 1e3,)))
