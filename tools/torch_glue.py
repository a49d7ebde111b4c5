"""Which torch (ATen) kernels still run inside one bench train step, and from where: torch.profiler with stacks, libvsx kernels filtered out."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import core, supernet_config as sc  # noqa: E402
from vit_search_b200.engine import FusedAdamW, TrainStep  # noqa: E402
from vit_search_b200.nets import create_model  # noqa: E402

space, B = 'sr_tiny', 256
nd, ks = sc.network_def(space), sc.num_channels_to_keep(space)
torch.manual_seed(0)
m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_classes=1000, drop_path_rate=0.2,
                 num_channels_to_keep=ks, example_per_arch=B, num_warmup_epochs=0, single_arch=True).cuda()
m.set_epoch(0)
m.train()
core.set_precision('bf16')
step = TrainStep(m, FusedAdamW(m), arch_sample='single')
x = torch.randn(B, 3, 224, 224, device='cuda')
t = torch.softmax(torch.randn(B, 1000, device='cuda'), -1)
pt = t.unsqueeze(1).repeat(1, 16, 1).contiguous()
for _ in range(3):
    step(x, t, pt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step(x, t, pt)
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_stack_n=6)
rows = []
for e in ka:
    dt = getattr(e, 'self_device_time_total', 0) or getattr(e, 'self_cuda_time_total', 0)
    if e.key.startswith('aten::') and dt > 0:
        st = [s for s in (e.stack or []) if 'vit_search_b200' in s]
        rows.append((dt, e.count, e.key, st[0][-110:] if st else '?'))
rows.sort(reverse=True)
print('ATen ops with own device time in one step: %.0f us' % sum(r[0] for r in rows))
for r in rows[:50]:
    print('%7.1f us x%-3d %-26s %s' % r)
