"""One bench-configuration train step (no warm-up) -- the target of the ncu captures under profiles/."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import core, supernet_config as sc  # noqa: E402
from vit_search_b200.engine import FusedAdamW, TrainStep  # noqa: E402
from vit_search_b200.nets import create_model  # noqa: E402

space = sys.argv[1] if len(sys.argv) > 1 else 'sr_tiny'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
archs = int(sys.argv[4]) if len(sys.argv) > 4 else 1      # architectures per step (1: --single-arch semantics; > 1: multi-arch sampling)
nd, ks = sc.network_def(space), sc.num_channels_to_keep(space)
torch.manual_seed(0)
m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_classes=1000, drop_path_rate=0.2,
                 num_channels_to_keep=ks, example_per_arch=B // archs, num_warmup_epochs=0, single_arch=(archs == 1)).cuda()
m.set_epoch(0)
m.train()
core.set_precision('bf16')
step = TrainStep(m, FusedAdamW(m), arch_sample='single' if archs == 1 else 'multi')
x = torch.randn(B, 3, 224, 224, device='cuda')
t = torch.softmax(torch.randn(B, 1000, device='cuda'), -1)
pt = t.unsqueeze(1).repeat(1, 16, 1).contiguous()
for _ in range(steps):
    loss = step(x, t, pt)
torch.cuda.synchronize()
print('loss', float(loss))
