"""cProfile of the host side of the train step at a tiny batch (GPU work negligible) -> gpurun_out/host_profile.txt"""
import cProfile
import io
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import core, supernet_config as sc  # noqa: E402
from vit_search_b200.engine import FusedAdamW, TrainStep  # noqa: E402
from vit_search_b200.nets import create_model  # noqa: E402

B = 8
nd, ks = sc.network_def('sr_tiny'), sc.num_channels_to_keep('sr_tiny')
m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_classes=1000, drop_path_rate=0.2,
                 num_channels_to_keep=ks, example_per_arch=B, num_warmup_epochs=0, single_arch=True).cuda()
m.set_epoch(0)
m.train()
core.set_precision('bf16')
step = TrainStep(m, FusedAdamW(m), arch_sample='single')
x = torch.randn(B, 3, 224, 224, device='cuda')
t = torch.softmax(torch.randn(B, 1000, device='cuda'), -1)
pt = t.unsqueeze(1).repeat(1, 16, 1).contiguous()
for _ in range(5):
    step(x, t, pt)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step(x, t, pt)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(45)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'host_profile.txt')
open(out, 'w').write(s.getvalue())
print(s.getvalue()[:7000])
