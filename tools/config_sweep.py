"""Run the train step at full size for the other BASELINE.json configurations (parity-test cases, not bench lines): step time and a
finite loss for sr_small 4 archs/step (multi), sr_tiny_mh (head_dim 32/48/64: mma.sync attention path), sr_tiny multi 4 archs/step."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import core, supernet_config as sc  # noqa: E402
from vit_search_b200.engine import FusedAdamW, TrainStep  # noqa: E402
from vit_search_b200.nets import create_model  # noqa: E402

core.set_precision('bf16')
B = 256
x = torch.randn(B, 3, 224, 224, device='cuda')
t = torch.softmax(torch.randn(B, 1000, device='cuda'), -1)
pt = t.unsqueeze(1).repeat(1, 16, 1).contiguous()
for space, epa, mode, dp in (('sr_small', 64, 'multi', 0.3), ('sr_tiny', 64, 'multi', 0.2), ('sr_tiny_mh', 256, 'single', 0.2), ('sr_small', 256, 'single', 0.3)):
    nd, ks = sc.network_def(space), sc.num_channels_to_keep(space)
    torch.manual_seed(0)
    m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_classes=1000, drop_path_rate=dp,
                     num_channels_to_keep=ks, example_per_arch=epa, num_warmup_epochs=0, single_arch=(mode == 'single')).cuda()
    m.set_epoch(0)
    m.train()
    step = TrainStep(m, FusedAdamW(m), arch_sample=mode)
    for _ in range(3):
        loss = step(x, t, pt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        loss = step(x, t, pt)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    print('%-11s archs/step %d (%s): %.2f ms/step, %.0f img/s, loss %.4f finite=%s' % (space, B // epa, mode, ms, B / ms * 1e3, loss.item(), bool(torch.isfinite(loss))))
    del m, step
    torch.cuda.empty_cache()
