"""Is a train step bound by the launching thread or by the GPU?  Per step: wall time the Python thread needs to ISSUE the step (no
synchronisation inside) next to the CUDA-event time between step boundaries, for 1 and 4 architectures per step.
Usage: PYTHONPATH=. python tools/host_vs_gpu.py [space] [batch]"""
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import _lib, core, supernet_config as sc  # noqa: E402
from vit_search_b200.engine import FusedAdamW, TrainStep  # noqa: E402
from vit_search_b200.nets import create_model  # noqa: E402

space = sys.argv[1] if len(sys.argv) > 1 else 'sr_tiny'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
core.set_precision('bf16')
x = torch.randn(B, 3, 224, 224, device='cuda')
t = torch.softmax(torch.randn(B, 1000, device='cuda'), -1)
pt = t.unsqueeze(1).repeat(1, 16, 1).contiguous()
for archs, mode in ((1, 'single'), (4, 'multi')):
    nd, ks = sc.network_def(space), sc.num_channels_to_keep(space)
    torch.manual_seed(0)
    m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_classes=1000, drop_path_rate=0.2,
                     num_channels_to_keep=ks, example_per_arch=B // archs, num_warmup_epochs=0, single_arch=(mode == 'single')).cuda()
    m.set_epoch(0)
    m.train()
    step = TrainStep(m, FusedAdamW(m), arch_sample=mode)
    for _ in range(4):
        step(x, t, pt)
    torch.cuda.synchronize()
    n = 12
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    host = []
    l0 = _lib.lib().vsx_launch_count()
    ev[0].record()
    for i in range(n):
        t0 = time.perf_counter()
        step(x, t, pt)
        host.append((time.perf_counter() - t0) * 1e3)
        ev[i + 1].record()
    torch.cuda.synchronize()
    gpu = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    # pure host cost: issue ONE step into an empty queue (no back-pressure from a full launch queue), several times
    pure = []
    for _ in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step(x, t, pt)
        pure.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    print('   host-only issue time of one step into an empty queue: p50 %.2f ms (min %.2f)' % (statistics.median(pure), min(pure)))
    print('%s, %d arch(s)/step: host issue p50 %.2f ms, step (events) p50 %.2f ms, libvsx launches/step %d'
          % (space, archs, statistics.median(host), statistics.median(gpu), (_lib.lib().vsx_launch_count() - l0) // n))
    del m, step
    torch.cuda.empty_cache()
