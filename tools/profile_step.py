"""Kernel-time breakdown of the bench train step (torch.profiler / CUPTI) -> gpurun_out/profile_step.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import core, supernet_config as sc  # noqa: E402
from vit_search_b200.engine import FusedAdamW, TrainStep  # noqa: E402
from vit_search_b200.nets import create_model  # noqa: E402

space = sys.argv[1] if len(sys.argv) > 1 else 'sr_tiny'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
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
import time
t0 = time.perf_counter()
for _ in range(5):
    step(x, t, pt)
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step(x, t, pt)
    torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'profile_step.txt')
with open(out, 'w') as f:
    f.write('host launch time per step %.1f ms, wall per step %.1f ms (5 steps)\n' % (t_host / 5 * 1e3, t_all / 5 * 1e3))
    f.write(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=90))
print(open(out).read()[:6000])
# ---- idle-gap analysis from the Kineto trace: for every kernel, gap = start - end of the previous kernel; the gap is
# "host-bound" when the kernel's launch call returned less than 15 us before the kernel started (the GPU was waiting for the CPU)
import json
trace = os.path.join(os.path.dirname(out), 'trace.json')
prof.export_chrome_trace(trace)
ev = json.load(open(trace))['traceEvents']
kern = sorted([e for e in ev if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')], key=lambda e: e['ts'])
launch = {e['args'].get('correlation'): e for e in ev if e.get('cat') == 'cuda_runtime' and 'correlation' in e.get('args', {})}
busy = sum(e['dur'] for e in kern)
span = kern[-1]['ts'] + kern[-1]['dur'] - kern[0]['ts']
gaps, host_gaps, n_host = 0.0, 0.0, 0
hist = {}
for a, b in zip(kern, kern[1:]):
    g = b['ts'] - (a['ts'] + a['dur'])
    if g <= 0:
        continue
    gaps += g
    l = launch.get(b['args'].get('correlation'))
    hb = l is not None and b['ts'] - (l['ts'] + l['dur']) < 15
    if hb:
        host_gaps += g
        n_host += 1
    k = min(int(g // 2) * 2, 20)
    hist[k] = hist.get(k, 0) + 1
with open(out, 'a') as f:
    f.write('\n2 steps: span %.2f ms, busy %.2f ms, idle gaps %.2f ms of which host-bound %.2f ms (%d of %d kernels)\n' % (span / 1e3, busy / 1e3, gaps / 1e3, host_gaps / 1e3, n_host, len(kern)))
    f.write('gap histogram (us bucket: count): %s\n' % sorted(hist.items()))
print(open(out).read()[-600:])
os.remove(trace)
