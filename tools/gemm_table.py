"""Per-shape time table of every tensor-core GEMM launch in one bench train step -> gpurun_out/gemm_table.txt"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import core, ops, supernet_config as sc  # noqa: E402
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
# VSX_GEMM_TILE=128 / 256 forces the single-CTA tile rows (with VSX_GEMM_CTA_GROUP=1), for per-shape comparisons of the three tile shapes
if os.environ.get('VSX_GEMM_TILE'):
    from vit_search_b200 import _lib
    _lib.lib().vsx_gemm_force_tile_rows(int(os.environ['VSX_GEMM_TILE']))
    for _ in range(2):
        step(x, t, pt)
ops.PROFILE = []
step(x, t, pt)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for e0, e1, fl, key, _nbytes in ops.PROFILE:
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
    a[2] += fl
EPI = ['STORE', 'GELU', 'RESID', 'GELUGRAD', 'ATOMIC']
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', os.environ.get('VSX_GEMM_TABLE_OUT', 'gemm_table.txt'))
with open(out, 'w') as f:
    tot = sum(v[1] for v in agg.values())
    f.write('total GEMM ms/step %.3f over %d launches, %.1f TFLOP/s\n' % (tot, sum(v[0] for v in agg.values()), sum(v[2] for v in agg.values()) / tot / 1e9))
    f.write('%8s %6s %8s %-8s %2s %2s %5s %4s | %4s %9s %9s %9s\n' % ('M', 'N', 'K', 'epi', 'aL', 'bL', 'n_out', 'splt', 'cnt', 'ms_total', 'us_each', 'TFLOP/s'))
    for (M, N, K, epi, al, bl, n_out, sk, terms), (cnt, ms, fl) in rows:
        f.write('%8d %6d %8d %-8s %2d %2d %5d %4d | %4d %9.3f %9.1f %9.1f\n' % (M, N, K, EPI[epi], al, bl, n_out, sk, cnt, ms, ms / cnt * 1e3, fl / ms / 1e9))
print(open(out).read()[:300])
