"""CUDA-event times of the tcgen05 attention kernels at the three stage shapes of sr_tiny / sr_tiny_mh / sr_small (B = 256).
Usage: PYTHONPATH=. python tools/attn_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import ops  # noqa: E402

B = 256
IMPL = {'tc': ops.ATTN_TCGEN05, 'mma': ops.ATTN_MMA_SYNC}[os.environ.get('VSX_ATTN_BENCH_IMPL', 'tc')]
for N, H, D in ((257, 4, 64), (65, 8, 64), (17, 12, 64), (257, 6, 32), (65, 12, 48), (257, 8, 32)):
    qkv = torch.randn(B * N, 3 * H * D, device='cuda').to(torch.bfloat16)
    do = torch.randn(B * N, H * D, device='cuda').to(torch.bfloat16)
    o = torch.empty(B * N, H * D, device='cuda', dtype=torch.bfloat16)
    lse = torch.zeros(B, H, N, device='cuda')
    dq = torch.empty_like(qkv)
    db = torch.zeros(3 * H * D, device='cuda')
    res = []
    for fn in (lambda: ops.attn_fwd(qkv, o, lse, B, N, H, D, H, D ** -0.5, impl=IMPL), lambda: ops.attn_bwd(qkv, o, do, lse, dq, B, N, H, D, H, D ** -0.5, dbias=db, impl=IMPL)):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 100)
    fl = 4.0 * N * N * D * B * H
    print('N=%3d H=%2d D=%2d: fwd %6.1f us (%5.1f TFLOP/s)   bwd %6.1f us (%5.1f TFLOP/s)' % (N, H, D, res[0], fl / res[0] / 1e6, res[1], 2.5 * fl / res[1] / 1e6))
