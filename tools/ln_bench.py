"""CUDA-event times of the masked LayerNorm kernels at the stage shapes of sr_tiny (B = 256): forward (fp32 rows -> bf16) and backward
(bf16 dy + fp32 x + fp32 gradient stream in -> fp32 gradient stream out + the fused bf16 cast), against their HBM floors."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import ops  # noqa: E402

for rows, C, keep in ((256 * 257, 256, 224), (256 * 65, 512, 448), (256 * 17, 1024, 896)):
    x = torch.randn(rows, C, device='cuda')
    g = torch.ones(C, device='cuda')
    b = torch.zeros(C, device='cuda')
    y = torch.empty(rows, C, device='cuda', dtype=torch.bfloat16)
    mean, rstd = torch.empty(rows, device='cuda'), torch.empty(rows, device='cuda')
    dy = torch.randn(rows, C, device='cuda').to(torch.bfloat16)
    gi, go = torch.randn(rows, C, device='cuda'), torch.empty(rows, C, device='cuda')
    dg, db = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    spare = [torch.randn(rows, C, device='cuda') for _ in range(max(1, int(3e8 // (rows * C * 4))))]      # > L2 between launches
    res = []
    for fn in (lambda i: ops.masked_ln_fwd(spare[i % len(spare)], C, g, b, y, C, mean, rstd, rows, C, keep, 1e-6),
               lambda i: ops.masked_ln_bwd(dy, C, spare[i % len(spare)], C, mean, rstd, g, gi, go, C, dg, db, rows, C, keep)):
        for i in range(3):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 50)
    fb, bb = rows * keep * 6, rows * keep * 14
    print('rows %6d C %4d keep %4d: fwd %6.1f us (%4.2f TB/s)   bwd %6.1f us (%4.2f TB/s)' % (rows, C, keep, res[0], fb / res[0] / 1e6, res[1], bb / res[1] / 1e6))
