"""RESIDUAL-epilogue GEMM (proj / fc2 forward) under the three tile shapes at the stage shapes of sr_tiny, several reduction lengths."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import _lib, ops  # noqa: E402

lib = _lib.lib()
print('%-34s %10s %10s %12s' % ('shape', '128x128', '256x128', 'pair 256x256'))
for M, C, keep, Ks, rps in ((65792, 256, 224, (64, 128, 192, 256, 384, 512, 768), 257), (16640, 512, 448, (128, 256, 384, 512, 1024, 1536), 65),
                            (4352, 1024, 896, (256, 512, 768, 2048, 3072), 17)):
    Kmax = max(Ks)
    A = torch.randn(M, Kmax, device='cuda').to(torch.bfloat16)
    W = (torch.randn(C, Kmax, device='cuda') * 0.05).to(torch.bfloat16)
    xr, xo = torch.randn(M, C, device='cuda'), torch.empty(M, C, device='cuda')
    bias, rs = torch.zeros(C, device='cuda'), torch.ones(256, device='cuda')
    for K in Ks:
        fn = lambda: ops.gemm(A, W, Kmax, Kmax, M, keep, K, ops.EPI_RESIDUAL, xo, C, n_out=C, aux=xr, ld_aux=C, bias=bias, row_scale=rs,
                              rows_per_sample=rps, n_keep=keep)
        t = []
        for rows, cg in ((128, 1), (256, 1), (0, 2)):
            lib.vsx_gemm_force_tile_rows(rows)
            lib.vsx_gemm_force_cta_group(cg)
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1) * 100)
        lib.vsx_gemm_force_tile_rows(0)
        lib.vsx_gemm_force_cta_group(0)
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print('M=%5d N=%4d (of %4d) K=%4d   %8.1f us %8.1f us %8.1f us   heuristic %6.1f us' % (M, keep, C, K, t[0], t[1], t[2], e0.elapsed_time(e1) * 100))
