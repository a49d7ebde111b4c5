"""Stage-1 GEMM shapes (K <= 768: HBM / epilogue bound) under the three tile shapes: 128 x 128 single CTA, 256 x 128 single CTA, 256 x 256 CTA pair.
Wave quantisation: 65792 rows = 257 tiles of 256 rows = 3.47 waves of 74 pairs / 148 SMs, but 514 x (N / 128) tiles of 128 rows."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import _lib, ops  # noqa: E402

M = 65792
A256 = torch.randn(M, 256, device='cuda').to(torch.bfloat16)
A768 = torch.randn(M, 768, device='cuda').to(torch.bfloat16)
W = (torch.randn(768, 768, device='cuda') * 0.05).to(torch.bfloat16)
xr, xo = torch.randn(M, 256, device='cuda'), torch.empty(M, 256, device='cuda')
o768 = torch.empty(M, 768, device='cuda', dtype=torch.bfloat16)
o768b = torch.empty(M, 768, device='cuda', dtype=torch.bfloat16)
o256 = torch.empty(M, 256, device='cuda', dtype=torch.bfloat16)
bias = torch.zeros(768, device='cuda')
rs = torch.ones(256, device='cuda')
cs = torch.zeros(768, device='cuda')
cases = {
    'qkv  STORE  N=768 K=224': lambda: ops.gemm(A256, W, 256, 768, M, 768, 224, ops.EPI_STORE, o768, 768, bias=bias),
    'proj RESID  N=224 K=256': lambda: ops.gemm(A256, W, 256, 768, M, 224, 256, ops.EPI_RESIDUAL, xo, 256, n_out=256, aux=xr, ld_aux=256, bias=bias,
                                                row_scale=rs, rows_per_sample=257, n_keep=224),
    'fc1  GELU   N=768 K=224': lambda: ops.gemm(A256, W, 256, 768, M, 768, 224, ops.EPI_GELU, o768, 768, out2=o768b, ldo2=768, bias=bias),
    'fc2  RESID  N=224 K=768': lambda: ops.gemm(A768, W, 768, 768, M, 224, 768, ops.EPI_RESIDUAL, xo, 256, n_out=256, aux=xr, ld_aux=256, bias=bias,
                                                row_scale=rs, rows_per_sample=257, n_keep=224),
    'dfc2 GELUG  N=768 K=224': lambda: ops.gemm(A256, W, 256, 768, M, 768, 224, ops.EPI_GELUGRAD, o768, 768, aux=o768b, ld_aux=768, colsum=cs,
                                                b_layout=ops.MNMAJOR),
    'dqkv STORE  N=224 K=768': lambda: ops.gemm(A768, W, 768, 768, M, 224, 768, ops.EPI_STORE, o256, 256, b_layout=ops.MNMAJOR),
    'dprj STORE  N=256 K=224': lambda: ops.gemm(A256, W, 256, 768, M, 256, 224, ops.EPI_STORE, o256, 256, b_layout=ops.MNMAJOR),
}
lib = _lib.lib()
print('%-26s %10s %10s %10s' % ('shape', '128x128', '256x128', 'pair 256x256'))
for name, fn in cases.items():
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
    print('%-26s %8.1f us %8.1f us %8.1f us' % (name, t[0], t[1], t[2]))
