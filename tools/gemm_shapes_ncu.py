"""The stage-2 / stage-3 GEMM shapes of the bench train step (sr_tiny, B = 256), two launches each, for `ncu --set full -k regex:gemm_tc`.
Shapes: (M, N, K, epilogue, b_layout) as tools/gemm_table.py prints them; the launcher's own heuristic picks single-CTA or CTA-pair tiles."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import ops  # noqa: E402

SHAPES = [
    (4352, 896, 3072, 'resid', 0),      # stage 3 fc2 (K = 3072)
    (4352, 896, 3072, 'store', 1),      # stage 3 fc2 data gradient
    (4352, 3072, 896, 'gelu', 0),       # stage 3 fc1
    (4352, 2304, 896, 'store', 0),      # stage 3 qkv
    (16640, 384, 1536, 'resid', 0),     # stage 2 fc2
    (16640, 1536, 384, 'gelu', 0),      # stage 2 fc1
    (16640, 1152, 384, 'store', 0),     # stage 2 qkv
    (65792, 768, 224, 'store', 0),      # stage 1 qkv (for contrast: epilogue bound)
]
for M, N, K, kind, bl in SHAPES:
    A = torch.randn(M, K, device='cuda').to(torch.bfloat16)
    W = (torch.randn(N, K, device='cuda') * 0.05).to(torch.bfloat16) if bl == 0 else (torch.randn(K, N, device='cuda') * 0.05).to(torch.bfloat16)
    ldb = K if bl == 0 else N
    bias = torch.zeros(N, device='cuda')
    for _ in range(2):
        if kind == 'store':
            out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
            ops.gemm(A, W, K, ldb, M, N, K, ops.EPI_STORE, out, N, b_layout=bl, bias=bias)
        elif kind == 'gelu':
            out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
            out2 = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
            ops.gemm(A, W, K, ldb, M, N, K, ops.EPI_GELU, out, N, out2=out2, ldo2=N, bias=bias)
        else:
            C = (N + 127) // 128 * 128
            res = torch.randn(M, C, device='cuda')
            out = torch.empty(M, C, device='cuda')
            ops.gemm(A, W, K, ldb, M, N, K, ops.EPI_RESIDUAL, out, C, n_out=C, bias=bias, aux=res, ld_aux=C, rows_per_sample=1, n_keep=N)
    torch.cuda.synchronize()
print('done')
