"""Per-unit timeline (CTA 0) of one tcgen05 GEMM launch: clock64 stamps of the epilogue warp, the store warp and the MMA thread."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import _lib, ops  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else 'gelugrad'
M, N, K = 65792, 768, 224
A = torch.randn(M, 256, device='cuda').to(torch.bfloat16)
W = (torch.randn(768, 256, device='cuda') * 0.05).to(torch.bfloat16)       # fc1 weight [hidden, C] / used as W2^T for the dgrad
u = torch.randn(M, 768, device='cuda').to(torch.bfloat16)
out = torch.empty(M, 768, device='cuda', dtype=torch.bfloat16)
out2 = torch.empty(M, 768, device='cuda', dtype=torch.bfloat16)
bias = torch.zeros(768, device='cuda')
cs = torch.zeros(768, device='cuda')


xr = torch.randn(M, 256, device='cuda')
xo = torch.empty(M, 256, device='cuda')
Wp = (torch.randn(256, 256, device='cuda') * 0.05).to(torch.bfloat16)
rs = torch.ones(256, device='cuda')


def run():
    if kind == 'resid':      # proj / fc2 of stage 1: x_new = x + row_scale * mask * (a W^T + b), fp32 residual stream in and out
        ops.gemm(A, Wp, 256, 256, M, 224, 256, ops.EPI_RESIDUAL, xo, 256, n_out=256, aux=xr, ld_aux=256, bias=bias, row_scale=rs,
                 rows_per_sample=257, n_keep=224)
    elif kind == 'gelugrad':
        ops.gemm(A, W, 256, 256, M, N, K, ops.EPI_GELUGRAD, out, 768, n_out=768, aux=u, ld_aux=768, colsum=cs)
    elif kind == 'gelu':
        ops.gemm(A, W, 256, 256, M, N, K, ops.EPI_GELU, out, 768, n_out=768, out2=out2, ldo2=768, bias=bias)
    else:
        ops.gemm(A, W, 256, 256, M, N, K, ops.EPI_STORE, out, 768, n_out=768, bias=bias)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print(kind, 'us per launch', e0.elapsed_time(e1) * 100)
dbg = torch.zeros(64 * 8, dtype=torch.int64, device='cuda')
_lib.lib().vsx_gemm_debug_buffer(dbg.data_ptr())
run()
torch.cuda.synchronize()
_lib.lib().vsx_gemm_debug_buffer(None)
t = dbg.view(64, 8).cpu()
t0 = int(t[0, 0])
print('unit | epi: wait-ready  ready   math-done  staged | acc_full(tile) | store: wait-staged  issued | mma tile issued')
for un in range(44):
    r = [int(v) - t0 if int(v) else -1 for v in t[un]]
    print('%3d | %8d %8d %8d %8d | %8d | %8d %8d | %8d' % (un, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]))
