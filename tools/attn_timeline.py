"""Per-block timeline of the tcgen05 attention backward kernel (CTA 0): clock64 stamps of the MMA thread and one softmax warp."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import _lib, ops  # noqa: E402

B, N, H, D = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 257, 4, 64
qkv = torch.randn(B * N, 3 * H * D, device='cuda').to(torch.bfloat16)
do = torch.randn(B * N, H * D, device='cuda').to(torch.bfloat16)
o = torch.empty(B * N, H * D, device='cuda', dtype=torch.bfloat16)
lse = torch.zeros(B, H, N, device='cuda')
dq = torch.empty_like(qkv)
ops.attn_fwd(qkv, o, lse, B, N, H, D, H, D ** -0.5)
dbg = torch.zeros(64 * 8, dtype=torch.int64, device='cuda')
for it in range(3):
    ops.attn_bwd(qkv, o, do, lse, dq, B, N, H, D, H, D ** -0.5)
_lib.lib().vsx_attn_debug_buffer(dbg.data_ptr())
ops.attn_bwd(qkv, o, do, lse, dq, B, N, H, D, H, D ** -0.5)
torch.cuda.synchronize()
_lib.lib().vsx_attn_debug_buffer(None)
t = dbg.view(64, 8).cpu()
t0 = int(t[0, 3])
print('block | mma: pds_ready  sdp(next) issued  dvdkdq issued | softmax: enter-wait  sdp_full  staged  arrived   (cycles since start)')
for g in range(48):
    r = [int(v) - t0 if int(v) else -1 for v in t[g, :7]]
    print('%3d | %8d %8d %8d | %8d %8d %8d %8d   softmax work %5d  wait %5d' % (g, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[5] - r[4], r[4] - r[3]))
