"""Timeline of the tcgen05 attention backward kernel: per-CTA start / end / SM (globaltimer) and, for one CTA (the slowest of a first
instrumented launch), clock64 stamps of the MMA thread, one softmax warp, the epilogue warps and the odd-token side warps.
Usage: python tools/attn_timeline.py [batch] [tokens]   (VSX_ATTN_ODD selects the odd-token mode, see csrc/attn_tc.cu)"""
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
for it in range(3):
    ops.attn_bwd(qkv, o, do, lse, dq, B, N, H, D, H, D ** -0.5)


def run(cta_index):
    os.environ['VSX_ATTN_DBG_CTA'] = str(cta_index)
    dbg = torch.zeros(64 * 16 + 3 * 160, dtype=torch.int64, device='cuda')
    _lib.lib().vsx_attn_debug_buffer(dbg.data_ptr())
    ops.attn_bwd(qkv, o, do, lse, dq, B, N, H, D, H, D ** -0.5)
    torch.cuda.synchronize()
    _lib.lib().vsx_attn_debug_buffer(None)
    return dbg[:1024].view(64, 16).cpu(), dbg[1024:1024 + 2 * 148].view(148, 2).cpu(), dbg[1024 + 320:1024 + 320 + 148].cpu()


_, cta, smid = run(0)
dur = (cta[:, 1] - cta[:, 0]).float() / 1e3
order = torch.argsort(dur)
print('per-CTA duration (us): min %.1f  median %.1f  max %.1f' % (dur.min(), dur.median(), dur.max()))
print('fastest: ' + '  '.join('cta %d sm %d %.1f' % (int(i), int(smid[i]), dur[i]) for i in order[:8]))
print('slowest: ' + '  '.join('cta %d sm %d %.1f' % (int(i), int(smid[i]), dur[i]) for i in order[-8:]))
for which in (int(order[0]), int(order[-1])):
    t, cta2, _ = run(which)
    print('---- CTA %d (this launch: %.1f us)' % (which, (int(cta2[which, 1]) - int(cta2[which, 0])) / 1e3))
    t0 = int(t[0, 3])
    print('block | mma: pds_ready  sdp(next) issued  accumulators free  dvdkdq issued | softmax: enter-wait  sdp_full  staged  arrived   (cycles)')
    for g in range(44):
        r = [int(v) - t0 if int(v) else -1 for v in t[g, :8]]
        print('%3d | %8d %8d %8d %8d | %8d %8d %8d %8d   softmax work %5d  wait %5d' % (g, r[0], r[1], r[7], r[2], r[3], r[4], r[5], r[6], r[5] - r[4], r[4] - r[3]))
    print('key block | epilogue: wait-enter  dkv_full  side_ready  drained | side warps: wait-enter  kv_full  block done')
    for g in range(22):
        r = [int(v) - t0 if int(v) else -1 for v in t[g, 8:15]]
        print('%3d | %8d %8d %8d %8d | %8d %8d %8d' % (g, r[0], r[1], r[2], r[3], r[4], r[5], r[6]))
