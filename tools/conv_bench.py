"""Time the stem's 3x3 convolution kernels (legacy direct vs TMA warp-specialised) at the train-step size.
Usage: PYTHONPATH=. python tools/conv_bench.py [B]"""
import sys

import torch

from vit_search_b200 import _lib, core, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = W = 112
C = 24
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(B, H, W, C, device='cuda', generator=g).to(torch.bfloat16)
yp = torch.randn(B, H, W, C, device='cuda', generator=g).to(torch.bfloat16)
add = torch.randn(B, H, W, C, device='cuda', generator=g).to(torch.bfloat16)
out = torch.empty_like(x)
sc, sh = torch.rand(C, device='cuda') + 0.5, torch.randn(C, device='cuda') * 0.1
gam, bet, mean, rstd = torch.rand(C, device='cuda') + 0.5, torch.randn(C, device='cuda') * 0.1, torch.randn(C, device='cuda') * 0.1, torch.rand(C, device='cuda') + 0.5
wp = torch.nn.Parameter(torch.randn(C, C, 3, 3, device='cuda') * 0.1)
wf, wb = core.weights.get(wp, 'conv3x3_fwd'), core.weights.get(wp, 'conv3x3_bwd')
sums = torch.zeros(2 * C, device='cuda', dtype=torch.float64)
dw = torch.zeros(C, 9 * C, device='cuda')
mb = x.numel() * 2 / 1e6
modes = {
    'fwd act+stats (2 maps)': (lambda: ops.call('conv3x3', x, sc, sh, wf, None, out, B, H, W, C, 1, None, None, None, None, None, sums), 2),
    'dgrad +yprev (3 maps)': (lambda: ops.call('conv3x3', x, None, None, wb, None, out, B, H, W, C, 2, yp, gam, bet, mean, rstd, sums), 3),
    'dgrad +add +yprev (4 maps)': (lambda: ops.call('conv3x3', x, None, None, wb, add, out, B, H, W, C, 2, yp, gam, bet, mean, rstd, sums), 4),
    'wgrad (2 maps)': (lambda: ops.call('conv3x3_wgrad', x, yp, sc, sh, dw, B, H, W, C), 2),
}
for impl in (1, 2):
    _lib.check(_lib.lib().vsx_conv3x3_force_impl(impl))
    for name, (fn, nmaps) in modes.items():
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print('impl %d  %-28s %7.1f us   %6.2f TB/s algorithmic' % (impl, name, us, nmaps * mb / us))
