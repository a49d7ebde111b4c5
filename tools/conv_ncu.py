"""Two launches of every mode of the TMA conv kernels at the train-step size, for `ncu --set full -k regex:tma`."""
import torch

from vit_search_b200 import core, ops

B, H, W, C = 256, 112, 112, 24
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
for _ in range(2):
    ops.call('conv3x3', x, sc, sh, wf, None, out, B, H, W, C, 1, None, None, None, None, None, sums)
    ops.call('conv3x3', x, None, None, wb, None, out, B, H, W, C, 2, yp, gam, bet, mean, rstd, sums)
    ops.call('conv3x3', x, None, None, wb, add, out, B, H, W, C, 2, yp, gam, bet, mean, rstd, sums)
    ops.call('conv3x3_wgrad', x, yp, sc, sh, dw, B, H, W, C)
torch.cuda.synchronize()
