"""BASELINE configs[3]: the searched ViT-ResNAS-Medium network (4.6 G MACs; searched_net/medium_mac@4.6G.sh:18) -- dense model, no masks,
drop-path 0.3, EMA on -- one train step at B = 256 on one GPU: step time and a finite loss."""
import time

import torch

from vit_search_b200 import core
from vit_search_b200.engine import FusedAdamW, ModelEma, TrainStep
from vit_search_b200.nets import create_model

from vit_search_b200.supernet_config import VIT_RESNAS_MEDIUM as MEDIUM

core.set_precision('bf16')
B = 256
torch.manual_seed(0)
m = create_model('flexible_vit_sr_patch14_224_patch_output', network_def=MEDIUM, num_classes=1000, drop_path_rate=0.3).cuda()
m.train()
ema = ModelEma(m)
step = TrainStep(m, FusedAdamW(m), model_ema=ema)
x = torch.randn(B, 3, 224, 224, device='cuda')
t = torch.softmax(torch.randn(B, 1000, device='cuda'), -1)
pt = t.unsqueeze(1).repeat(1, 16, 1).contiguous()
for _ in range(3):
    loss = step(x, t, pt)
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    loss = step(x, t, pt)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
w, we = m.blocks[0].attn.qkv.weight, ema.module.blocks[0].attn.qkv.weight
print('medium searched net (4.6 G MACs), dense, EMA on: %.2f ms/step, %.0f img/s, loss %.4f finite=%s, |w - ema|/|w| = %.2e'
      % (ms, B / ms * 1e3, loss.item(), bool(torch.isfinite(loss)), ((w - we).norm() / w.norm()).item()))
