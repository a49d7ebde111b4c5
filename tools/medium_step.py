"""BASELINE configs[3]: the searched ViT-ResNAS-Medium network (4.6 G MACs; searched_net/medium_mac@4.6G.sh:18) -- dense model, no masks,
drop-path 0.3, EMA on -- one train step at B = 256 on one GPU: step time and a finite loss."""
import time

import torch

from vit_search_b200 import core
from vit_search_b200.engine import FusedAdamW, ModelEma, TrainStep
from vit_search_b200.nets import create_model

MEDIUM = ((4, 240), (1, (240, 7, 32), (240, 960), 1), (1, (240, 6, 32), (240, 960), 1), (1, (240, 7, 32), (240, 800), 1),
          (1, (240, 8, 32), (240, 960), 1), (1, (240, 7, 32), (240, 880), 1), (1, (240, 8, 32), (240, 880), 1), (1, (240, 6, 32), (240, 800), 1),
          (3, 240, 640), (1, (640, 10, 48), (640, 1120), 1), (1, (640, 14, 48), (640, 1760), 1), (1, (640, 14, 48), (640, 1920), 1),
          (1, (640, 16, 48), (640, 1760), 1), (1, (640, 14, 48), (640, 1440), 1), (1, (640, 16, 48), (640, 1760), 1), (1, (640, 16, 48), (640, 1920), 1),
          (3, 640, 880), (1, (880, 16, 64), (880, 3200), 1), (1, (880, 10, 64), (880, 3840), 1), (1, (880, 16, 64), (880, 3840), 1),
          (1, (880, 12, 64), (880, 3200), 1), (1, (880, 16, 64), (880, 3520), 1), (1, (880, 14, 64), (880, 3520), 1), (2, 880, 1000))

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
