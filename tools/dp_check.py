"""2-GPU consistency check of the native gradient exchange (engine.TrainStep, world_size > 1, no DDP wrapper) against
torch DistributedDataParallel: same seeds, same data, 4 steps; prints the per-step losses and the max parameter difference.
Launch: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import core, ops, supernet_config as sc  # noqa: E402
from vit_search_b200.engine import FusedAdamW, TrainStep, broadcast_parameters  # noqa: E402
from vit_search_b200.nets import create_model  # noqa: E402

rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
ops.set_device(local)
dist.init_process_group('nccl')
dev = torch.device('cuda', local)
B = 32
nd, ks = sc.network_def('sr_tiny'), sc.num_channels_to_keep('sr_tiny')
PREC = os.environ.get('VSX_DP_CHECK_PREC', 'fp32')
core.set_precision(PREC)            # fp32: fp32-exact GEMMs, the two exchanges agree to reduction-order rounding; bf16: the training path with the
                                    # per-stage autograd nodes and the STAGED exchange (looser: bf16 rounding of different summation orders)
g = torch.Generator().manual_seed(100 + rank)
x = torch.randn(B, 3, 224, 224, generator=g).to(dev)
t = torch.softmax(torch.randn(B, 1000, generator=g), -1).to(dev)
pt = t.unsqueeze(1).repeat(1, 16, 1).contiguous()
res = {}
for mode in ('native', 'ddp'):
    torch.manual_seed(0)
    m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_classes=1000, drop_path_rate=0.0,
                     num_channels_to_keep=ks, example_per_arch=B, num_warmup_epochs=0, single_arch=True).to(dev)
    m.set_epoch(0)
    m.train()
    net = None
    if mode == 'ddp':
        net = torch.nn.parallel.DistributedDataParallel(m, device_ids=[local], gradient_as_bucket_view=True)
    else:
        broadcast_parameters(m)
    step = TrainStep(m, FusedAdamW(m, lr=1e-3), arch_sample='single', world_size=world, ddp_model=net)
    losses = [step(x, t, pt, epoch=0).item() for _ in range(4)]
    res[mode] = (losses, {k: v.detach().clone() for k, v in m.named_parameters()}, getattr(getattr(step, '_exchange', None), 'calls', None))
if rank == 0:
    print('losses native', res['native'][0])
    print('losses ddp   ', res['ddp'][0])
    worst = max(((res['native'][1][k] - res['ddp'][1][k]).norm() / res['ddp'][1][k].norm().clamp_min(1e-30)).item() for k in res['ddp'][1])
    print('max relative parameter difference after 4 steps: %.3e' % worst)
    print('staged exchange calls in the last native step:', res['native'][2])
    dl = max(abs(a - b) for a, b in zip(res['native'][0], res['ddp'][0]))
    print('max loss difference over the 4 steps: %.3e' % dl)
    # bf16: parameters that start at zero (biases) move by ~lr * sign(g) per step, so their relative difference is dominated by the sign
    # of noise-level gradients; the losses are the meaningful comparison there
    assert (worst < 1e-3) if PREC == 'fp32' else (dl < 5e-3)
dist.destroy_process_group()
