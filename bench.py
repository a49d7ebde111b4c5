#!/usr/bin/env python
"""bench.py -- images/sec of one ViT-ResNAS-Tiny supernet TRAIN STEP (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): ViT-ResNAS-Tiny supernet, search space supernet_config/sr_tiny (7/7/4 blocks,
conv stem), 1 sub-architecture per step (`--single-arch` semantics: engine.py:121-122 seeds the draw per iteration so
all ranks agree), batch 256 per GPU, bf16 operands / fp32 accumulate, drop-path 0.2, AdamW -- forward + 2x soft-target CE
+ backward + gradient all-reduce + optimizer step.  Synthetic ImageNet-shaped data, random-init weights.

One JSON line on stdout (rank 0).  `value` = whole-job images/sec with inputs resident in HBM; `e2e` = the same through
the public API with pinned HOST buffers copied in every step and the loss read back; `roofline` = the tensor-core GEMM
kernel's achieved algorithmic TFLOP/s against the measured bf16 peak; `cpu_baseline` = the oracle port on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPACE = 'sr_tiny'
BATCH = 256
DROP_PATH = 0.2
LR, WD = 5e-4, 0.05
METRIC = 'images/sec ViT-ResNAS-Tiny supernet train step'
WORKLOAD = 'ViT-ResNAS-Tiny supernet (supernet_config/sr_tiny) train, 1 arch/step, bs=256/GPU'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--space', default=SPACE)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-batch', type=int, default=8)
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(tflops=p.get('bf16_tflops_sustained', p['bf16_tflops']), hbm=p['hbm_gbs'], src='measured (MEASURED_PEAKS.json, sustained bf16)')
    except Exception:
        return dict(tflops=1400.0, hbm=6650.0, src='fallback (B200_PROFILING.md)')


def gemm_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch (average over the launches of one train step), from the ncu pass
    summarised by tools/summarize_ncu.py into profiles/gemm_traffic.json; None when that capture is missing."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'gemm_traffic.json')) as f:
            return json.load(f)['dram_bytes_per_launch']
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_steps(space, batch, steps, warmup):
    """The reference's algorithm for this path on the host cores: oracle/vit_res_oracle.py (a restatement of the reference's
    PyTorch modules; the reference itself is Python and cannot travel to the GPU box).  fp32, all host threads.
    Returns (img/s, ms/step, cores)."""
    import torch
    from oracle import vit_res_oracle as O
    from vit_search_b200 import supernet_config as sc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nd, ks = sc.network_def(space), sc.num_channels_to_keep(space)
    torch.manual_seed(0)
    p = O.keyed_fill(O.param_shapes(nd))
    p = {k: v.requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in p.items()}
    smp = O.Sampler(nd, ks, batch, 0, single_arch=True)
    smp.set_epoch(0)
    x, t, pt = O.synthetic_batch(batch, seed=1234)
    state, times = {}, []
    names = [k for k, v in p.items() if v.requires_grad]
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        torch.manual_seed(it)
        keeps = smp.sample(batch)
        loss, _, _ = O.train_loss(p, nd, x, t, pt, keeps)
        grads = torch.autograd.grad(loss, [p[k] for k in names])
        with torch.no_grad():
            O.adamw_step({k: p[k] for k in names}, dict(zip(names, grads)), state, LR, WD, it + 1)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, sec * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ips, ms, cores = cpu_steps(args.space, args.cpu_batch, args.steps, args.warmup)
    sample = 'oracle port, fp32, %d host threads, %d-image steps of the same model/space (bounded sample of the 256-image step)' % (cores, args.cpu_batch)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': 'images/sec', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': {'workload': WORKLOAD, 'space': args.space, 'sample_batch': args.cpu_batch},
        'cpu_baseline': {'value': ips, 'unit': 'images/sec', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': ips, 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))


# ---------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, dev):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(dev), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '50'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [v.strip() for v in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if sm:
            out = {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from vit_search_b200 import _lib, core, ops, macs
    from vit_search_b200 import supernet_config as sc
    from vit_search_b200.engine import FusedAdamW, TrainStep
    from vit_search_b200.nets import create_model

    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    ops.set_device(local)
    _lib.check(_lib.lib().vsx_device_ok(local))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl')
    dev = torch.device('cuda', local)
    B = args.batch
    nd, ks = sc.network_def(args.space), sc.num_channels_to_keep(args.space)
    torch.manual_seed(0)
    model = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_classes=1000, drop_rate=0.,
                         drop_path_rate=DROP_PATH, num_channels_to_keep=ks, example_per_arch=B, num_warmup_epochs=0,
                         single_arch=True).to(dev)
    model.set_epoch(0)
    model.train()
    core.set_precision('bf16')
    use_ddp = bool(int(os.environ.get('VSX_BENCH_DDP', '0')))      # 1: torch DistributedDataParallel wrapper instead of the native exchange
    net = model
    if world > 1 and use_ddp:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
    elif world > 1:
        from vit_search_b200.engine import broadcast_parameters
        broadcast_parameters(model)
    opt = FusedAdamW(model, lr=LR * B * world / 512.0, weight_decay=WD)
    step = TrainStep(model, opt, arch_sample='single', world_size=world, ddp_model=net if (world > 1 and use_ddp) else None)

    g = torch.Generator().manual_seed(1234 + rank)
    hx = torch.randn(B, 3, 224, 224, generator=g).pin_memory()
    y = torch.randint(0, 1000, (B,), generator=g)
    ht = torch.full((B, 1000), 0.1 / 1000)
    ht[torch.arange(B), y] += 0.9
    hpt = ht.unsqueeze(1).repeat(1, 16, 1).contiguous().pin_memory()
    ht = ht.pin_memory()
    x, t, pt = hx.to(dev), ht.to(dev), hpt.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def dev_step():
        step(x, t, pt, epoch=0)

    h_loss = torch.empty((), dtype=torch.float32).pin_memory()

    # end to end through the public API: every step uploads its own batch from pinned host memory (engine.DeviceFeeder: side
    # stream, two device slots, so batch i+1 travels while batch i computes) and reads the loss back
    from vit_search_b200.engine import DeviceFeeder
    feeder = DeviceFeeder(dev)

    def e2e_step():
        feeder.submit(hx, ht, hpt)             # batch i+1 (the first call of a window primes the pipeline with an extra submit)
        xs, ts, pts = feeder.next()
        loss = step(xs, ts, pts, epoch=0)
        feeder.release()
        h_loss.copy_(loss, non_blocking=True)

    n_warm = max(args.warmup, 3 if world == 1 else 8)   # DDP rebuilds its buckets after the first backward and the caching
    clocks = ClockSampler(local) if rank == 0 else None   # started before the warm-up: nvidia-smi needs ~0.2 s to deliver its first sample
    for _ in range(n_warm):                              # allocator needs a few steps to settle: extra untimed steps when N > 1
        dev_step()
    ops.LAUNCHES = 0
    launches0 = _lib.lib().vsx_launch_count()
    ms = timed(dev_step, args.steps)
    launches = _lib.lib().vsx_launch_count() - launches0      # kernels launched by libvsx.so (counted in csrc/api.cu: check_launch)
    clk = clocks.stop() if clocks is not None else None
    feeder.submit(hx, ht, hpt)
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)       # K steps = K uploads + K train steps + K loss read-backs inside the window
    feeder.next(), feeder.release()            # drain the batch left in flight
    torch.cuda.synchronize()
    loss_val = float(h_loss)

    # ---- roofline of the dominant kernel: CUDA events around every tensor-core GEMM launch of 2 further steps
    ops.PROFILE = []
    for _ in range(2):
        dev_step()
    torch.cuda.synchronize()
    gemm_ms = sum(r[0].elapsed_time(r[1]) for r in ops.PROFILE)
    gemm_flops = sum(r[2] for r in ops.PROFILE)
    gemm_bytes = sum(r[4] for r in ops.PROFILE)
    n_gemm = len(ops.PROFILE)
    pk0 = peaks()
    # two-resource roofline of the same launches: each launch can be no faster than max(flops / tensor peak, bytes / HBM peak)
    gemm_floor_ms = sum(max(r[2] / (pk0['tflops'] * 1e12), r[4] / (pk0['hbm'] * 1e9)) for r in ops.PROFILE) * 1e3
    n_hbm_bound = sum(1 for r in ops.PROFILE if r[4] / (pk0['hbm'] * 1e9) > r[2] / (pk0['tflops'] * 1e12))
    ops.PROFILE = None
    keeps = model.last_keeps
    step_macs = sum(macs.network_macs(macs.effective_network_def(nd, keeps, b)) for b in range(B))
    pk = peaks()
    ms_step = ms / args.steps
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0

    if rank == 0:
        out = {
            'metric': METRIC, 'value': world * B * args.steps / (ms * 1e-3), 'unit': 'images/sec', 'n_gpus': world, 'steps': args.steps,
            'warmup': n_warm, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'space': args.space, 'batch_per_gpu': B, 'archs_per_step': 1, 'drop_path': DROP_PATH,
                       'optimizer': 'AdamW (fused)', 'parallelism': 'dp%d' % world,
                       'grad_exchange': 'none' if world == 1 else ('torch DDP' if use_ddp else 'in-place NCCL all-reduce of the flat gradient pool, overlapped with the stem backward'),
                       'l2': 'inputs larger than L2: every step streams >10 GB of activations, no tensor survives in the 126 MB L2'},
            'e2e': {'value': world * B * args.steps / (ms_e2e * 1e-3), 'unit': 'images/sec',
                    'h2d_bytes_per_step': hx.numel() * 4 + ht.numel() * 4 + hpt.numel() * 4, 'd2h_bytes_per_step': 4,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches,
            'roofline': {'bound': 'tensor', 'kernel': 'gemm_tc_kernel (tcgen05 GEMM, all epilogues: fwd, dgrad, wgrad)',
                         'achieved': achieved, 'peak': pk['tflops'], 'unit': 'TFLOP/s', 'frac': achieved / pk['tflops'],
                         'traffic': gemm_traffic(), 'algorithmic_bytes_per_launch': gemm_bytes / max(n_gemm, 1),
                         'hbm_view': {'achieved': gemm_bytes / (gemm_ms * 1e-3) / 1e9 if gemm_ms > 0 else 0.0, 'peak': pk['hbm'], 'unit': 'GB/s',
                                      'launches_hbm_bound': n_hbm_bound, 'launches': n_gemm,
                                      'frac_of_two_resource_floor': gemm_floor_ms / gemm_ms if gemm_ms > 0 else 0.0},
                         'peak_source': pk['src'], 'launches_per_step': n_gemm // 2, 'kernel_ms_per_step': gemm_ms / 2,
                         'kernel_share_of_step': (gemm_ms / 2) / ms_step,
                         'how': 'algorithmic 2*M*N*K of the kept extents per launch / CUDA-event time of each launch, 2 instrumented steps after the timed region',
                         'step_algorithmic_tflops': 6.0 * step_macs / (ms_step * 1e-3) / 1e12,
                         'step_frac': 6.0 * step_macs / (ms_step * 1e-3) / 1e12 / pk['tflops']},
            'clocks': clk, 'loss': loss_val,
        }
        if not args.no_cpu_baseline and world == 1:
            ips, cms, cores = cpu_steps(args.space, args.cpu_batch, 3, 1)
            out['cpu_baseline'] = {'value': ips, 'unit': 'images/sec', 'cores': cores, 'kind': 'port',
                                   'sample': 'oracle port (restated reference modules), fp32, %d-image steps x3 of the same model/space, %.0f ms/step' % (args.cpu_batch, cms)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
