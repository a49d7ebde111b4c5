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


def bench_lr(batch, world):
    """The reference scales lr = 5e-4 * global_batch / 512 and reaches it after warm-up epochs (main.py: --lr, --warmup-epochs); a bench of a
    few dozen steps from random init has no warm-up, so the rate is capped at 5e-4 (at 8 x 256 the uncapped 2e-3 diverges to a non-finite
    loss within ~10 steps, which TrainStep.check_finite() reports).  The arithmetic per step does not depend on the value."""
    return min(LR * batch * world / 512.0, LR)
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
    ap.add_argument('--cpu-batch', type=int, default=32)      # the largest sample of the 256-image step the host cores finish in a few seconds
    ap.add_argument('--no-extra', dest='extra', action='store_false')      # skip the extra_configs block (other BASELINE configurations)
    ap.add_argument('--extra-steps', type=int, default=6)
    ap.add_argument('--evo-candidates', type=int, default=64)
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


def bench_config(args, world):
    """The `config` block both arms print: the workload BASELINE.json's metric is quoted on."""
    return {'workload': WORKLOAD, 'space': args.space, 'batch_per_gpu': args.batch, 'archs_per_step': 1, 'drop_path': DROP_PATH,
            'optimizer': 'AdamW', 'parallelism': 'dp%d' % world}


def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores (the oracle port: the reference is Python and does not
    travel to the GPU box), all host threads, each step a BOUNDED sample (--cpu-batch images of the 256-image step)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ips, ms, cores = cpu_steps(args.space, args.cpu_batch, args.steps, args.warmup)
    sample = ('oracle port (restated reference modules), fp32, %d host threads, %d steps (after %d warm-up) of %d images = a bounded sample of the '
              '%d-image step of the same model / space / optimizer' % (cores, args.steps, args.warmup, args.cpu_batch, args.batch))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': 'images/sec', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': dict(bench_config(args, args.gpus), sample_batch=args.cpu_batch),
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


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def _synthetic_batch(B, rank, dev):
    """ImageNet-shaped synthetic batch: pinned host tensors (images uint8 [B,3,224,224] as a decoder produces them AND the fp32 batch
    ToTensor + Normalize makes of them, soft targets [B,1000], per-patch soft targets [B,16,1000] as SwitchTokenMix produces them) and
    device-resident copies of the fp32 images and the targets."""
    import torch
    g = torch.Generator().manual_seed(1234 + rank)
    hu8 = torch.randint(0, 256, (B, 3, 224, 224), generator=g, dtype=torch.uint8).pin_memory()
    mean, std = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1), torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    hx = ((hu8.float() / 255.0 - mean) / std).pin_memory()
    y = torch.randint(0, 1000, (B,), generator=g)
    ht = torch.full((B, 1000), 0.1 / 1000)
    ht[torch.arange(B), y] += 0.9
    hpt = ht.unsqueeze(1).repeat(1, 16, 1).contiguous().pin_memory()
    ht = ht.pin_memory()
    return (hx, ht, hpt, hu8), (hx.to(dev), ht.to(dev), hpt.to(dev))


def _timed_steps(fn, n, world, dev):
    """n calls of fn bracketed by barrier + synchronize on both sides; one CUDA event per step boundary on the compute stream.
    -> (window ms = MAX over ranks, per-step ms list of this rank)."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    barrier()
    ms = torch.tensor([ev[0].elapsed_time(ev[n])], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item(), [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]


def _p50(v):
    return statistics.median(v) if v else None


def _extra_configs(args, world, rank, dev):
    """The other BASELINE.json configurations at full size (few steps each, same --gpus N, device-resident inputs):
    configs[2] sr_small with 4 architectures per step, configs[3] the searched Medium network with EMA, configs[4] forward-only
    evaluation of sampled sub-networks on the resident sr_tiny_mh super-network.  Not the headline: parity cases made measurable."""
    import random
    import torch
    import torch.distributed as dist
    from vit_search_b200 import _lib, core, macs
    from vit_search_b200 import supernet_config as sc
    from vit_search_b200.engine import FusedAdamW, ModelEma, TrainStep, broadcast_parameters
    from vit_search_b200.evo_eval import CandidateEvaluator, sample_candidate
    from vit_search_b200.nets import create_model
    B = args.batch
    out = {}
    _, (x, t, pt) = _synthetic_batch(B, rank, dev)
    W, K = 3, args.extra_steps

    def train_case(name, model, mode, ema=None, note=''):
        model.train()
        if world > 1:
            broadcast_parameters(model)
        opt = FusedAdamW(model, lr=bench_lr(B, world), weight_decay=WD)
        step = TrainStep(model, opt, arch_sample=mode, world_size=world, model_ema=ema)
        for _ in range(W):
            step(x, t, pt, epoch=0)
        l0 = _lib.lib().vsx_launch_count()
        ms, per = _timed_steps(lambda: step(x, t, pt, epoch=0), K, world, dev)
        launches = _lib.lib().vsx_launch_count() - l0
        loss = step(x, t, pt, epoch=0)
        step.check_finite()
        out[name] = {'value': world * B * K / (ms * 1e-3), 'unit': 'images/sec', 'ms_per_step': ms / K, 'ms_per_step_p50': _p50(per),
                     'steps': K, 'warmup': W, 'loss': float(loss), 'gpu_launches_per_step': launches // K, 'config': note}
        del step, opt

    # configs[2]: ViT-ResNAS-Small supernet, multi-arch sampling, 4 architectures per step
    nd, ks = sc.network_def('sr_small'), sc.num_channels_to_keep('sr_small')
    torch.manual_seed(0)
    m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_classes=1000, drop_rate=0., drop_path_rate=0.3,
                     num_channels_to_keep=ks, example_per_arch=B // 4, num_warmup_epochs=0, single_arch=False).to(dev)
    m.set_epoch(0)
    train_case('sr_small_4archs', m, 'multi', note='ViT-ResNAS-Small supernet (sr_small), multi-arch sampling, 4 archs/step, bs=%d/GPU, bf16, drop-path 0.3' % B)
    del m
    # the headline space with 4 architectures per step (north star: the multi-architectural-sampling inner loop)
    nd, ks = sc.network_def('sr_tiny'), sc.num_channels_to_keep('sr_tiny')
    torch.manual_seed(0)
    m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_classes=1000, drop_rate=0., drop_path_rate=DROP_PATH,
                     num_channels_to_keep=ks, example_per_arch=B // 4, num_warmup_epochs=0, single_arch=False).to(dev)
    m.set_epoch(0)
    train_case('sr_tiny_4archs', m, 'multi', note='ViT-ResNAS-Tiny supernet (sr_tiny), multi-arch sampling, 4 archs/step, bs=%d/GPU, bf16' % B)
    del m
    # configs[3]: searched ViT-ResNAS-Medium network (4.6 G MACs), dense, EMA on
    torch.manual_seed(0)
    m = create_model('flexible_vit_sr_patch14_224_patch_output', network_def=sc.VIT_RESNAS_MEDIUM, num_classes=1000, drop_path_rate=0.3).to(dev)
    ema = ModelEma(m)
    train_case('medium_4g6_dense_ema', m, None, ema=ema,
               note='searched ViT-ResNAS-Medium network_def (4.6 G MACs, %.2f G counted), dense, EMA on, bs=%d/GPU, bf16, drop-path 0.3' % (macs.network_macs(sc.VIT_RESNAS_MEDIUM) / 1e9, B))
    del m, ema
    torch.cuda.empty_cache()
    # configs[4]: evolutionary-search candidate evaluation, forward only, on the RESIDENT super-network weights
    nd, ks = sc.network_def('sr_tiny_mh'), sc.num_channels_to_keep('sr_tiny_mh')
    torch.manual_seed(0)
    m = create_model('flexible_vit_sr_patch14_224_patch_output', network_def=nd, num_classes=1000).to(dev).eval()
    if world > 1:
        broadcast_parameters(m)
    evl = CandidateEvaluator(m, dev)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    loader = [(torch.randn(B, 3, 224, 224, device=dev, generator=g), torch.randint(0, 1000, (B,), device=dev, generator=g)) for _ in range(2)]
    rng = random.Random(0)                         # every rank scores the same candidates on its own shard of the sub-val set
    cands = [sample_candidate(nd, ks, rng) for _ in range(args.evo_candidates)]
    evl.score(cands[0], loader)
    evl.score(nd, loader)
    state = {'i': 0, 'acc': []}

    def one():
        state['acc'].append(evl.score(cands[state['i']], loader)['acc1'])
        state['i'] += 1
    ms, per = _timed_steps(one, len(cands), world, dev)
    n_img = world * len(cands) * len(loader) * B
    out['evo_eval_sr_tiny_mh'] = {'value': n_img / (ms * 1e-3), 'unit': 'images/sec', 'candidates': len(cands), 'images_per_candidate': world * len(loader) * B,
                                  'ms_per_candidate': ms / len(cands), 'ms_per_candidate_p50': _p50(per),
                                  'config': 'evolutionary-search eval: %d sampled sub-nets (uniform draws from sr_tiny_mh), forward only, synthetic sub-val of %d images per candidate, resident '
                                            'super-network weights, CE / top-1 / top-5 meters summed over ranks once per candidate' % (len(cands), world * len(loader) * B)}
    del m, evl
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import gc
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
    _lib.check(_lib.lib().vsx_device_ok(local))
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl')
        warm = torch.ones(1 << 20, device=dev)          # NCCL communicator / channel set-up is not a train step: do it before the warm-up
        dist.all_reduce(warm)
        torch.cuda.synchronize()
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
    opt = FusedAdamW(model, lr=bench_lr(B, world), weight_decay=WD)
    step = TrainStep(model, opt, arch_sample='single', world_size=world, ddp_model=net if (world > 1 and use_ddp) else None)
    (hx, ht, hpt, hu8), (x, t, pt) = _synthetic_batch(B, rank, dev)
    step_macs = []

    def dev_step():
        step(x, t, pt, epoch=0)
        step_macs.append(model.last_keeps)

    h_loss = torch.empty((), dtype=torch.float32).pin_memory()

    # end to end through the public API: every step uploads its own batch from pinned host memory (engine.DeviceFeeder: side
    # stream, two device slots, so batch i+1 travels while batch i computes) and reads the loss back.  Images travel as uint8 (what a
    # decoder produces: 1 byte per pixel) and ToTensor + Normalize runs on the device (vsx_image_normalize_u8); the fp32-image variant
    # (the reference's engine.py:104-105 uploads 4 bytes per pixel) is timed as well.
    from vit_search_b200.engine import DeviceFeeder
    feeder = DeviceFeeder(dev, normalize=(IMAGENET_MEAN, IMAGENET_STD))
    feeder32 = DeviceFeeder(dev)

    def make_e2e(fd, himg):
        def e2e_step():
            fd.submit(himg, ht, hpt)               # batch i+1 (the first call of a window primes the pipeline with an extra submit)
            xs, ts, pts = fd.next()
            loss = step(xs, ts, pts, epoch=0)
            fd.release()
            h_loss.copy_(loss, non_blocking=True)
        return e2e_step
    e2e_step, e2e_step32 = make_e2e(feeder, hu8), make_e2e(feeder32, hx)

    n_warm = max(args.warmup, 3)                          # exactly the driver's W (the contract asks for W >= 3)
    clocks = ClockSampler(local) if rank == 0 else None   # started before the warm-up: nvidia-smi needs ~0.2 s to deliver its first sample
    for _ in range(n_warm):
        dev_step()
    gc.collect()
    gc.disable()                                          # a collector pause of the launching thread inside the window is not a property of the step
    step_macs.clear()
    launches0 = _lib.lib().vsx_launch_count()
    ms, per_step = _timed_steps(dev_step, args.steps, world, dev)
    launches = _lib.lib().vsx_launch_count() - launches0      # kernels launched by libvsx.so (counted in csrc/api.cu: check_launch)
    timed_keeps = list(step_macs)
    clk = clocks.stop() if clocks is not None else None
    feeder.submit(hx, ht, hpt)
    e2e_step()
    ms_e2e, per_step_e2e = _timed_steps(e2e_step, args.steps, world, dev)       # K steps = K uploads + K train steps + K loss read-backs
    feeder.next(), feeder.release()            # drain the batch left in flight
    feeder32.submit(hx, ht, hpt)
    e2e_step32()
    ms_e2e32, _ = _timed_steps(e2e_step32, args.steps, world, dev)
    feeder32.next(), feeder32.release()
    torch.cuda.synchronize()
    gc.enable()
    loss_val = float(h_loss)
    step.check_finite()

    # ---- roofline of the dominant kernel: CUDA events around every tensor-core GEMM launch of 2 further steps
    ops.PROFILE = []
    for _ in range(2):
        # the instrumented step is issued from Python launch by launch (slower than the GPU): park the stream behind a ~60 ms spin kernel
        # first, so that the whole step is queued before the first event is processed and an interval measures the kernel (plus its
        # launch latency on the device), not the wait for the launching thread
        torch.cuda.synchronize()
        torch.cuda._sleep(int(0.06 * 1.9e9))
        dev_step()
    torch.cuda.synchronize()
    gemm_ms = sum(r[0].elapsed_time(r[1]) for r in ops.PROFILE)
    gemm_flops = sum(r[2] for r in ops.PROFILE)
    gemm_bytes = sum(r[4] for r in ops.PROFILE)
    n_gemm = len(ops.PROFILE)
    pk = peaks()
    # two-resource roofline of the same launches: each launch can be no faster than max(flops / tensor peak, bytes / HBM peak)
    gemm_floor_ms = sum(max(r[2] / (pk['tflops'] * 1e12), r[4] / (pk['hbm'] * 1e9)) for r in ops.PROFILE) * 1e3
    n_hbm_bound = sum(1 for r in ops.PROFILE if r[4] / (pk['hbm'] * 1e9) > r[2] / (pk['tflops'] * 1e12))
    ops.PROFILE = None
    # algorithmic MACs of the TIMED steps (every step samples its own sub-network): mean over exactly those steps
    mean_macs = statistics.mean(B * macs.network_macs(macs.effective_network_def(nd, k, 0)) for k in timed_keeps) if timed_keeps else 0.0
    ms_step = ms / args.steps
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    extra = _extra_configs(args, world, rank, dev) if args.extra else None

    if rank == 0:
        out = {
            'metric': METRIC, 'value': world * B * args.steps / (ms * 1e-3), 'unit': 'images/sec', 'n_gpus': world, 'steps': args.steps,
            'warmup': n_warm, 'ms_per_step': ms_step, 'ms_per_step_p50': _p50(per_step), 'ms_per_step_min': min(per_step), 'ms_per_step_max': max(per_step),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16', 'data': 'synthetic',
            'config': dict(bench_config(args, world),
                       grad_exchange= 'none' if world == 1 else ('torch DDP' if use_ddp else 'in-place NCCL all-reduce of the flat gradient pool, staged per stage and overlapped with the backward'),
                       l2='inputs larger than L2: every step streams >10 GB of activations, no tensor survives in the 126 MB L2',
                       timing='one CUDA event per step boundary; ms_per_step = window / K (the value), ms_per_step_p50 = median of the K step times of rank 0'),
            'e2e': {'value': world * B * args.steps / (ms_e2e * 1e-3), 'unit': 'images/sec',
                    'h2d_bytes_per_step': hu8.numel() + ht.numel() * 4 + hpt.numel() * 4, 'd2h_bytes_per_step': 4,
                    'ms_per_step': ms_e2e / args.steps, 'ms_per_step_p50': _p50(per_step_e2e),
                    'input': 'uint8 images [B,3,224,224] + fp32 soft targets from pinned host memory; ToTensor + Normalize on the device',
                    'fp32_image_input': {'value': world * B * args.steps / (ms_e2e32 * 1e-3), 'ms_per_step': ms_e2e32 / args.steps,
                                         'h2d_bytes_per_step': hx.numel() * 4 + ht.numel() * 4 + hpt.numel() * 4}},
            'gpu_launches': launches,
            'roofline': {'bound': 'tensor', 'kernel': 'gemm_tc_kernel (tcgen05 GEMM, all epilogues: fwd, dgrad, wgrad; single-CTA and cta_group::2 tiles)',
                         'achieved': achieved, 'peak': pk['tflops'], 'unit': 'TFLOP/s', 'frac': achieved / pk['tflops'],
                         'traffic': gemm_traffic(), 'algorithmic_bytes_per_launch': gemm_bytes / max(n_gemm, 1),
                         'hbm_view': {'achieved': gemm_bytes / (gemm_ms * 1e-3) / 1e9 if gemm_ms > 0 else 0.0, 'peak': pk['hbm'], 'unit': 'GB/s',
                                      'launches_hbm_bound': n_hbm_bound, 'launches': n_gemm,
                                      'frac_of_two_resource_floor': gemm_floor_ms / gemm_ms if gemm_ms > 0 else 0.0},
                         'peak_source': pk['src'], 'launches_per_step': n_gemm // 2, 'kernel_ms_per_step': gemm_ms / 2,
                         'kernel_share_of_step': (gemm_ms / 2) / ms_step,
                         'how': 'algorithmic 2*M*N*K of the kept extents per launch / CUDA-event time of each launch (events on the launching stream around every GEMM launch), 2 instrumented steps after the timed region, each queued behind a spin kernel so that the intervals do not contain waits for the launching thread',
                         'step_algorithmic_tflops': 6.0 * mean_macs / (ms_step * 1e-3) / 1e12,
                         'step_frac': 6.0 * mean_macs / (ms_step * 1e-3) / 1e12 / pk['tflops'],
                         'step_macs_how': 'mean algorithmic MACs of the sub-networks sampled in the K timed steps'},
            'clocks': clk, 'loss': loss_val,
        }
        if extra is not None:
            out['extra_configs'] = extra
        if not args.no_cpu_baseline and world == 1:
            ips, cms, cores = cpu_steps(args.space, args.cpu_batch, 2, 1)
            out['cpu_baseline'] = {'value': ips, 'unit': 'images/sec', 'cores': cores, 'kind': 'port',
                                   'sample': 'oracle port (restated reference modules), fp32, %d-image steps x2 (after 1 warm-up) of the same model/space, %.0f ms/step' % (args.cpu_batch, cms)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
