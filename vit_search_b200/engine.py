"""The train step of the ViT-Res super-network -- the body of the reference's engine.train_one_epoch loop
(engine.py:101-185) re-built around the libvsx kernels: RNG bracket for architecture sampling (:119-131, :164-165),
forward, the two soft-target cross-entropies (:152-157), backward, data-parallel gradient all-reduce (DDP at
main.py:366-368), fused AdamW (:175-177).  Compared with the reference there is no host sync between forward and
backward (`loss.item()` at :168 becomes a device-side scalar read once per logging interval) and the optimizer is one
kernel launch instead of a per-tensor loop.
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib, core, ops

_SAMPLE_EPOCH_OFFSET = 10000   # engine.py:98


# ------------------------------------------------------------------------------------------------ loss
class _SoftCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        core.require_cuda(logits, 'SoftTargetCrossEntropy')
        K = logits.shape[-1]
        x = logits.contiguous().view(-1, K).float()
        t = target.contiguous().view(-1, K).float()
        rows = x.shape[0]
        loss = torch.zeros((), device=x.device)
        need_grad = logits.requires_grad
        dl = torch.empty_like(x) if need_grad else None
        ops.call('soft_ce', x, K, t, K, rows, K, 1.0 / rows, 1.0 / rows, loss, dl, K)
        ctx.dl, ctx.shape = dl, logits.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        dl = ctx.dl
        ctx.dl = None
        ops.call('scale_by_scalar', dl, dl.numel(), g.contiguous().float())
        return dl.view(ctx.shape), None


class SoftTargetCrossEntropy(nn.Module):
    """timm's SoftTargetCrossEntropy (main.py:392-394): mean over rows of sum(-target * log_softmax(x))."""

    def forward(self, x, target):
        return _SoftCEFn.apply(x, target)


# ------------------------------------------------------------------------------------------------ optimizer
class FusedAdamW:
    """torch.optim.AdamW semantics with timm's parameter grouping (no decay for biases, 1-D tensors and the names in
    model.no_weight_decay(); main.py:385, :434), executed as ONE kernel launch over all parameters, which also refreshes
    the bf16 GEMM-operand copies of the weights held by core.weights."""

    def __init__(self, model, lr=5e-4, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.step_count = 0
        skip = model.no_weight_decay() if hasattr(model, 'no_weight_decay') else set()
        self.entries = []
        for name, p in model.named_parameters():
            if not p.requires_grad:
                continue
            no_decay = p.ndim <= 1 or name.endswith('.bias') or name in skip
            self.entries.append((name, p, 0.0 if no_decay else weight_decay))
        self.state = {}
        self.shadow = {}          # name -> bf16 GEMM-operand copy refreshed by the optimizer kernel itself
        self.ema = None           # ModelEma attached with attach_ema(): parameter averages are then updated inside the same kernel
        self._table_key = None
        # timm's add_weight_decay (optim_factory.py, used by create_optimizer at main.py:385): group 0 = no-decay parameters, group 1 =
        # decayed ones, each in named_parameters() order.  `params` holds torch.optim-style indices so that state_dict() has the
        # layout of torch.optim.AdamW.state_dict() for the reference's optimizer and checkpoints interoperate.
        order = [e for e in self.entries if e[2] == 0.0] + [e for e in self.entries if e[2] != 0.0]
        self._index = {e[0]: i for i, e in enumerate(order)}
        n0 = sum(1 for e in self.entries if e[2] == 0.0)
        common = dict(lr=lr, betas=tuple(betas), eps=eps, amsgrad=False)
        self.param_groups = [dict(common, weight_decay=0.0, params=list(range(n0))),
                             dict(common, weight_decay=weight_decay, params=list(range(n0, len(order))))]
        self.chunk = _lib.lib().vsx_adamw_chunk_elems()
        self.guard = None         # (loss device scalar, int32 device counter) of the current step: see TrainStep / vsx_adamw

    # ---- checkpointing (main.py saves optimizer.state_dict() in every checkpoint and restores it on --resume)
    def state_dict(self):
        """torch.optim.AdamW layout: {'state': {index: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]}."""
        state = {}
        for name, p, _ in self.entries:
            st = self.state.get(name)
            if st is not None:
                state[self._index[name]] = {'step': torch.tensor(float(self.step_count)), 'exp_avg': st[0].clone(), 'exp_avg_sq': st[1].clone()}
        return {'state': state, 'param_groups': [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd):
        groups = sd['param_groups']
        if [len(g['params']) for g in groups] != [len(g['params']) for g in self.param_groups]:
            raise ValueError('FusedAdamW.load_state_dict: parameter groups do not match this model (%s vs %s)' %
                             ([len(g['params']) for g in groups], [len(g['params']) for g in self.param_groups]))
        for mine, theirs in zip(self.param_groups, groups):
            for k in ('lr', 'betas', 'eps', 'weight_decay'):
                if k in theirs:
                    mine[k] = tuple(theirs[k]) if k == 'betas' else theirs[k]
        self.betas, self.eps = tuple(self.param_groups[0]['betas']), self.param_groups[0]['eps']
        steps = set()
        for name, p, _ in self.entries:
            st = sd['state'].get(self._index[name], sd['state'].get(str(self._index[name])))
            if st is None:
                self.state.pop(name, None)
                continue
            if tuple(st['exp_avg'].shape) != tuple(p.shape):
                raise ValueError('FusedAdamW.load_state_dict: moment shape %s of %s does not match the parameter %s' %
                                 (tuple(st['exp_avg'].shape), name, tuple(p.shape)))
            self.state[name] = (st['exp_avg'].to(device=p.device, dtype=torch.float32).clone().contiguous(),
                                st['exp_avg_sq'].to(device=p.device, dtype=torch.float32).clone().contiguous())
            steps.add(int(float(st['step'])))
        if len(steps) > 1:
            raise ValueError('FusedAdamW.load_state_dict: per-parameter step counts differ (%s); the fused kernel keeps one' % sorted(steps))
        self.step_count = steps.pop() if steps else 0
        self._table_key = None      # moment buffers were replaced: rebuild the pointer table

    def attach_ema(self, ema):
        """Fold `ema.update(model)` for the PARAMETERS into the optimizer kernel (buffers are averaged by ModelEma.update_buffers)."""
        self.ema = ema
        self._table_key = None

    def zero_grad(self, set_to_none=True):
        for _, p, _ in self.entries:
            p.grad = None

    def _build_table(self):
        """Pointer table of the fused kernel: one 64-byte record per tensor (struct vsx_adamw_tensor), packed with numpy.  Everything
        but the gradient pointers is static between `rewiring` calls; step() refreshes the gradient column only."""
        import numpy as np
        dev = self.entries[0][1].device
        n = len(self.entries)
        rec = np.zeros(n, dtype=np.dtype([('param', '<u8'), ('grad', '<u8'), ('m', '<u8'), ('v', '<u8'), ('hi', '<u8'), ('lo', '<u8'),
                                          ('numel', '<i8'), ('wd', '<f4'), ('ema_decay', '<f4'), ('ema', '<u8')]))
        assert rec.itemsize == C.sizeof(_lib.AdamWTensor)
        chunks = []
        for i, (name, p, wd) in enumerate(self.entries):
            st = self.state.get(name)
            if st is None or st[0].shape != p.shape:
                st = (torch.zeros_like(p), torch.zeros_like(p))
                self.state[name] = st
            hi = 0
            if core.get_precision() == 'bf16' and p.ndim == 2 and p.shape[1] % 8 == 0:
                # Linear weights: the kernel also writes the bf16 operand copy the next forward's GEMMs read (saves one cast kernel per
                # weight per step); conv weights keep their re-laid-out copies in core.weights
                sh = self.shadow.get(name)
                if sh is None or sh.shape != p.shape:
                    sh = torch.empty(p.shape, device=p.device, dtype=torch.bfloat16)
                    self.shadow[name] = sh
                hi = sh.data_ptr()
            ema_ptr, ema_decay = 0, 0.0
            if self.ema is not None:
                ema_ptr, ema_decay = self.ema.param_of(name).data_ptr(), self.ema.decay
            rec[i] = (p.data_ptr(), 0, st[0].data_ptr(), st[1].data_ptr(), hi, 0, p.numel(), wd, ema_decay, ema_ptr)
            chunks.append(math.ceil(p.numel() / self.chunk))
        if getattr(self, '_chunks', None) != chunks:
            self._chunks = chunks
            ct = np.repeat(np.arange(n, dtype=np.int32), chunks)
            ci = np.concatenate([np.arange(c, dtype=np.int32) for c in chunks])
            self._ct = core.h2d(torch.from_numpy(ct), dev)
            self._ci = core.h2d(torch.from_numpy(ci), dev)
            self._nchunks = int(ct.shape[0])
        self._rec = rec
        self._grad_ptrs = None
        self._tab_has_shadow = bool((rec['hi'] != 0).any())

    def step(self, grad_scale=None):
        """grad_scale: optional fp32 device scalar multiplied into every gradient inside the kernel (1 / world_size after a SUM
        all-reduce; a loss-scaler's inverse scale)."""
        # parameters are re-allocated by `rewiring` (and moments by load_state_dict): re-derive the static part of the table when a
        # parameter pointer changed (a host-side comparison of ~250 integers)
        key = (core.get_precision(), id(self.ema)) + tuple(p.data_ptr() for _, p, _ in self.entries)
        if key != self._table_key:
            self._build_table()
            self._table_key = key
        # gradient pointers: stable from step to step when they live in the persistent gradient pool (core.grad_pool); the table is
        # uploaded again (asynchronously, through pinned memory: a blocking copy would drain the GPU once per step) only when one moved
        ptrs = []
        for _, p, _ in self.entries:
            g = p.grad
            if g is None:
                g = p.grad = torch.zeros_like(p)
            elif not g.is_contiguous():
                g = p.grad = g.contiguous()
            ptrs.append(g.data_ptr())
        if ptrs != self._grad_ptrs:
            self._rec['grad'] = ptrs
            self._tab = core.h2d(torch.from_numpy(self._rec.view('uint8').copy()), self.entries[0][1].device)
            self._grad_ptrs = ptrs
        self.step_count += 1
        lrs = set(float(g['lr']) for g in self.param_groups)
        if len(lrs) != 1:
            raise ValueError('FusedAdamW: all parameter groups must share one learning rate (got %s)' % sorted(lrs))
        self.lr = lrs.pop()
        guard, flag = self.guard if self.guard is not None else (None, None)
        self.guard = None
        ops.call('adamw', self._tab, self._ct, self._ci, self._nchunks, float(self.lr), float(self.betas[0]), float(self.betas[1]),
                 float(self.eps), self.step_count, grad_scale, guard, flag)
        core.weights.generation += 1      # parameters were written through raw pointers: operand copies are stale ...
        if self._tab_has_shadow:          # ... except the shadows this launch has just rewritten
            for name, p, _ in self.entries:
                sh = self.shadow.get(name)
                if sh is not None:
                    core.weights.adopt_shadow(p, sh)


class ModelEma:
    """timm's ModelEmaV2 (main.py:26,69,357-363; engine.py:179-180): `module` is a deep copy of the model whose state_dict entries
    follow `ema = decay * ema + (1 - decay) * model` after every optimizer step.  With `FusedAdamW.attach_ema(ema)` the parameter part
    of that update happens inside the optimizer kernel (one extra 8 B/parameter of traffic instead of a second pass over all
    weights); `update(model)` then only averages the buffers (BatchNorm running statistics), or everything when not attached."""

    def __init__(self, model, decay=0.99996):
        import copy
        self.module = copy.deepcopy(model)
        self.module.eval()
        for p in self.module.parameters():
            p.requires_grad_(False)
        self.decay = decay
        self._params = dict(self.module.named_parameters())
        self.fused = False

    def param_of(self, name):
        self.fused = True
        return self._params[name]

    @torch.no_grad()
    def update(self, model):
        ema_sd, sd = self.module.state_dict(), model.state_dict()
        for k, e in ema_sd.items():
            if self.fused and k in self._params:
                continue                      # already averaged by the optimizer kernel
            m = sd[k]
            if e.is_floating_point():
                e.mul_(self.decay).add_(m.detach().to(e.dtype), alpha=1.0 - self.decay)
            else:
                e.copy_((self.decay * e + (1.0 - self.decay) * m).to(e.dtype))      # num_batches_tracked: timm applies the same formula


# ------------------------------------------------------------------------------------------------ train step
class TrainStep:
    """One data-parallel training step, the drop-in for the body of engine.train_one_epoch."""

    def __init__(self, model, optimizer=None, criterion=None, arch_sample='multi', world_size=1, ddp_model=None, model_ema=None,
                 broadcast_buffers=True):
        self.model = model
        self.broadcast_buffers = broadcast_buffers    # native data parallelism only (a DDP wrapper broadcasts its buffers itself)
        self.net = ddp_model if ddp_model is not None else model
        self.optimizer = optimizer if optimizer is not None else FusedAdamW(model)
        self.criterion = criterion if criterion is not None else SoftTargetCrossEntropy()
        self.arch_sample = arch_sample
        self.world_size = world_size
        self.train_iter = 0           # iteration WITHIN the current epoch (the reference resets it per epoch, engine.py:100)
        self._epoch = None
        self._pool_numel = None
        self.nonfinite = None         # int32 device counter of steps whose loss was inf / nan (those steps update nothing)
        self.model_ema = model_ema
        if model_ema is not None and hasattr(self.optimizer, 'attach_ema'):
            self.optimizer.attach_ema(model_ema)
        # native data parallelism (world_size > 1 and no DDP wrapper): all parameter gradients of a step live in ONE flat pool
        # (core.grad_pool), so the exchange is an in-place NCCL all-reduce of that buffer -- no bucket copies, no per-parameter hooks.
        # The part produced by the transformer / SR blocks is reduced on a side stream while the stem backward still runs.
        self._native_dp = world_size > 1 and ddp_model is None
        self._comm_stream = None
        self._inv_world = None
        self._exchange = None

    def __call__(self, samples, targets, patch_targets, epoch=0):
        """samples [B,3,224,224], targets [B,K], patch_targets [B,16,K] on the GPU.  Returns the loss as a device scalar
        (no host sync)."""
        if epoch != self._epoch:                                 # engine.py:100: `train_iter = 0` at the top of every epoch, so the
            self._epoch, self.train_iter = epoch, 0              # sampling seed is epoch * 10000 + iteration-in-epoch
            self.sync_buffers()
        rng = None
        if self.arch_sample is not None:                         # engine.py:119-131
            rng = torch.random.get_rng_state()
            if self.arch_sample in ('single', 'hybrid'):
                torch.manual_seed(epoch * _SAMPLE_EPOCH_OFFSET + self.train_iter)
            elif self.arch_sample != 'multi':
                raise ValueError('arch_sample has invalid value {}.'.format(self.arch_sample))
        cls_pred, patch_pred = self.net(samples, patch_output_type='seq')
        loss = self.criterion(cls_pred, targets) + self.criterion(patch_pred, patch_targets)     # engine.py:153-157
        if rng is not None:
            torch.random.set_rng_state(rng)                      # engine.py:164-165
        self.train_iter += 1
        self.optimizer.zero_grad()
        if self._pool_numel is None:          # every gradient buffer padded to a multiple of 4 elements + slack for meta placeholders
            self._pool_numel = sum((p.numel() + 3) // 4 * 4 + 8 for p in self.model.parameters()) + 4096
        core.grad_pool.begin(self._pool_numel, samples.device)
        flat = core.grad_pool.flat
        if self._native_dp:
            self._arm_overlap(flat)
        try:
            loss.backward()
            used = core.grad_pool.off
        finally:
            core.grad_pool.end()
            core.trunk_grads_ready_hook = None
            core.pool_prefix_ready_hook = None
        if self.nonfinite is None:
            self.nonfinite = torch.zeros(1, dtype=torch.int32, device=samples.device)
        if hasattr(self.optimizer, 'guard'):
            self.optimizer.guard = (loss.detach(), self.nonfinite)
        if self._native_dp:
            self._finish_allreduce(flat, used)
            self.optimizer.step(grad_scale=self._inv_world)
        else:
            self.optimizer.step()
        if self.model_ema is not None:
            self.model_ema.update(self.model)                    # engine.py:179-180
        return loss.detach()

    def sync_buffers(self, src=0):
        """Rank `src`'s BatchNorm running statistics on every rank (see broadcast_buffers).  Called at the start of every epoch; call it
        before evaluating or checkpointing on a rank other than `src` in the middle of one."""
        if self._native_dp and self.broadcast_buffers:
            broadcast_buffers(self.model, src)

    def reset_epoch(self, epoch=None):
        """Start of an epoch (engine.py:100): the sampling seed counts iterations from 0 again."""
        self._epoch, self.train_iter = epoch, 0

    def check_finite(self):
        """The reference aborts on a non-finite loss after a host read-back every step (engine.py:168-173).  Here the optimizer kernel
        skips such a step on the device and counts it; call this at the logging interval (one 4-byte read-back) to abort like the
        reference does."""
        if self.nonfinite is not None:
            n = int(self.nonfinite.item())
            if n:
                raise FloatingPointError('Loss was not finite in %d step(s) since the last check, stopping training' % n)

    # ------------------------------------------------------------------ native data parallelism
    def _arm_overlap(self, flat):
        if self._comm_stream is None and flat.is_cuda:
            self._comm_stream = torch.cuda.Stream(device=flat.device)
        if self._inv_world is None:
            self._inv_world = torch.full((1,), 1.0 / self.world_size, device=flat.device)
        ex = StagedExchange(flat, self._comm_stream)
        self._exchange = ex
        core.pool_prefix_ready_hook = ex.prefix_ready       # after every stage's backward (stage 3 holds 60 % of the gradient bytes
        core.trunk_grads_ready_hook = ex.tensor_hook        # and is final a fifth of the way into the backward) and before the stem

    def _finish_allreduce(self, flat, used):
        core.pool_prefix_ready_hook = None
        exchange_pool_gradients(flat, used, self._exchange.done, self.model.parameters(), self._comm_stream)


class StagedExchange:
    """Gradient all-reduce (SUM) of the per-step gradient pool in stages (the reference's DDP overlaps bucket all-reduces with the
    backward, main.py:366-368): every gradient of a step is a view of ONE flat buffer that fills from the front in backward order
    (heads, stage 3, SR, stage 2, SR, stage 1, stem), so whenever a stage's backward has been queued the prefix flat[:off] is final
    and its not-yet-exchanged part goes to NCCL on a side stream while the earlier stages' backward keeps the SMs busy."""

    def __init__(self, flat, comm_stream=None, min_elems=1 << 20):
        self.flat, self.comm, self.done, self.min_elems = flat, comm_stream, 0, min_elems
        self.calls = 0

    def prefix_ready(self):
        import torch.distributed as dist
        off = core.grad_pool.off
        if off - self.done < self.min_elems:          # tiny ranges ride along with the next one
            return
        if self.comm is not None:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ready)
                dist.all_reduce(self.flat[self.done:off])
        else:
            dist.all_reduce(self.flat[self.done:off])
        self.done = off
        self.calls += 1

    def tensor_hook(self, grad):
        self.prefix_ready()
        return grad


def exchange_pool_gradients(flat, used, reduced_upto, params, comm_stream=None):
    """SUM all-reduce of one step's gradients: flat[reduced_upto:used] in place (flat[:reduced_upto] was already put on
    `comm_stream` by the backward hooks).  Gradients that do not live in the pool (re-laid-out conv weights, BN affine parameters:
    a few hundred KB) are first moved behind `used` -- one multi-tensor copy, p.grad re-pointed at the pool views -- so that ONE
    collective finishes the step and every gradient address is the same in every step.  The 1/world factor is applied by the
    optimizer kernel (FusedAdamW.step(grad_scale=...))."""
    import torch.distributed as dist
    lo, hi = flat.data_ptr(), flat.data_ptr() + used * flat.element_size()
    rest = [p for p in params if p.grad is not None and not (lo <= p.grad.data_ptr() < hi)]
    if rest:
        need = sum((p.grad.numel() + 3) // 4 * 4 for p in rest)
        if used + need <= flat.numel() and all(p.grad.dtype == flat.dtype for p in rest):
            views, off = [], used
            for p in rest:
                views.append(flat[off:off + p.grad.numel()].view(p.grad.shape))
                off += (p.grad.numel() + 3) // 4 * 4
            torch._foreach_copy_(views, [p.grad for p in rest])
            for p, v in zip(rest, views):
                p.grad = v
            used, rest = off, []
    if used > reduced_upto:
        dist.all_reduce(flat[reduced_upto:used])
    if rest:                                  # no room behind the pool (foreign tensors): one flattened bucket
        grads = [p.grad for p in rest]
        bucket = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(bucket)
        for g, f in zip(grads, torch._utils._unflatten_dense_tensors(bucket, grads)):
            g.copy_(f)
    if comm_stream is not None and flat.is_cuda:
        torch.cuda.current_stream(flat.device).wait_stream(comm_stream)   # the overlapped part must have landed before the optimizer
        flat.record_stream(comm_stream)


def broadcast_parameters(model, src=0):
    """Make every rank start from rank `src`'s parameters and buffers (what DistributedDataParallel does at construction)."""
    import torch.distributed as dist
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)


def broadcast_buffers(model, src=0):
    """DistributedDataParallel(broadcast_buffers=True), the reference's setting (main.py:366-368): every rank's buffers -- the BatchNorm
    running statistics of the conv stem -- are overwritten with rank `src`'s.  DDP does it before every forward; the buffers never enter
    the training arithmetic (train-mode BatchNorm normalises with batch statistics), and rank `src`'s own copy only ever sees its own
    batches either way, so doing it lazily -- at the start of every epoch and before evaluation / checkpoints (`TrainStep.sync_buffers`)
    -- leaves every observable state identical.  One collective per dtype: the buffers travel as one flat tensor."""
    import torch.distributed as dist
    by_dtype = {}
    for t in model.buffers():
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        flat = torch.cat([t.detach().reshape(-1) for t in ts])
        dist.broadcast(flat, src)
        off = 0
        with torch.no_grad():
            for t in ts:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()


class DeviceFeeder:
    """Host -> device input pipeline of the train step (the reference does `samples.to(device, non_blocking=True)` on the compute
    stream, engine.py:104-105, so every step waits for its own 170 MB upload).  Uploads run on a side stream into one of two
    device slots while the previous step computes; `next()` hands the slot to the compute stream through an event.

        feeder = DeviceFeeder(device)
        feeder.submit(samples, targets, patch_targets)        # pinned host tensors of batch i+1
        x, t, pt = feeder.next()                              # device tensors of batch i (waits only for its copy)
    """

    def __init__(self, device, normalize=None):
        """normalize=(mean, std) (per-channel sequences): uint8 image batches [B, C, H, W] submitted as the FIRST tensor are converted on
        the device, on the upload stream, to the fp32 batch torchvision's ToTensor + Normalize would have produced on the host
        (vsx_image_normalize_u8) -- the step then uploads 1 byte per pixel instead of 4."""
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.staging = [None, None]                              # uint8 device copies of the image batch (normalize mode)
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]     # recorded on the compute stream when a slot's step has been queued
        self.used = [False, False]
        self.head = self.tail = 0
        self.normalize = None
        if normalize is not None:
            mean, std = normalize
            self.normalize = ((C.c_float * len(mean))(*[float(v) for v in mean]), (C.c_float * len(std))(*[float(v) for v in std]))

    def submit(self, *host_tensors):
        i = self.head & 1
        assert self.head - self.tail < 2, 'DeviceFeeder: two batches are already in flight'
        u8 = self.normalize is not None and host_tensors[0].dtype == torch.uint8
        with torch.cuda.stream(self.stream):
            if self.used[i]:
                self.stream.wait_event(self.free[i])        # the step that read this slot has finished
            want = [(h.shape, torch.float32 if (u8 and j == 0) else h.dtype) for j, h in enumerate(host_tensors)]
            if self.slots[i] is None or [(d.shape, d.dtype) for d in self.slots[i]] != want:
                self.slots[i] = tuple(torch.empty(sh, dtype=dt, device=self.device) for sh, dt in want)
            for j, (d, h) in enumerate(zip(self.slots[i], host_tensors)):
                if u8 and j == 0:
                    st = self.staging[i]
                    if st is None or st.shape != h.shape:
                        st = self.staging[i] = torch.empty(h.shape, dtype=torch.uint8, device=self.device)
                    st.copy_(h, non_blocking=True)
                    B_, C_ = h.shape[0], h.shape[1]
                    _lib.check(_lib.lib().vsx_image_normalize_u8(st.data_ptr(), d.data_ptr(), B_, C_, h[0, 0].numel(), self.normalize[0], self.normalize[1],
                                                                 self.stream.cuda_stream))
                else:
                    d.copy_(h, non_blocking=True)
            self.events[i].record(self.stream)
        self.head += 1

    def next(self):
        assert self.tail < self.head, 'DeviceFeeder: nothing submitted'
        i = self.tail & 1
        torch.cuda.current_stream(self.device).wait_event(self.events[i])
        self.tail += 1
        self._last = i
        return self.slots[i]

    def release(self):
        """Call after the step that consumed the last `next()` has been queued on the compute stream."""
        i = self._last
        self.free[i].record(torch.cuda.current_stream(self.device))
        self.used[i] = True


def allreduce_gradients(model, world_size):
    """Gradient all-reduce (SUM / world) over NCCL when the model is not wrapped in DistributedDataParallel: one flat
    fp32 bucket per ~64 MB, reduced in place."""
    import torch.distributed as dist
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    bucket, size = [], 0
    for g in grads + [None]:
        if g is not None:
            bucket.append(g)
            size += g.numel()
        if bucket and (g is None or size >= 16 * 1024 * 1024):
            flat = torch._utils._flatten_dense_tensors(bucket)
            dist.all_reduce(flat)
            flat.div_(world_size)
            for b, f in zip(bucket, torch._utils._unflatten_dense_tensors(flat, bucket)):
                b.copy_(f)
            bucket, size = [], 0
