"""CPU, world_size 2 over gloo: the data-parallel gradient exchange of engine.TrainStep."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from vit_search_b200.engine import allreduce_gradients
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5))
    for i, p in enumerate(m.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    allreduce_gradients(m, world)
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(m.parameters()))
    # the native exchange of engine.TrainStep: gradients are views of one flat pool (all-reduced in place, the first part possibly
    # already done by the backward hook) + gradients living outside the pool (one flattened bucket); SUM, scaled later by the optimizer
    from vit_search_b200.engine import exchange_pool_gradients
    ps = list(m.parameters())
    flat = torch.zeros(2048)
    off = 0
    for i, p in enumerate(ps[:3]):
        n = (p.numel() + 3) // 4 * 4
        p.grad = flat[off:off + p.numel()].view(p.shape)
        p.grad.fill_(float(rank + 1) * (i + 1))
        off += n
    ps[3].grad = torch.full_like(ps[3], float(rank + 1) * 4)          # outside the pool
    first = (ps[0].numel() + 3) // 4 * 4
    dist.all_reduce(flat[:first])                                      # what the hook does for the early part
    exchange_pool_gradients(flat, off, first, ps)
    ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 3.0 * (i + 1))) for i, p in enumerate(ps))
    ok = ok and all(p.grad.data_ptr() >= flat.data_ptr() and p.grad.data_ptr() < flat.data_ptr() + 4 * off for p in ps[:3])
    # the gradient that lived outside the pool was moved behind it (one collective, stable addresses) ...
    ok = ok and flat.data_ptr() + 4 * off <= ps[3].grad.data_ptr() < flat.data_ptr() + 4 * flat.numel()
    # ... and the staged exchange driven by the stage-backward hooks: prefixes of the pool are reduced as they become final
    from vit_search_b200 import core
    from vit_search_b200.engine import StagedExchange
    flat2 = torch.zeros(4096)
    core.grad_pool.flat, core.grad_pool.off = flat2, 0
    ex = StagedExchange(flat2, None, min_elems=256)
    a = core.grad_pool.take(1000, flat2.device)
    a.fill_(float(rank + 1))
    ex.prefix_ready()                       # 1000 elements final -> exchanged
    b = core.grad_pool.take(100, flat2.device)
    b.fill_(10.0 * (rank + 1))
    ex.prefix_ready()                       # too small: rides along with the next range
    c = core.grad_pool.take(600, flat2.device)
    c.fill_(100.0 * (rank + 1))
    ex.tensor_hook(None)                    # the trunk hook
    used = core.grad_pool.off
    ok = ok and ex.calls == 2 and ex.done == used == 1700
    exchange_pool_gradients(flat2, used, ex.done, [])
    ok = ok and bool((a == 3.0).all() and (b == 30.0).all() and (c == 300.0).all()) and bool((flat2[1700:] == 0).all())
    core.grad_pool.end()
    # DDP's broadcast_buffers (main.py:366-368): rank 0's BatchNorm running statistics on every rank, float and integer buffers alike
    from vit_search_b200.engine import broadcast_buffers
    bn = torch.nn.Sequential(torch.nn.BatchNorm2d(6), torch.nn.BatchNorm2d(3))
    for i, t in enumerate(bn.buffers()):
        t.fill_(7 * (rank + 1) + i)
    broadcast_buffers(bn)
    ok = ok and all(bool((t == 7 + i).all()) for i, t in enumerate(bn.buffers()))
    ok = ok and bn[0].num_batches_tracked.dtype == torch.int64
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _meters_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from vit_search_b200.evo_eval import EvalMeters
    m = EvalMeters('cpu')
    # what vsx_eval_metrics would have accumulated on this rank's shard: [sum of batch-mean losses, top-1, top-5, samples, batches]
    m.totals.copy_(torch.tensor([2.0 * (rank + 1), 10.0 + rank, 40.0 + rank, 100.0, 2.0], dtype=torch.float64))
    m.synchronize_between_processes()            # utils.MetricLogger.synchronize_between_processes (engine.py:233)
    r = m.result()
    ok = abs(r['loss'] - 6.0 / 4.0) < 1e-12 and abs(r['acc1'] - 100.0 * 21 / 200) < 1e-12 and abs(r['acc5'] - 100.0 * 81 / 200) < 1e-12
    q.put((rank, ok))
    dist.destroy_process_group()


def test_eval_meters_sum_over_ranks_world2_gloo():
    """Candidate evaluation shards the validation set over ranks; the meters are summed over ranks like the reference's MetricLogger."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_meters_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
