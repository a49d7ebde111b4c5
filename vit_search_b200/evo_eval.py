"""Forward-only evaluation of evolutionary-search candidates on the RESIDENT super-network weights (SURVEY.md §8(f) row 2,
BASELINE configs[4]).

The reference evaluates every candidate by building a fresh dense sub-network on the CPU, slicing the super-network checkpoint
into it (nets/net_utils.py:34-57), uploading it, wrapping it in DDP and running engine.evaluate (evo_search.py:256-273,
engine.py:194-261).  Here a candidate is a list of prefix extents applied to the one super-network that already lives in HBM
(`FlexibleDistillVisionTransformerSR.set_active_subnet`): no model build, no weight copy, no re-wrap; the cross entropy /
top-1 / top-5 meters accumulate on the device (`vsx_eval_metrics`) and are read back once per candidate.
"""
import torch

from . import core, ops


class EvalMeters:
    """Device-side accumulators of engine.evaluate's meters (engine.py:222-233, utils.MetricLogger): `loss` is the mean of per-batch
    mean losses, `acc1` / `acc5` are per-sample percentages."""

    def __init__(self, device):
        self.totals = torch.zeros(5, dtype=torch.float64, device=device)
        self._scratch = None

    def reset(self):
        self.totals.zero_()

    def update(self, logits, labels):
        core.require_cuda(logits, 'EvalMeters')
        x = logits if logits.dtype == torch.float32 and logits.stride(-1) == 1 else logits.float().contiguous()
        rows, cols = x.shape
        if self._scratch is None or self._scratch[0].numel() < rows:
            self._scratch = (torch.empty(rows, device=x.device), torch.empty(rows, dtype=torch.int32, device=x.device))
        lab = labels if labels.dtype == torch.int64 else labels.long()
        ops.call('eval_metrics', x, x.stride(0), lab, rows, cols, self._scratch[0], self._scratch[1], self.totals)

    def synchronize_between_processes(self):
        """utils.MetricLogger.synchronize_between_processes (engine.py:233): sums over ranks."""
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(self.totals)

    def result(self):
        """-> {'loss', 'acc1', 'acc5'}: ONE device->host read."""
        loss_sum, top1, top5, n, nb = self.totals.tolist()
        if n == 0:
            return dict(loss=float('nan'), acc1=float('nan'), acc5=float('nan'))
        return dict(loss=loss_sum / nb, acc1=100.0 * top1 / n, acc5=100.0 * top5 / n)


@torch.no_grad()
def evaluate(data_loader, model, device, meters=None):
    """engine.evaluate (engine.py:194-261) without the per-batch host synchronisations: eval mode, forward, hard-label cross entropy
    and top-1 / top-5 accuracy over the loader, summed over ranks.  Returns {'loss', 'acc1', 'acc5'}."""
    model.eval()
    meters = meters or EvalMeters(device)
    meters.reset()
    for images, target in data_loader:
        images = images.to(device, non_blocking=True)
        target = target.to(device, non_blocking=True)
        output = model(images)
        if isinstance(output, tuple):
            output = output[0]
        meters.update(output, target)
    meters.synchronize_between_processes()
    return meters.result()


class CandidateEvaluator:
    """The candidate loop of evo_search.py:250-285 for one resident super-network.

        ev = CandidateEvaluator(supernet, device)
        for cand in popu_evolve.popu:
            cand.score = ev.score(cand.network_def, data_loader_val)['acc1']

    `supernet` holds the checkpoint (`load_state_dict(checkpoint['model'])` once); candidates only change integer extents."""

    def __init__(self, supernet, device):
        self.model = supernet
        self.device = device
        self.meters = EvalMeters(device)
        supernet.eval()

    @torch.no_grad()
    def logits(self, sub_network_def, images):
        m = self.model
        m.eval()
        m.set_active_subnet(sub_network_def)
        try:
            out = m(images)
        finally:
            m.set_active_subnet(None)
        return out[0] if isinstance(out, tuple) else out

    def score(self, sub_network_def, data_loader):
        m = self.model
        m.set_active_subnet(sub_network_def)
        try:
            return evaluate(data_loader, m, self.device, self.meters)
        finally:
            m.set_active_subnet(None)


def sample_candidate(nd, ks, rng):
    """A random dense sub-network definition of the search space (uniform choice per entry; a removed block removes the removable
    blocks that follow it in the stage, like search_utils/gen_utils.update_depth)."""
    out, width, removing = [], None, False
    for d, k in zip(nd, ks):
        if d[0] in (0, 4, 5):
            width = int(rng.choice(list(k)))
            out.append((d[0], width) + tuple(d[2:]))
        elif d[0] == 1:
            hd = d[1][2]
            heads = int(rng.choice(list(k['attn']))) // hd
            feat = int(rng.choice(list(k['mlp'])))
            exists = 1
            if k.get('layer') is None:
                removing = False
            elif removing or int(rng.choice(list(k['layer']))) == 0:
                exists, removing = 0, True
            out.append((1, (width, heads, hd), (width, feat), exists))
        elif d[0] == 3:
            nxt = int(rng.choice(list(k)))
            out.append((3, width, nxt))
            width, removing = nxt, False
        else:
            out.append((2, width, d[2]))
    return tuple(out)
