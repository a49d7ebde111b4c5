"""Prefix-mask bookkeeping.  Every mask the reference's ChannelDrop can emit keeps a PREFIX of the channels
(nets/channel_drop.py:153-157), so a [B,1,C] boolean mask is fully described by one integer per sample.  The
kernels consume those integers; the boolean tensors only exist for API compatibility with the reference's
module signatures, and carry their keep counts along as a Python attribute so no device read-back is needed."""
import torch


def make_mask(keep, width, device):
    """list[int] -> torch.bool [B,1,width] tagged with its keep counts."""
    k = torch.tensor(keep, dtype=torch.int32)
    k = (k.pin_memory().to(device, non_blocking=True) if torch.device(device).type == 'cuda' else k.to(device)).view(-1, 1, 1)
    m = torch.arange(width, device=device, dtype=torch.int32).view(1, 1, -1) < k
    m._vsx_keep = [int(v) for v in keep]
    return m


def keep_of(mask):
    """mask (None | tagged tensor | arbitrary bool tensor) -> list[int] | None.  Untagged tensors cost a device
    read-back and must be prefix masks (anything else cannot come out of the reference's training path)."""
    if mask is None:
        return None
    k = getattr(mask, '_vsx_keep', None)
    if k is not None:
        return k
    m = mask.reshape(mask.shape[0], -1)
    cnt = m.sum(dim=1)
    width = m.shape[1]
    prefix = torch.arange(width, device=m.device).view(1, -1) < cnt.view(-1, 1)
    if not bool((prefix == m).all()):
        raise ValueError('only prefix masks are supported (reference: nets/channel_drop.py:153-157)')
    return [int(v) for v in cnt.tolist()]


def and_keep(a, b):
    """Logical AND of two prefix masks given as keep lists (None = all true)."""
    if a is None:
        return b
    if b is None:
        return a
    return [min(x, y) for x, y in zip(a, b)]
