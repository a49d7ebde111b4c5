"""Import the UNMODIFIED reference hot-path modules from /root/reference (test infrastructure).

Only usable where /root/reference is mounted (the build container) -- used by
``oracle/make_golden.py`` to pin the oracle and to write ``tests/golden``.  Nothing is copied:
the reference files are executed from where they lie.

Three shims are needed:
  1. ``timm`` 0.3.2 is not installed: stub the five symbols the model files import.
  2. ``nets/__init__.py`` pulls unrelated timm-heavy models: register an empty package object and load
     the six hot-path files individually.
  3. the reference hard-codes ``.cuda()``: neutralise ``Tensor.cuda`` when there is no GPU.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get('VIT_SEARCH_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF, 'nets'))


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class _PatchEmbed(nn.Module):
    """Behaviour of timm 0.3.2 PatchEmbed (only used by embed type 0)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size, patch_size = _to_2tuple(img_size), _to_2tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0])
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


_loaded = {}


def load():
    """-> dict of reference modules: drop, channel_drop, masked_layer_norm, patch_conv, supernet_blocks,
    vit_sr_supernet, supernet_config, compute_flop_mac."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not mounted at %s' % REF)

    def mk(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    mk('timm')
    mk('timm.models')
    mk('timm.models.vision_transformer', _cfg=lambda **kw: dict(kw), PatchEmbed=_PatchEmbed)
    mk('timm.models.layers', to_2tuple=_to_2tuple,
       trunc_normal_=lambda t, mean=0., std=1., a=-2., b=2.: nn.init.trunc_normal_(t, mean, std, a, b))
    mk('timm.models.registry', register_model=lambda fn: fn)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    pkg = types.ModuleType('nets')
    pkg.__path__ = [REF + '/nets']
    sys.modules['nets'] = pkg
    for n in ['drop', 'channel_drop', 'masked_layer_norm', 'patch_conv', 'supernet_blocks', 'vit_sr_supernet']:
        spec = importlib.util.spec_from_file_location('nets.' + n, '%s/nets/%s.py' % (REF, n))
        m = importlib.util.module_from_spec(spec)
        sys.modules['nets.' + n] = m
        spec.loader.exec_module(m)
        _loaded[n] = m
    for n, rel in [('supernet_config', 'supernet_config/__init__.py'),
                   ('compute_flop_mac', 'network_utils/compute_flop_mac.py')]:
        if n == 'supernet_config':
            sys.path.insert(0, REF)
            import supernet_config  # pure numpy
            sys.path.pop(0)
            _loaded[n] = supernet_config
        else:
            spec = importlib.util.spec_from_file_location('ref_' + n, '%s/%s' % (REF, rel))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            _loaded[n] = m
    return _loaded
