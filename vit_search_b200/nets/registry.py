"""Minimal model registry with timm's calling convention (timm is not a dependency of the hot path).

`create_model(name, **kwargs)` drops None-valued kwargs like timm 0.3.2 does (main.py:329-348 passes
`drop_block_rate=None`), so the reference's main.py can build these models unchanged; when timm IS importable the
factories are additionally registered with timm's own registry.
"""
_MODELS = {}


def register_model(name_or_fn, fn=None):
    if fn is None:                      # decorator form
        fn = name_or_fn
        name = fn.__name__
    else:
        name = name_or_fn
        fn.__name__ = name
    _MODELS[name] = fn
    try:                                 # pragma: no cover - timm absent in this image
        from timm.models.registry import register_model as timm_register
        fn.__module__ = __name__
        timm_register(fn)
    except Exception:
        pass
    return fn


def create_model(model_name, pretrained=False, **kwargs):
    if model_name not in _MODELS:
        raise RuntimeError('Unknown model (%s); known: %s' % (model_name, sorted(_MODELS)))
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return _MODELS[model_name](pretrained=pretrained, **kwargs)


def list_models():
    return sorted(_MODELS)
