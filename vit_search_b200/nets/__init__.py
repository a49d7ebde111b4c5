from .channel_drop import ChannelDrop  # noqa: F401
from .drop import DropPath, drop_path  # noqa: F401
from .masked_layer_norm import MaskedLayerNorm, MaskedLayerNormFunc  # noqa: F401
from .patch_conv import PatchConvEmbed, PatchEmbed  # noqa: F401
from .registry import create_model, list_models, register_model  # noqa: F401
from .supernet_blocks import Attention, Block, Mlp  # noqa: F401
from .vit_sr_supernet import (BypassBlock, FlexibleDistillVisionTransformerSR,  # noqa: F401
                              SpatialReductionPatchEmbedding)
from . import vit_sr_supernet  # noqa: F401
from .vision_transformer_supernet import FlexibleDistillVisionTransformer  # noqa: F401
from . import vision_transformer_supernet  # noqa: F401
