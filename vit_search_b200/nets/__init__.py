from .channel_drop import ChannelDrop  # noqa: F401
from .drop import DropPath, drop_path  # noqa: F401
from .masked_layer_norm import MaskedLayerNorm, MaskedLayerNormFunc  # noqa: F401
from .supernet_blocks import Attention, Block, Mlp  # noqa: F401
