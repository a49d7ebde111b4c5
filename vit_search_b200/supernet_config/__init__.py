from . import sr_tiny, sr_tiny_mh, sr_tiny_666, sr_small, sr_small_mh  # noqa: F401
from ._tables import SPACES, VIT_RES_TINY, VIT_RESNAS_MEDIUM, network_def, num_channels_to_keep  # noqa: F401
