from ._tables import num_channels_to_keep as _k, network_def as _nd

num_channels_to_keep = _k('sr_tiny_666')
network_def = _nd('sr_tiny_666')
