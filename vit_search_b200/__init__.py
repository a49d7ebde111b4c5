"""vit_search_b200 -- B200-native (sm_100a) implementation of the ViT-Res super-network training hot path of
yilunliao/vit-search, behind the reference's own nn.Module surface.  See DESIGN.md."""
__version__ = '0.1.0'
