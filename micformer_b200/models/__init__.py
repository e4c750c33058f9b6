from .MICFormer_self import (Head, MicFormer, BasicLayer, CrossTransformerBlock3D, TransformerBlock3D,  # noqa: F401
                             CrossWindowAttention3D, WindowAttention3D, PatchEmbed3D, PatchMerging, PatchExpand, Mlp,
                             LayerNormProxy, DropPath, window_partition, window_reverse, get_window_size)
from .STN import SpatialTransformer, Re_SpatialTransformer  # noqa: F401
