"""crop2seg_b200 -- B200-native (sm_100a) L-TAE + TemporalAggregator hot path of Many98/Crop2Seg.

Public surface (mirrors the reference's interface for this path):

    LTAE, LTAE4WTAE, TemporalAggregator     drop-in nn.Modules (modules.py)
    ops.ltae_forward, ops.temporal_aggregate tensor-level calls over the C ABI (ops.py)
    install()                               swap the three classes into the reference's model files
    shard_patches()                         patch sharding for multi-GPU inference (sharding.py)
    copy_valid_frames_()                    host -> device staging that skips padded frames (staging.py)
    CrossEntropyLoss, FocalCELoss, boundary_target   loss side of the training step (losses.py)
    TilePatchifier, ClassMap                tile pipeline edges on the device: raw tile -> model inputs, logits -> class map (tile.py)

The arithmetic lives in ``lib/libcrop2seg_b200.so`` (``include/crop2seg_b200.h``), built in-tree by
``python -m crop2seg_b200.build``.  There is no CPU or PyTorch fallback.
"""
from .modules import LTAE, LTAE4WTAE, TemporalAggregator  # noqa: F401
from .install import install, uninstall  # noqa: F401
from .sharding import GradientBucket, gather_shards, shard_bounds, shard_patches  # noqa: F401
from .staging import copy_valid_frames_, frame_slots, gather_frames, scatter_frames, smart_forward, valid_lengths  # noqa: F401
from . import ops  # noqa: F401
from .tile import ClassMap, TilePatchifier, patch_grid  # noqa: F401
from .ops import pad_mask_from_input  # noqa: F401
from . import losses  # noqa: F401
from . import conv  # noqa: F401
from .conv import ConvBlock, ConvLayer, DownConvBlock  # noqa: F401
from .losses import CrossEntropyLoss, FocalCELoss, boundary_target  # noqa: F401

__all__ = ["LTAE", "LTAE4WTAE", "TemporalAggregator", "install", "uninstall", "shard_patches", "shard_bounds", "gather_shards", "copy_valid_frames_", "valid_lengths", "smart_forward", "pad_mask_from_input", "ops", "TilePatchifier", "ClassMap", "patch_grid", "frame_slots", "gather_frames", "scatter_frames", "GradientBucket", "losses", "CrossEntropyLoss", "FocalCELoss", "boundary_target", "conv", "ConvLayer", "ConvBlock", "DownConvBlock"]
