"""B200-native (sm_100a) SE-UNet hot path: drop-in `SE_UNet` module over the C ABI of include/seunet_b200.h."""
from .SE_UNet import SE_UNet, SSEConv, SSEConv2, CATConv, DropLayer, get_model, config  # noqa: F401

__all__ = ["SE_UNet", "SSEConv", "SSEConv2", "CATConv", "DropLayer", "get_model", "config"]
