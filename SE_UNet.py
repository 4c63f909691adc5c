"""`from SE_UNet import SE_UNet` shim: put this repo's root on sys.path in place of the reference checkout and
train.py / test.py / prediction.py pick up the B200-native module (same names as /root/reference/SE_UNet.py)."""
from se_unet_airseg_b200.SE_UNet import SE_UNet, SSEConv, SSEConv2, CATConv, DropLayer, get_model, config  # noqa: F401
