"""One sliding-window step at the bench's window batch for ncu captures: a 128x128x512 volume is exactly 7 windows along the
last axis (stride 64), i.e. ONE seunet_forward_window call at batch 7 per predict_device.  usage: python tools/window_step.py [runs]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se_unet_airseg_b200 import SE_UNet
from se_unet_airseg_b200.inference import SlidingWindowPredictor
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
m = SE_UNet(2, 1).cuda().eval()
img = (torch.randn(128, 128, 512, device="cuda") * 400 + 424).clamp_(0, 4095).round().to(torch.int16)
sw = SlidingWindowPredictor(m, streams=1)
for _ in range(runs):
    sw.predict_device(img, reuse_output=True)
torch.cuda.synchronize()
print("done")
