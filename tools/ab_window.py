"""A/B on ONE box: sliding-window volume time with the fused window step (seunet_forward_window) against the two-call path.
usage: python tools/ab_window.py [rounds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se_unet_airseg_b200 import SE_UNet
from se_unet_airseg_b200.inference import SlidingWindowPredictor
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
m = SE_UNet(2, 1).cuda().eval()
img = (torch.randn(512, 512, 400, device="cuda") * 400 + 424).clamp_(0, 4095).round().to(torch.int16)
variants = {"fused": dict(fuse_head=True), "plain": dict(fuse_head=False)}
for k, v in list(variants.items()):
    variants[k + "_1s"] = dict(streams=1, **v)
sws = {k: SlidingWindowPredictor(m, **v) for k, v in variants.items()}
for sw in sws.values():
    sw.predict_device(img, reuse_output=True)
torch.cuda.synchronize()
for r in range(rounds):
    for k, sw in sws.items():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            sw.predict_device(img, reuse_output=True)
        e1.record(); torch.cuda.synchronize()
        print(f"round {r} {k:10s} {e0.elapsed_time(e1) / 2:8.2f} ms/volume", flush=True)
