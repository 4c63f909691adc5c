"""Experiment: the fused training step (forward + losses + backward + AdamW, ~250 launches) replayed as ONE CUDA graph against
the eager C-ABI calls.  Timing only: the captured AdamW step number is a constant.  python tools/graph_step.py [B] [S]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se_unet_airseg_b200 import SE_UNet
from se_unet_airseg_b200.trainer import DataParallelTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
m = SE_UNet(2, 1).cuda().eval()      # eval: the DropLayer draw is a host-side RNG call + H2D copy, which a capture cannot contain
tr = DataParallelTrainer(m, stage=2)
x = torch.rand(B, 2, S, S, S, device="cuda")
label = (torch.rand(B, 1, S, S, S, device="cuda") > 0.97).float()
weight = torch.where(label > 0, torch.rand_like(label) * 2 + 0.5, torch.ones_like(label))


def timed(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


eager = timed(lambda: tr.step(x, label, weight))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    tr.step(x, label, weight)
graph = timed(g.replay)
print(f"B={B} S={S}: eager {eager:.3f} ms/step, graph replay {graph:.3f} ms/step")
