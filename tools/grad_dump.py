"""One fused training step (DataParallelTrainer.step, stage 3) on fixed synthetic data; saves the flat gradient and the updated
parameters.  Used by tests/test_gpu_knobs.py to compare builds / environment knobs of the backward schedule in separate
processes (the knobs are read once per process):  python tools/grad_dump.py OUT.pt [B] [S]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import seunet_oracle as oracle
from se_unet_airseg_b200 import SE_UNet
from se_unet_airseg_b200.trainer import DataParallelTrainer

out = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = int(sys.argv[3]) if len(sys.argv) > 3 else 32
g = torch.Generator().manual_seed(11)
x = torch.rand(B, 2, S, S, S, generator=g)
label = (torch.rand(B, 1, S, S, S, generator=g) > 0.9).float()
weight = torch.where(label > 0, torch.rand(B, 1, S, S, S, generator=g) * 2 + 0.5, torch.ones_like(label))
skel = label * (torch.rand(B, 1, S, S, S, generator=g) > 0.5).float()
m = SE_UNet(2, 1)
m.load_state_dict(oracle.init_params(2, 1, seed=5))
m = m.cuda().eval()          # eval: no DropLayer draws, the step is a pure function of the inputs
tr = DataParallelTrainer(m, stage=3)
loss = tr.step(x.cuda(), label.cuda(), weight.cuda(), skel.cuda())
torch.cuda.synchronize()
torch.save({"grads": tr.grads.cpu(), "flat": tr.flat.cpu(), "loss": loss.cpu()}, out)
print("GRAD DUMP OK", float(loss))
