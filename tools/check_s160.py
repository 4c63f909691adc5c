"""Config 5 sanity (S=160, stage 3): forward parity vs the fp32 oracle at 160^3 and a timed training step. Dev tool."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import seunet_oracle as oracle
from se_unet_airseg_b200 import SE_UNet
sd = oracle.init_params(2, 1, seed=777)
m = SE_UNet(2, 1); m.load_state_dict(sd); m = m.cuda().eval()
g = torch.Generator().manual_seed(160)
x = torch.rand(1, 2, 160, 160, 160, generator=g)
with torch.no_grad():
    r0, r1 = oracle.forward(sd, x)
    p0, p1 = m(x.cuda())
for n, p, r in (("pred0", p0, r0), ("pred1", p1, r1)):
    e = (p.cpu() - r).abs().max().item()
    a = ((p.cpu() >= 0) == (r >= 0)).float().mean().item()
    print(f"160^3 {n}: max|dlogit|={e:.3e} mask agreement={a:.5f}")
    assert e <= 2e-2
