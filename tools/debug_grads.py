import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import seunet_oracle as oracle
from se_unet_airseg_b200 import SE_UNet
in_ch, (B, D, H, W) = 2, (1, 16, 16, 16)
if len(sys.argv) > 1: D = H = W = int(sys.argv[1])
sd = oracle.init_params(in_ch, 1, seed=4242)
m = SE_UNet(in_ch, 1); m.load_state_dict(sd); m = m.cuda().eval()
g = torch.Generator().manual_seed(9)
x = torch.rand(B, in_ch, D, H, W, generator=g)
label = (torch.rand(B, 1, D, H, W, generator=g) > 0.9).float()
if len(sys.argv) > 2: oracle.EMULATE_STORAGE = torch.float16
sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
r0, r1 = oracle.forward(sdr, x)
oracle.stage_loss(1, r0, r1, label).backward()
p0, p1 = m(x.cuda())
oracle.stage_loss(1, p0, p1, label.cuda()).backward()
for n, p in m.named_parameters():
    r = sdr[n].grad
    if p.grad is None or r is None:
        print(f"{n:22s} ours {'None' if p.grad is None else 'set'} ref {'None' if r is None else 'set'}"); continue
    gg = p.grad.cpu().double(); r = r.double()
    rel = (gg - r).norm().item() / max(r.norm().item(), 1e-30)
    flag = "  <<<<" if rel > 1e-2 and not n.endswith("conv1.bias") else ""
    print(f"{n:22s} rel {rel:9.3e} |ref| {r.norm().item():9.3e} |ours| {gg.norm().item():9.3e}{flag}")
