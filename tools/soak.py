"""Soak test (dev tool): repeats the 128^3 forward and a training step many times and checks run-to-run consistency.
A race in the TMA/mbarrier/TMEM pipelines would show up as an occasional large deviation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se_unet_airseg_b200 import SE_UNet
from se_unet_airseg_b200.trainer import DataParallelTrainer
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
torch.manual_seed(0)
m = SE_UNet(2, 1).cuda().eval()
x = torch.rand(7, 2, 128, 128, 128, device="cuda")
with torch.no_grad():
    r0, r1 = [t.clone() for t in m(x)]
    worst = 0.0
    for i in range(iters):
        p0, p1 = m(x)
        d = max((p0 - r0).abs().max().item(), (p1 - r1).abs().max().item())
        worst = max(worst, d)
        assert d < 1e-4, f"forward run {i}: deviation {d}"
print(f"forward x{iters}: worst run-to-run deviation {worst:.3e}")
m.train()
label = (torch.rand(2, 1, 64, 64, 64, device="cuda") > 0.95).float()
xs = torch.rand(2, 2, 64, 64, 64, device="cuda")
ref = None
worst = 0.0
for i in range(iters // 4):
    torch.manual_seed(123)                      # same DropLayer draw every time
    mm = SE_UNet(2, 1)
    torch.manual_seed(5); mm.load_state_dict(m.state_dict()); mm = mm.cuda().train()
    tr = DataParallelTrainer(mm, stage=1)
    torch.manual_seed(77)
    tr.step(xs, label)
    g = tr.grads.clone()
    if ref is None:
        ref = g
    else:
        d = ((g - ref).norm() / ref.norm()).item()
        worst = max(worst, d)
        assert d < 1e-3, f"training run {i}: gradient deviation {d}"
print(f"training step x{iters // 4}: worst relative gradient deviation {worst:.3e}")
