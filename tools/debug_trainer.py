import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import seunet_oracle as oracle
from se_unet_airseg_b200 import SE_UNet
from se_unet_airseg_b200.trainer import DataParallelTrainer
from test_gpu_training import _targets
stage = 1
sd = oracle.init_params(2, 1, seed=31)
shape = (2, 1, 16, 16, 16)
g = torch.Generator().manual_seed(5)
x = torch.rand(2, 2, 16, 16, 16, generator=g).cuda()
label, weight, skel = (t.cuda() for t in _targets(shape, 6))
ma, mb = SE_UNet(2, 1), SE_UNet(2, 1)
ma.load_state_dict(sd); mb.load_state_dict(sd)
ma, mb = ma.cuda().train(), mb.cuda().train()
opt = torch.optim.AdamW(ma.parameters(), lr=1e-4)
tr = DataParallelTrainer(mb, stage=stage)
for it in range(1):
    torch.manual_seed(100 + it)
    pe, pd = ma(x)
    loss_a = oracle.stage_loss(stage, pe, pd, label, weight, skel)
    opt.zero_grad(); loss_a.backward()
    ga = {n: (p.grad.clone() if p.grad is not None else None) for n, p in ma.named_parameters()}
    opt.step()
    torch.manual_seed(100 + it)
    loss_b = tr.step(x, label, weight, skel)
    off = 0
    for (n, a), (_, b) in zip(ma.named_parameters(), mb.named_parameters()):
        k = a.numel()
        gb = tr.grads[off:off + k].view(a.shape); off += k
        dp = (a - b).abs().max().item()
        dg = (ga[n] - gb).abs().max().item() if ga[n] is not None else -1
        if dp > 1e-6 or dg > 1e-7 * max(1e-30, gb.abs().max().item()) * 100:
            print(f"{n:22s} dparam {dp:.3e} dgrad {dg:.3e} |g| {gb.abs().max().item():.3e}")
