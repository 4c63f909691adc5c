"""torchrun --nproc-per-node N tools/dp_check.py : N-rank DataParallelTrainer step (NCCL) == 1-rank step on the concatenated
batch (eval-mode DropLayer so the local-batch normaliser does not enter).  Prints max parameter difference after 2 steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from oracle import seunet_oracle as oracle
from se_unet_airseg_b200 import SE_UNet
from se_unet_airseg_b200.trainer import DataParallelTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, S = 2 * world, 32
g = torch.Generator().manual_seed(1)
x = torch.rand(B, 2, S, S, S, generator=g)
label = (torch.rand(B, 1, S, S, S, generator=g) > 0.9).float()
weight = torch.where(label > 0, torch.rand(B, 1, S, S, S, generator=g) * 2 + 0.5, torch.ones(B, 1, S, S, S))
skel = label * (torch.rand(B, 1, S, S, S, generator=g) > 0.5).float()
sd = oracle.init_params(2, 1, seed=3)

def run(trainer_world, tensors, graph=False, steps=2):
    m = SE_UNet(2, 1); m.load_state_dict(sd); m = m.to(dev).eval()
    tr = DataParallelTrainer(m, stage=3, graph=graph)
    tr.world = trainer_world
    losses, g1 = [], None
    for it in range(steps):
        losses.append(tr.step(*tensors).item())
        if it == 0:
            g1 = tr.grads.clone()      # identical parameters on both sides only at the first step
    return tr.flat.clone(), g1, losses

shard = [t.chunk(world)[rank].contiguous().to(dev) for t in (x, label, weight, skel)]
p_dp, g_dp, l_dp = run(world, shard)
# the same two-rank step replayed as a CUDA graph (both NCCL exchanges captured): steps 3 and 4 are replays
p_gr, _, l_gr = run(world, shard, graph=True, steps=4)
p_e4, _, l_e4 = run(world, shard, graph=False, steps=4)
if rank == 0:
    full = [t.to(dev) for t in (x, label, weight, skel)]
    p_1, g_1, l_1 = run(1, full)
    gerr = ((g_dp - g_1).norm() / g_1.norm()).item()
    d = (p_dp - p_1).abs()
    print(f"world={world}: losses dp {l_dp} vs single {l_1}; first-step grad rel err {gerr:.3e}; param diff mean {d.mean().item():.3e} max {d.max().item():.3e}")
    assert abs(l_dp[0] - l_1[0]) < 1e-5 and gerr < 1e-2
    dg = (p_gr - p_e4).abs()
    print(f"graph mode: losses {l_gr} vs eager {l_e4}; param diff mean {dg.mean().item():.3e} max {dg.max().item():.3e}")
    assert all(abs(a - b) < 2e-5 for a, b in zip(l_gr, l_e4)) and dg.mean().item() < 1e-5
    print("DP CHECK OK")
dist.barrier()
dist.destroy_process_group()
