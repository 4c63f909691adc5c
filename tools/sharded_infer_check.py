"""torchrun --nproc-per-node N tools/sharded_infer_check.py [X Y Z cube step]
Patch-sharded sliding-window inference of ONE volume over N ranks (NCCL reduce of the fixed-point partial volumes,
prediction.py:80-110 sharded by window) must give the bit-identical mask and mean probability as the 1-rank run.
Prints "SHARDED INFER OK" on success (tests/test_gpu_multi_device.py runs it with N = 2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from oracle import seunet_oracle as oracle
from se_unet_airseg_b200 import SE_UNet
from se_unet_airseg_b200.inference import SlidingWindowPredictor

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
args = [int(a) for a in sys.argv[1:]]
X, Y, Z, cube, step = (args + [112, 64, 80, 32, 16][len(args):])[:5]

sd = oracle.init_params(2, 1, seed=777)
m = SE_UNet(2, 1)
m.load_state_dict(sd)
m = m.to(dev).eval()
rng = np.random.RandomState(5)
img = np.clip(rng.randn(X, Y, Z) * 400.0 + 424.0, 0, 4095).round().astype(np.int16)
sw = SlidingWindowPredictor(m, cube=cube, step=step, batch=7, streams=3)

res = sw.predict_device_sharded(torch.from_numpy(img).to(dev), return_prob=True)
host = sw.predict_sharded(img)
if rank == 0:
    mask_s, prob_s = res
    mask_1, prob_1 = sw.predict_device(torch.from_numpy(img).to(dev), return_prob=True)
    same_mask = torch.equal(mask_s, mask_1)
    same_prob = torch.equal(prob_s, prob_1)
    same_host = torch.equal(host, mask_s.cpu())
    dmax = (prob_s - prob_1).abs().max().item()
    print(f"world={world} volume={X}x{Y}x{Z} windows={len(sw._geometry((X, Y, Z), dev)['wins'])}: mask identical {same_mask}, "
          f"mean probability identical {same_prob} (max diff {dmax:.2e}), host path identical {same_host}, "
          f"foreground {mask_1.float().mean().item():.4f}")
    # The exchange itself is exact (integer sums).  The per-window forward is invariant to how windows are batched up to
    # the order of the fp64 statistics atomics, so the two runs agree to ~1e-7 and bit for bit in practice.
    assert same_host and dmax <= 1e-6 and (mask_s != mask_1).sum().item() <= 2
    print("SHARDED INFER OK")
else:
    assert res is None and host is None
dist.barrier()
dist.destroy_process_group()
