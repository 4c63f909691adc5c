"""Dev timing of SE_UNet.forward on one GPU (not the bench): python tools/time_forward.py [B] [S] [iters]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se_unet_airseg_b200 import SE_UNet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
it = int(sys.argv[3]) if len(sys.argv) > 3 else 10
torch.manual_seed(0)
m = SE_UNet(2, 1).cuda().eval()
x = torch.rand(B, 2, S, S, S, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        m(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / it
flops = 6.306e11 * B * (S / 128) ** 3
print(f"B={B} S={S}: {ms:.3f} ms/forward, {ms/B:.3f} ms/patch, {flops/ms/1e9:.1f} TFLOP/s, mem {torch.cuda.max_memory_allocated()/2**30:.2f} GiB")
if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
