"""Pure-write and pure-read HBM bandwidth probes (denominators for write-dominated kernels such as the trunk up-sampling)."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(f, reps=10):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for _ in range(3): a.zero_(); b.copy_(a)
print(f"memset 1 GiB: {n / t(lambda: a.zero_()) / 1e6:.0f} GB/s (write only)")
print(f"copy 1 GiB:   {2 * n / t(lambda: b.copy_(a)) / 1e6:.0f} GB/s (read + write)")
af = a.view(torch.float32)
print(f"sum 1 GiB:    {n / t(lambda: af.sum()) / 1e6:.0f} GB/s (read only)")
