"""Per-launch timing table of one SE_UNet forward via the plan's CUDA-event timing (dev tool). args: B S"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se_unet_airseg_b200 import SE_UNet, _lib
L = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
m = SE_UNet(2, 1).cuda().eval()
x = torch.rand(B, 2, S, S, S, device="cuda")
with torch.no_grad():
    for _ in range(2): m(x)
    plan = m._plan(B, S, S, S, 0, x.device)
    L.seunet_plan_set_timing(plan.handle, 1)
    acc = {}
    R = 3
    for _ in range(R):
        m(x); torch.cuda.synchronize()
        for i in range(L.seunet_plan_timing_count(plan.handle)):
            lab, ms, fl = ctypes.c_char_p(), ctypes.c_float(), ctypes.c_double()
            L.seunet_plan_timing_get(plan.handle, i, ctypes.byref(lab), ctypes.byref(ms), ctypes.byref(fl))
            k = lab.value.decode()
            a = acc.setdefault(k, [0.0, 0.0]); a[0] += ms.value / R; a[1] = fl.value
tot = sum(v[0] for v in acc.values())
for k, (ms, fl) in acc.items():
    extra = f"{fl/ms/1e9:8.1f} TFLOP/s" if fl else ""
    print(f"{k:12s} {ms*1e3/B:9.1f} us/patch {100*ms/tot:5.1f}% {extra}")
conv = sum(v[0] for k, v in acc.items() if k.startswith("conv:"))
print(f"total {tot*1e3/B:.1f} us/patch; conv {conv*1e3/B:.1f}; other {(tot-conv)*1e3/B:.1f}")
