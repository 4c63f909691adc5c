"""Dev timing of the fused training step (DataParallelTrainer.step) on one GPU: python tools/time_train.py [B] [S] [stage]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se_unet_airseg_b200 import SE_UNet, _lib
from se_unet_airseg_b200.trainer import DataParallelTrainer
L = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
stage = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
m = SE_UNet(2, 1).cuda().train()
tr = DataParallelTrainer(m, stage=stage)
x = torch.rand(B, 2, S, S, S, device="cuda")
label = (torch.rand(B, 1, S, S, S, device="cuda") > 0.97).float()
weight = torch.where(label > 0, torch.rand_like(label) * 2 + 0.5, torch.ones_like(label))
skel = label * (torch.rand_like(label) > 0.5).float()
for _ in range(2):
    loss = tr.step(x, label, weight, skel)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
it = 3
e0.record()
for _ in range(it):
    loss = tr.step(x, label, weight, skel)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / it
print(f"B={B} S={S} stage={stage}: {ms:.1f} ms/step, {B/ms*1e3:.1f} patches/s, loss {loss.item():.4f}, mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB, "
      f"{1.89e12*B*(S/128)**3/ms/1e9:.0f} TFLOP/s")
plan = m._plan(B, S, S, S, 1, x.device)
L.seunet_plan_set_timing(plan.handle, 1)
tr.step(x, label, weight, skel); torch.cuda.synchronize()
acc = {}
for i in range(L.seunet_plan_timing_count(plan.handle)):
    lab, t, fl = ctypes.c_char_p(), ctypes.c_float(), ctypes.c_double()
    L.seunet_plan_timing_get(plan.handle, i, ctypes.byref(lab), ctypes.byref(t), ctypes.byref(fl))
    k = lab.value.decode().split(":")[0]
    a = acc.setdefault(k, [0.0, 0.0]); a[0] += t.value; a[1] += fl.value
print("backward breakdown (ms/step):", {k: round(v[0], 2) for k, v in acc.items()})
print("wgrad TFLOP/s", acc["wgrad"][1] / acc["wgrad"][0] / 1e9, "dgrad TFLOP/s", acc["dgrad"][1] / acc["dgrad"][0] / 1e9)
if os.environ.get("DETAIL"):
    for i in range(L.seunet_plan_timing_count(plan.handle)):
        lab, t, fl = ctypes.c_char_p(), ctypes.c_float(), ctypes.c_double()
        L.seunet_plan_timing_get(plan.handle, i, ctypes.byref(lab), ctypes.byref(t), ctypes.byref(fl))
        print(f"  {lab.value.decode():16s} {t.value*1e3/B:9.1f} us/patch")
