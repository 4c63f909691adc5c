import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import seunet_oracle as oracle
from se_unet_airseg_b200 import SE_UNet, _lib
L = _lib.lib()
sd = oracle.init_params(2, 1, seed=777)
m = SE_UNet(2, 1); m.load_state_dict(sd); m = m.cuda().eval()
B, S = int(sys.argv[1]), int(sys.argv[2])
x = torch.rand(B, 2, S, S, S, device="cuda")
names = ["XB","CAT1","DC5IN","D2","P1","CAT2","DC3IN","DC42IN","D1F","P2","CAT3","DC1IN","DC22IN","D0F","P3","CAT4","E7F"]
with torch.no_grad():
    p0, p1 = m(x)
    plan = m._plan(B, S, S, S, 0, x.device)
    base = plan.ws.data_ptr()
    spans = {}
    for nm in names:
        ptr, ch, lv = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
        _lib.check(L.seunet_plan_debug_buffer(plan.handle, nm.encode(), ctypes.byref(ptr), ctypes.byref(ch), ctypes.byref(lv)), "dbg")
        s = S >> lv.value
        spans[nm] = (ptr.value - base, B * ch.value * s ** 3 * 16)
    last = max(o + n for o, n in spans.values())
    spans["REST"] = (last, plan.ws.numel() - last)
    for nm, (o, n) in spans.items():
        plan.ws[o:o + n].fill_(255)
        _lib.check(L.seunet_plan_bind(plan.handle, _lib.ptr(plan.ws), _lib.ptr(plan.wimg), _lib.stream_ptr()), "bind")
        q0, q1 = m(x)
        torch.cuda.synchronize()
        bad = torch.isnan(q1).any().item() or torch.isnan(q0).any().item()
        print(nm, "NaN" if bad else "ok", (q1 - p1).abs().max().item())
        if bad:
            plan.ws.zero_()
            _lib.check(L.seunet_plan_bind(plan.handle, _lib.ptr(plan.ws), _lib.ptr(plan.wimg), _lib.stream_ptr()), "bind")
            m(x)
