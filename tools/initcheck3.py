import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import seunet_oracle as oracle
from se_unet_airseg_b200 import SE_UNet, _lib
L = _lib.lib()
sd = oracle.init_params(2, 1, seed=777)
m = SE_UNet(2, 1); m.load_state_dict(sd); m = m.cuda().eval()
B, S, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.rand(B, 2, S, S, S, device="cuda")
order = ["XB", "raw:ec1", "CAT1", "raw:ec33", "DC5IN", "P1", "raw:ec4", "raw:ec6", "CAT2", "raw:ec63", "DC3IN", "P2", "CAT3", "DC1IN", "P3", "CAT4", "E7F",
         "DC22IN", "D0F", "DC42IN", "D1F", "D2", "T0:0", "T0:1", "T0:2", "T0:3", "T1:0", "T1:1", "T1:2"]
with torch.no_grad():
    plan = m._plan(B, S, S, S, mode, x.device)
    flat = m._flat_params(m._param_tensors())
    plan.pack(flat)
    plan.ws.fill_(255)
    _lib.check(L.seunet_plan_bind(plan.handle, _lib.ptr(plan.ws), _lib.ptr(plan.wimg), _lib.stream_ptr()), "bind")
    ones0, ones1 = torch.ones(B, 24, device="cuda"), torch.ones(B, 12, device="cuda")
    q0 = torch.empty(B, 1, S, S, S, device="cuda"); q1 = torch.empty_like(q0)
    strides = (ctypes.c_int64 * 5)(*x.stride())
    _lib.check(L.seunet_forward(plan.handle, _lib.ptr(x), strides, None, _lib.ptr(flat), _lib.ptr(ones0), _lib.ptr(ones1), _lib.ptr(q0), _lib.ptr(q1), _lib.stream_ptr()), "fwd")
    torch.cuda.synchronize()
    print("pred nan:", torch.isnan(q0).float().mean().item(), torch.isnan(q1).float().mean().item())
    for nm in order:
        ptr, ch, lv = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
        _lib.check(L.seunet_plan_debug_buffer(plan.handle, nm.encode(), ctypes.byref(ptr), ctypes.byref(ch), ctypes.byref(lv)), "dbg")
        s = S >> lv.value
        if ch.value == 0:
            t = torch.empty(B, s, s, s, device="cuda")
            ctypes.memmove  # noqa
            src = (ctypes.c_float * 1).from_address  # noqa
            # copy via torch: build tensor from pointer offset inside ws
            off = ptr.value - plan.ws.data_ptr()
            t = plan.ws[off:off + B * s ** 3 * 4].view(torch.float32).view(B, s, s, s)
            print(f"{nm:8s} nan frac {torch.isnan(t).float().mean().item():.4f}")
            continue
        out = torch.empty(B, ch.value * 8, s, s, s, device="cuda")
        _lib.check(L.seunet_from_chunks(ptr, ch.value, 0, B, ch.value * 8, s, s, s, _lib.ptr(out), _lib.stream_ptr()), "from")
        torch.cuda.synchronize()
        fr = torch.isnan(out).flatten(2).float().mean(dim=2).view(B, ch.value, 8).mean(dim=2)
        print(f"{nm:8s} nan frac per (sample, chunk): {[[round(v, 3) for v in r] for r in fr.tolist()]}")
