import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from oracle import seunet_oracle as oracle
from se_unet_airseg_b200 import SE_UNet, _lib
from se_unet_airseg_b200.inference import SlidingWindowPredictor
L = _lib.lib()
sd = oracle.init_params(2, 1, seed=777)
m = SE_UNet(2, 1); m.load_state_dict(sd); m = m.cuda().eval()
g = torch.Generator().manual_seed(11)
img = (torch.randn((40, 56, 48), generator=g) * 400 + 424).clamp_(0, 4095).round().to(torch.int16)
xref = oracle.two_channel(img.double() - 1024).float().unsqueeze(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "sw"
if mode == "sw":
    sw = SlidingWindowPredictor(m, cube=32, step=16, batch=1)
    mask, prob = sw.predict_device(img.cuda(), return_prob=True)
    x2 = sw._geom[1]["x2"]
    print("x2 == ref", torch.equal(x2.cpu(), xref), "prob nan frac", torch.isnan(prob).float().mean().item())
else:
    x2 = xref.cuda()
def check_buffers(plan, tag):
    bad = []
    for nm in ["XB","CAT1","DC5IN","D2","P1","CAT2","DC3IN","DC42IN","D1F","P2","CAT3","DC1IN","DC22IN","D0F","P3","CAT4","E7F"]:
        ptr, ch, lv = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
        _lib.check(L.seunet_plan_debug_buffer(plan.handle, nm.encode(), ctypes.byref(ptr), ctypes.byref(ch), ctypes.byref(lv)), "dbg")
        s = 32 >> lv.value
        out = torch.empty(1, ch.value * 8, s, s, s, device="cuda")
        _lib.check(L.seunet_from_chunks(ptr, ch.value, 0, 1, ch.value * 8, s, s, s, _lib.ptr(out), _lib.stream_ptr()), "from")
        torch.cuda.synchronize()
        fr = torch.isnan(out).flatten(2).float().mean(dim=2)[0].view(ch.value, 8).mean(dim=1)
        if fr.sum() > 0: bad.append((nm, [round(v, 3) for v in fr.tolist()]))
    print(tag, "NaN channels:", bad)
with torch.no_grad():
    for w in [(0,0,0)]:
        v = x2[:, :, w[0]:w[0]+32, w[1]:w[1]+32, w[2]:w[2]+32].contiguous()
        p0, p1 = m(v)
        torch.cuda.synchronize()
        print(w, "nan", torch.isnan(p1).any().item(), torch.isnan(p0).any().item())
        check_buffers(m._plan(1, 32, 32, 32, 0, v.device), str(w))
