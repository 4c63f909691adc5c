"""Poor man's initcheck: poison workspace + weight image with 0xFF (NaN in every dtype), re-bind, run forward."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import seunet_oracle as oracle
from se_unet_airseg_b200 import SE_UNet, _lib
L = _lib.lib()
sd = oracle.init_params(2, 1, seed=777)
m = SE_UNet(2, 1); m.load_state_dict(sd); m = m.cuda().eval()
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 32
x = torch.rand(B, 2, S, S, S, device="cuda")
with torch.no_grad():
    p0, p1 = m(x)
    plan = m._plan(B, S, S, S, 0, x.device)
    for poison_ws, poison_w in ((True, False), (False, True)):
        if poison_ws:
            plan.ws.fill_(255)
            _lib.check(L.seunet_plan_bind(plan.handle, _lib.ptr(plan.ws), _lib.ptr(plan.wimg), _lib.stream_ptr()), "bind")
        if poison_w:
            plan.wimg.fill_(255)
            plan.packed_for = None
        q0, q1 = m(x)
        torch.cuda.synchronize()
        print("poison ws" if poison_ws else "poison wimg", "-> nan:", torch.isnan(q1).any().item(), torch.isnan(q0).any().item(),
              "max diff", (q1 - p1).abs().max().item())
