"""Layer-by-layer comparison of the CUDA plan's intermediates with the oracle (dev tool)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import seunet_oracle as o
from se_unet_airseg_b200 import SE_UNet, _lib
L = _lib.lib()
sd = o.init_params(2, 1, seed=777)
m = SE_UNet(2, 1); m.load_state_dict(sd); m = m.cuda().eval()
g = torch.Generator().manual_seed(11)
img = (torch.randn((40, 56, 48), generator=g) * 400 + 424).clamp_(0, 4095).round().to(torch.int16)
x = o.two_channel(img.double() - 1024).float().unsqueeze(0)[:, :, 0:32, 0:32, 0:32].contiguous()
# oracle intermediates: raw conv1 outputs in call order
raws = []
orig = F.conv3d
def conv(inp, w, b=None, **kw):
    y = orig(inp, w, b, **kw)
    if w.shape[2] == 3 or w.shape[1] > 8 and w.shape[0] > 2:
        raws.append(y)
    return y
F.conv3d = conv
with torch.no_grad():
    r0, r1 = o.forward(sd, x)
F.conv3d = orig
xd = x.cuda()
x.requires_grad_(False)
with torch.enable_grad():
    for p in m.parameters(): p.requires_grad_(False)
with torch.no_grad():
    p0, p1 = m(xd)
# use a training-mode plan to keep per-layer raws: run with grad enabled flag off is inference; create training plan manually
plan = m._plan(1, 32, 32, 32, 1, xd.device)
plan.pack(m._flat_params(m._param_tensors()))
ones0, ones1 = torch.ones(1, 24, device="cuda"), torch.ones(1, 12, device="cuda")
q0, q1 = torch.empty_like(p0), torch.empty_like(p1)
strides = (ctypes.c_int64 * 5)(*xd.stride())
_lib.check(L.seunet_forward(plan.handle, _lib.ptr(xd), strides, None, _lib.ptr(m._flat_params(m._param_tensors())), _lib.ptr(ones0), _lib.ptr(ones1), _lib.ptr(q0), _lib.ptr(q1), _lib.stream_ptr()), "fwd")
torch.cuda.synchronize()
names = ["ec1","ec2","ec3","ec33","ec4","ec5","ec6","ec63","ec7","ec8","ec9","ec93","ec10","ec11","ec12","ec123","dc1","dc2","dc22","dc3","dc4","dc42","dc5","dc6"]
# oracle call order includes x33/x63/x93 (cin=2, filtered out by w.shape[1] > 8) -> raws align with names
print(len(raws), len(names))
for nm, ref in zip(names, raws):
    ptr, ch, lv = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
    _lib.check(L.seunet_plan_debug_buffer(plan.handle, ("raw:" + nm).encode(), ctypes.byref(ptr), ctypes.byref(ch), ctypes.byref(lv)), "dbg")
    C = ref.shape[1]; s = ref.shape[2:]
    out = torch.empty(1, C, *s, device="cuda")
    _lib.check(L.seunet_from_chunks(ptr, ch.value, 0, 1, C, s[0], s[1], s[2], _lib.ptr(out), _lib.stream_ptr()), "from")
    torch.cuda.synchronize()
    out = out.cpu()
    # the reference conv1 has a bias which cancels in IN; compare after removing per-channel mean
    a = out - out.mean(dim=(2, 3, 4), keepdim=True); b = ref - ref.mean(dim=(2, 3, 4), keepdim=True)
    print(f"{nm:6s} nan={torch.isnan(out).any().item()} inf={torch.isinf(out).any().item()} max|d|={(a-b).abs().max().item():.3e} ref max={b.abs().max().item():.3f}")
print("pred1 nan train-plan", torch.isnan(q1).any().item(), "inference-plan", torch.isnan(p1).any().item(), torch.isnan(p0).any().item())
