"""Compare dn / dY of dc6 (first block of the backward) with autograd of the oracle."""
import sys, os, ctypes
os.environ["SEUNET_DEBUG_BWD_STOP_AFTER_DC6"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, torch.nn.functional as F
from oracle import seunet_oracle as oracle
from se_unet_airseg_b200 import SE_UNet, _lib
from test_gpu_conv import from_chunk_planes
L = _lib.lib()
S = 16
sd = oracle.init_params(2, 1, seed=4242)
m = SE_UNet(2, 1); m.load_state_dict(sd); m = m.cuda().eval()
g = torch.Generator().manual_seed(9)
x = torch.rand(1, 2, S, S, S, generator=g)
label = (torch.rand(1, 1, S, S, S, generator=g) > 0.9).float()
sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
caps = []
orig = F.conv3d
def conv(inp, w, b=None, **kw):
    y = orig(inp, w, b, **kw)
    if w.shape[2] == 3:
        y.retain_grad(); caps.append((inp, y))
    return y
F.conv3d = conv
r0, r1 = oracle.forward(sdr, x)
F.conv3d = orig
oracle.stage_loss(1, r0, r1, label).backward()
xin_ref, y_ref = caps[-1]          # dc6.conv1 is the last 3x3x3 conv of the forward
dy_ref = y_ref.grad
p0, p1 = m(x.cuda())
oracle.stage_loss(1, p0, p1, label.cuda()).backward()
torch.cuda.synchronize()
plan = m._plan(1, S, S, S, 1, torch.device("cuda", 0))
def buf(name, dtype, C):
    ptr, ch, lv = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
    _lib.check(L.seunet_plan_debug_buffer(plan.handle, name.encode(), ctypes.byref(ptr), ctypes.byref(ch), ctypes.byref(lv)), "dbg")
    off = ptr.value - plan.ws.data_ptr()
    esz = torch.tensor([], dtype=dtype).element_size()
    n = ch.value * S ** 3 * 8
    t = plan.ws[off:off + n * esz].view(dtype).view(1, ch.value, S, S, S, 8)
    return from_chunk_planes(t, C).cpu()
ptr, ch, lv = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
_lib.check(L.seunet_plan_debug_buffer(plan.handle, b"scale:17", ctypes.byref(ptr), ctypes.byref(ch), ctypes.byref(lv)), "dbg")
off = ptr.value - plan.ws.data_ptr()
scale = plan.ws[off:off + 8].view(torch.float32).cpu()
print("scale", scale.tolist())
dy = buf("dy:0", torch.float16, 16) / scale[0]
dn = buf("dn:0", torch.float32, 16)
x_in = buf("D2", torch.float16, 32)
print("x input  rel err", ((x_in - xin_ref.detach()).norm() / xin_ref.detach().norm()).item())
print("dy       rel err", ((dy - dy_ref).norm() / dy_ref.norm()).item(), "|ref|", dy_ref.norm().item())
# reference dn: dL/d(norm output) - recompute from dy? use autograd of instance norm alone
yv = y_ref.detach().clone().requires_grad_(True)
nv = F.instance_norm(yv, eps=1e-5)
# dn_ref such that IN-backward(dn_ref) = dy_ref is not unique; instead push our dn through torch's IN backward
(dy_from_dn,) = torch.autograd.grad(nv, yv, dn)
print("IN_bwd(our dn) vs ref dy", ((dy_from_dn - dy_ref).norm() / dy_ref.norm()).item())
print("IN_bwd(our dn) vs our dy", ((dy_from_dn - dy).norm() / dy.norm()).item())
w = torch.zeros(16, 32, 3, 3, 3, requires_grad=True)
(dw_from_ours,) = torch.autograd.grad(F.conv3d(x_in, w, padding=1), w, dy)
(dw_ref,) = torch.autograd.grad(F.conv3d(xin_ref.detach(), w, padding=1), w, dy_ref)
print("torch wgrad(our x, our dy) vs ref dW", ((dw_from_ours - dw_ref).norm() / dw_ref.norm()).item())
ours = m.dc6.conv1.weight.grad.cpu()
print("our dW vs torch wgrad(our x, our dy)", ((ours - dw_from_ours).norm() / dw_from_ours.norm()).item())
print("our dW vs ref dW", ((ours - sdr['dc6.conv1.weight'].grad).norm() / dw_ref.norm()).item())
