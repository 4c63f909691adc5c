"""Test helper: read the CUDA plan's stored forward state (conv inputs, raw conv outputs, InstanceNorm statistics) through
the seunet_plan_debug_buffer test hook and turn it into an oracle.INJECT dictionary (see oracle/seunet_oracle.py)."""
import ctypes

import torch

# layer -> (activation buffer, first chunk plane, input channels, output channels, resolution level)
LAYERS = {
    "ec1": ("XB", 0, None, 8, 0), "ec2": ("CAT1", 4, 8, 16, 0), "ec3": ("CAT1", 5, 16, 32, 0), "ec33": ("CAT1", 0, 56, 32, 0),
    "ec4": ("P1", 0, 32, 32, 1), "ec5": ("CAT2", 8, 32, 32, 1), "ec6": ("CAT2", 12, 32, 64, 1), "ec63": ("CAT2", 0, 128, 64, 1),
    "ec7": ("P2", 0, 64, 64, 2), "ec8": ("CAT3", 8, 64, 64, 2), "ec9": ("CAT3", 16, 64, 64, 2), "ec93": ("CAT3", 0, 192, 64, 2),
    "ec10": ("P3", 0, 64, 64, 3), "ec11": ("CAT4", 8, 64, 64, 3), "ec12": ("CAT4", 16, 64, 64, 3), "ec123": ("CAT4", 0, 192, 64, 3),
    "dc1": ("DC1IN", 0, 128, 64, 2), "dc2": ("DC22IN", 8, 64, 64, 2), "dc22": ("DC22IN", 0, 128, 64, 2),
    "dc3": ("DC3IN", 0, 128, 64, 1), "dc4": ("DC42IN", 4, 64, 32, 1), "dc42": ("DC42IN", 0, 96, 32, 1),
    "dc5": ("DC5IN", 0, 64, 32, 0), "dc6": ("D2", 0, 32, 16, 0),
}


def _locate(L, plan, name):
    from se_unet_airseg_b200 import _lib
    ptr, ch, lv = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
    _lib.check(L.seunet_plan_debug_buffer(plan.handle, name.encode(), ctypes.byref(ptr), ctypes.byref(ch), ctypes.byref(lv)), name)
    return ptr.value - plan.ws.data_ptr(), ch.value, lv.value


def chunk_buffer(L, plan, name, shape, dtype):
    """Chunk-plane tensor [N][chunks][D][H][W][8] living inside the plan workspace."""
    B, D, H, W = shape
    off, chunks, lv = _locate(L, plan, name)
    d, h, w = D >> lv, H >> lv, W >> lv
    n = B * chunks * d * h * w * 8
    esz = torch.tensor([], dtype=dtype).element_size()
    return plan.ws[off:off + n * esz].view(dtype).view(B, chunks, d, h, w, 8)


def channels(buf, first_chunk, C):
    """[N][chunks][D][H][W][8] -> (N, C, D, H, W) fp32 starting at chunk plane first_chunk."""
    N, k, D, H, W, _ = buf.shape
    kk = (C + 7) // 8
    return buf[:, first_chunk:first_chunk + kk].permute(0, 1, 5, 2, 3, 4).reshape(N, kk * 8, D, H, W)[:, :C].float()


def forward_state(L, plan, shape, in_ch, sd, store_dtype):
    """oracle.INJECT dictionary describing the forward state the CUDA plan computed (and will differentiate)."""
    B, D, H, W = shape
    inj = {}
    for name, (bufname, c0, cin, cout, lv) in LAYERS.items():
        cin = in_ch if cin is None else cin
        src = chunk_buffer(L, plan, bufname, shape, store_dtype)
        inj["in:" + name] = channels(src, c0, cin).cpu()
        raw = chunk_buffer(L, plan, "raw:" + name, shape, store_dtype)
        inj["raw:" + name] = channels(raw, 0, cout).cpu()
        off, COUT, _ = _locate(L, plan, "stats:" + name)
        st = plan.ws[off:off + B * COUT * 2 * 8].view(torch.float64).view(B, COUT, 2)[:, :cout].cpu()
        V = (D >> lv) * (H >> lv) * (W >> lv)
        mean = st[..., 0] / V
        var = (st[..., 1] / V - mean * mean).clamp_min(0)
        rstd = 1.0 / torch.sqrt(var + 1e-5)
        inj["stats:" + name] = (mean.float().view(B, cout, 1, 1, 1), rstd.float().view(B, cout, 1, 1, 1))
        inj["w:" + name] = sd[name + ".conv1.weight"].to(store_dtype).float()
    return inj
