"""CPU/fp32 ORACLE for the SE_UNet hot path - TEST INFRASTRUCTURE ONLY.

This module is a plain-PyTorch (fp32, functional) restatement of the reference algorithm
(Beryl2000/SE-UNet-AirSeg).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it, and only as the checker / CPU baseline - never as the
product path.  The product (se_unet_airseg_b200) must not import anything from oracle/.

Pinning: the reference ships no tests, golden vectors or weights for this path, so
the oracle is pinned against outputs of the reference module itself, imported unmodified from
/root/reference in the build container:
  * tests/test_oracle_pin.py compares it live against /root/reference/SE_UNet.py when present;
  * oracle/make_golden.py (committed) generated tests/golden/*.npz from the reference module;
    tests/test_oracle_golden.py replays them anywhere (the GPU box has no /root/reference).

Every function cites the reference file:line it follows.
"""
import math

import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------
# parameter schema (SE_UNet.py:108-153; registration order == state_dict order)
# ---------------------------------------------------------------------------------------------
# (name, kind, cin, cout, dilation, side-branch up-sample factor)
SSE = "sse"    # SSEConv  (SE_UNet.py:9-35)   one sSE gate
SSE2 = "sse2"  # SSEConv2 (SE_UNet.py:51-82)  two sSE gates
CAT = "cat"    # CATConv  (SE_UNet.py:37-49)


def layer_schema(in_channel=1):
    ic = in_channel
    return [
        ("ec1", SSE, ic, 8, 1, 1), ("ec2", SSE, 8, 16, 1, 1), ("ec3", SSE, 16, 32, 2, 1),
        ("ec33", CAT, 56, 32, 0, 0), ("x33", CAT, ic, 32, 0, 0),
        ("ec4", SSE2, 32, 32, 1, 2), ("ec5", SSE2, 32, 32, 2, 2), ("ec6", SSE2, 32, 64, 2, 2),
        ("ec63", CAT, 128, 64, 0, 0), ("x63", CAT, ic, 64, 0, 0),
        ("ec7", SSE2, 64, 64, 1, 4), ("ec8", SSE2, 64, 64, 2, 4), ("ec9", SSE2, 64, 64, 2, 4),
        ("ec93", CAT, 192, 64, 0, 0), ("x93", CAT, ic, 64, 0, 0),
        ("ec10", SSE2, 64, 64, 1, 8), ("ec11", SSE2, 64, 64, 1, 8), ("ec12", SSE2, 64, 64, 1, 8),
        ("ec123", CAT, 192, 64, 0, 0),
        ("dc1", SSE2, 128, 64, 1, 4), ("dc2", SSE2, 64, 64, 1, 4), ("dc22", CAT, 128, 64, 0, 0),
        ("dc3", SSE2, 128, 64, 1, 2), ("dc4", SSE2, 64, 32, 1, 2), ("dc42", CAT, 96, 32, 0, 0),
        ("dc5", SSE, 64, 32, 1, 1), ("dc6", SSE, 32, 16, 1, 1), ("dc62", CAT, 48, 16, 0, 0),
    ]


def param_shapes(in_channel=1, n_classes=1):
    """Ordered {name: shape} of the 117 state_dict tensors."""
    out = {}
    for name, kind, cin, cout, _dil, _up in layer_schema(in_channel):
        if kind == CAT:
            out[f"{name}.conv1.weight"] = (cout, cin, 1, 1, 1)
            continue
        out[f"{name}.conv1.weight"] = (cout, cin, 3, 3, 3)
        out[f"{name}.conv1.bias"] = (cout,)
        out[f"{name}.conv2.weight"] = (2, cout, 1, 1, 1)
        out[f"{name}.conv2.bias"] = (2,)
        out[f"{name}.conv_se.weight"] = (1, cout, 1, 1, 1)
        if kind == SSE2:
            out[f"{name}.conv_se2.weight"] = (1, cout, 1, 1, 1)
    out["dc0_0.weight"] = (n_classes, 24, 1, 1, 1)
    out["dc0_0.bias"] = (n_classes,)
    out["dc0_1.weight"] = (n_classes, 12, 1, 1, 1)
    out["dc0_1.bias"] = (n_classes,)
    return out


def init_params(in_channel=1, n_classes=1, seed=777, dtype=torch.float32):
    """Deterministic parameters with nn.Conv3d's default init law (kaiming_uniform(a=sqrt(5)) ==
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias), drawn from a private generator."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(in_channel, n_classes).items():
        if name.endswith("weight"):
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]
        else:
            wshape = param_shapes(in_channel, n_classes)[name[:-4] + "weight"]
            fan_in = wshape[1] * wshape[2] * wshape[3] * wshape[4]
        bound = 1.0 / math.sqrt(fan_in)
        sd[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return sd


# ---------------------------------------------------------------------------------------------
# blocks
# ---------------------------------------------------------------------------------------------
def _inorm(x):
    # nn.InstanceNorm3d(C): eps 1e-5, affine=False, no running stats (SE_UNet.py:17,43,59)
    return F.instance_norm(x, eps=1e-5)


def _lrelu(x):
    # nn.LeakyReLU(): negative_slope 0.01 (SE_UNet.py:18,44,60)
    return F.leaky_relu(x, 0.01)


def _up(x, factor):
    # nn.Upsample(scale_factor, mode='trilinear', align_corners=True) (SE_UNet.py:19,61,136-138)
    if factor == 1:
        return x
    return F.interpolate(x, scale_factor=factor, mode="trilinear", align_corners=True)


class _RoundSTE(torch.autograd.Function):
    """Round to a 16-bit storage type in the forward, identity in the backward (straight-through)."""

    @staticmethod
    def forward(ctx, x, dtype):
        return x.to(dtype).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g, None


# Optional emulation of the CUDA path's 16-bit STORAGE points (conv operands and raw conv outputs) inside the fp32
# oracle.  Used only by the backward parity tests: LeakyReLU's derivative is discontinuous, so the gradient of ANY
# 16-bit-storage forward differs from the fp32 gradient by several percent (a fraction ~1e-3 of the activations sits
# within rounding distance of the kink); the backward kernels are therefore checked against autograd of the oracle
# evaluated at the same (rounded) forward state.  None = plain fp32 reference semantics.
EMULATE_STORAGE = None

# "Same forward state" injection (backward parity tests only).  INJECT maps
#   "in:<layer>"    -> the conv input actually consumed by the CUDA path (its stored 16-bit activations, as fp32)
#   "w:<layer>"     -> the conv weight as rounded to the storage type
#   "raw:<layer>"   -> the CUDA path's stored raw conv output
#   "stats:<layer>" -> (mean, rstd) tensors of shape (N, C, 1, 1, 1) the CUDA path normalised with
# Each value replaces the oracle's own forward value while gradients flow through the oracle's graph unchanged
# (v_used = v_oracle + (v_injected - v_oracle).detach()).  The oracle's backward is then the exact fp32 gradient of the
# network linearised at the CUDA path's forward state, which is what the backward kernels must reproduce: against the
# plain fp32 forward, LeakyReLU's discontinuous derivative makes ANY 16-bit-storage implementation differ by several
# percent (a ~1e-3 fraction of activations sits within rounding distance of the kink).
INJECT = None


def _inj(key, t):
    if INJECT is None or key not in INJECT:
        return t
    return t + (INJECT[key].to(t.dtype) - t).detach()


def _q(t):
    return t if EMULATE_STORAGE is None else _RoundSTE.apply(t, EMULATE_STORAGE)


def _conv1(name, x, w, b, **kw):
    if INJECT is not None:
        return _inj("raw:" + name, F.conv3d(_inj("in:" + name, x), _inj("w:" + name, w), None, **kw))
    if EMULATE_STORAGE is None:
        return F.conv3d(x, w, b, **kw)
    return _q(F.conv3d(_q(x), _q(w), None, **kw))   # the bias cancels in the following InstanceNorm


def _inorm_named(name, y):
    if INJECT is None or ("stats:" + name) not in INJECT:
        return _inorm(y)
    mean = y.mean(dim=(2, 3, 4), keepdim=True)
    var = y.var(dim=(2, 3, 4), unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + 1e-5)
    m_inj, r_inj = INJECT["stats:" + name]
    mean = mean + (m_inj.to(y.dtype) - mean).detach()
    rstd = rstd + (r_inj.to(y.dtype) - rstd).detach()
    return (y - mean) * rstd


def sse_block(sd, name, x, dil, up, gates):
    """SSEConv.forward (SE_UNet.py:24-35) / SSEConv2.forward (SE_UNet.py:68-82)."""
    e0 = _conv1(name, x, sd[f"{name}.conv1.weight"], sd[f"{name}.conv1.bias"], padding=dil, dilation=dil)
    e0 = _lrelu(_inorm_named(name, e0))
    e0 = e0 * torch.sigmoid(F.conv3d(e0, sd[f"{name}.conv_se.weight"]))
    if gates == 2:
        e0 = e0 * torch.sigmoid(F.conv3d(e0, sd[f"{name}.conv_se2.weight"]))
    e1 = F.conv3d(e0, sd[f"{name}.conv2.weight"], sd[f"{name}.conv2.bias"])
    return e0, _up(e1, up)


def cat_block(sd, name, x):
    """CATConv.forward (SE_UNet.py:45-49)."""
    if x.shape[1] <= 8:       # x33/x63/x93 injection branches stay fp32 in the CUDA path
        return _lrelu(_inorm(F.conv3d(x, sd[f"{name}.conv1.weight"])))
    return _lrelu(_inorm_named(name, _conv1(name, x, sd[f"{name}.conv1.weight"], None)))


def drop_scale(batch, channel_num, threshold=0.3, generator=None):
    """DropLayer's per-(sample, channel) factor (SE_UNet.py:91-94): r = rand(B,C,1,1,1) on the CPU
    generator, binarised at `threshold`, then r*C/(r.sum()+0.01) with the sum over the WHOLE batch."""
    r = torch.rand(batch, channel_num, 1, 1, 1, generator=generator)
    r = (r >= threshold).to(torch.float32)
    return r * channel_num / (r.sum() + 0.01)


def forward(sd, x, drop0=None, drop1=None):
    """SE_UNet.forward (SE_UNet.py:181-238). drop0/drop1: DropLayer factors of shape (B,24,1,1,1) /
    (B,12,1,1,1) in training mode, None in eval mode. Returns raw logits (pred0, pred1)."""
    sch = {n: (k, dil, up) for n, k, _ci, _co, dil, up in layer_schema(x.shape[1])}

    def sse(n, t):
        k, dil, up = sch[n]
        return sse_block(sd, n, t, dil, up, 2 if k == SSE2 else 1)

    e0, s0 = sse("ec1", x)                                            # :183
    e1, s1 = sse("ec2", e0)                                           # :184
    e1_1, s2 = sse("ec3", e1)                                         # :185
    e1 = cat_block(sd, "ec33", torch.cat((e1_1, e0, e1), 1))          # :186
    e1 = e1 + cat_block(sd, "x33", x)                                 # :187
    e2 = F.max_pool3d(e1, 2, 2)                                       # :188
    x = F.max_pool3d(x, 2, 2)                                         # :189
    e2, s3 = sse("ec4", e2)                                           # :192
    e3, s4 = sse("ec5", e2)                                           # :193
    e3_1, s5 = sse("ec6", e3)                                         # :194
    e3 = cat_block(sd, "ec63", torch.cat((e3_1, e2, e3), 1))          # :195
    e3 = e3 + cat_block(sd, "x63", x)                                 # :196
    e4 = F.max_pool3d(e3, 2, 2)                                       # :197
    x = F.max_pool3d(x, 2, 2)                                         # :198
    e4, s6 = sse("ec7", e4)                                           # :201
    e5, s7 = sse("ec8", e4)                                           # :202
    e5_1, s8 = sse("ec9", e5)                                         # :203
    e5 = cat_block(sd, "ec93", torch.cat((e5_1, e4, e5), 1))          # :204
    e5 = e5 + cat_block(sd, "x93", x)                                 # :205
    e6 = F.max_pool3d(e5, 2, 2)                                       # :206
    e6, s9 = sse("ec10", e6)                                          # :209
    e7, s10 = sse("ec11", e6)                                         # :210
    e7_1, s11 = sse("ec12", e7)                                       # :211
    e7 = cat_block(sd, "ec123", torch.cat((e7_1, e6, e7), 1))         # :212
    e8 = _up(e7, 2)                                                   # :214
    d0, s12 = sse("dc1", torch.cat((e8, e5), 1))                      # :216
    d0_1, s13 = sse("dc2", d0)                                        # :217
    d0 = cat_block(sd, "dc22", torch.cat((d0_1, d0), 1))              # :218
    d1 = _up(d0, 2)                                                   # :220
    d1, s14 = sse("dc3", torch.cat((d1, e3), 1))                      # :222
    d1_1, s15 = sse("dc4", d1)                                        # :223
    d1 = cat_block(sd, "dc42", torch.cat((d1_1, d1), 1))              # :224
    d2 = _up(d1, 2)                                                   # :226
    d2, s16 = sse("dc5", torch.cat((d2, e1), 1))                      # :228
    _d2_1, s17 = sse("dc6", d2)                                       # :229
    # :230 dc62(cat(d2_1, d2)) is computed by the reference but never used - omitted.
    h0 = torch.cat((s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11), 1)
    h1 = torch.cat((s12, s13, s14, s15, s16, s17), 1)
    if drop0 is not None:
        h0 = h0 * drop0                                               # :95
    if drop1 is not None:
        h1 = h1 * drop1
    pred0 = F.conv3d(h0, sd["dc0_0.weight"], sd["dc0_0.bias"])        # :232
    pred1 = F.conv3d(h1, sd["dc0_1.weight"], sd["dc0_1.bias"])        # :233
    return pred0, pred1


# ---------------------------------------------------------------------------------------------
# losses (train.py; inputs are sigmoid probabilities, sums over the whole batch)
# ---------------------------------------------------------------------------------------------
def dice_loss(pred, target):
    """train.py:51-57."""
    smooth = 1.0
    iflat = pred.reshape(-1)
    tflat = target.reshape(-1)
    intersection = (iflat * tflat).sum()
    return 1.0 - ((2.0 * intersection + smooth) / (iflat.sum() + tflat.sum() + smooth))


def general_union_loss_lib(pred, target, weight):
    """train.py:59-68 (alpha 0.2, root exponent 0.7; sigma1 == sigma2 == 1e-4)."""
    smooth = 1.0
    alpha = 0.2
    beta = 1 - alpha
    sigma1 = 0.0001
    sigma2 = 0.0001
    weight_i = target * sigma1 + (1 - target) * sigma2
    intersection = (weight * ((pred + weight_i) ** 0.7) * target).sum()
    intersection2 = (weight * (alpha * pred + beta * target)).sum()
    return 1 - (intersection + smooth) / (intersection2 + smooth)


def atr_loss(pred, target, skel, weight):
    """train.py:70-76 (the label argument is overwritten by the skeleton mask)."""
    smooth = 1.0
    target = skel
    pred = pred * skel
    intersection = (weight * pred * target).sum()
    intersection2 = (weight * (pred + target)).sum()
    return 1 - (intersection + smooth) / (intersection2 + smooth)


def stage_loss(stage, pred_en, pred_de, label, weight=None, skel=None):
    """Loss combinations of the three curriculum stages on raw logits:
    stage 1 train.py:595-599, stage 2 train.py:429-435, stage 3 train.py:234-243."""
    pe, pd = torch.sigmoid(pred_en), torch.sigmoid(pred_de)
    if stage == 1:
        return dice_loss(pd, label) + dice_loss(pe, label)
    loss = general_union_loss_lib(pd, label, weight) + 0.5 * general_union_loss_lib(pe, label, weight)
    if stage == 3:
        loss = loss + 0.5 * (atr_loss(pe, label, skel, weight) + atr_loss(pd, label, skel, weight))
    return loss


# ---------------------------------------------------------------------------------------------
# sliding-window helpers (prediction.py)
# ---------------------------------------------------------------------------------------------
def window_starts(length, cube=128, step=64):
    """prediction.py:80-100 window enumeration along one axis: starts at k*step, the last window is
    clamped to length-cube."""
    if (length - cube) % step == 0:
        n = (length - cube) // step + 1
    else:
        n = (length - cube) // step + 2
    out = []
    for i in range(n):
        lo = i * step
        if lo + cube > length:
            lo = length - cube
        out.append(lo)
    return out


def two_channel(img_hu):
    """prediction.py:39-49 on HU values (after the -1024 shift of prediction.py:69): windows
    [-1024, 1024] and [-1000, 500], clipped and scaled to [0, 1]."""
    a = (img_hu.clamp(-1024.0, 1024.0) + 1024.0) / 2048.0
    b = (img_hu.clamp(-1000.0, 500.0) + 1000.0) / 1500.0
    return torch.stack((a, b), 0)


def predict_volume(sd, img_stored, cube=128, step=64):
    """prediction.py:65-111 up to the 0.5 threshold: img_stored holds the stored CT values (HU + 1024);
    returns (mean probability float64 volume, mask).  Eval mode like prediction.py:64.  The forward runs on the device of
    the parameter tensors (the full-size GPU parity test puts them on the B200 in fp32 with TF32 disabled); the overlap
    accumulation is the reference's host float64 `+=` either way (prediction.py:104-107)."""
    import numpy as np
    img = img_stored.cpu().to(torch.float64) - 1024                                 # :69
    x = two_channel(img).to(torch.float32).unsqueeze(0)                            # :72-77
    x = x.to(next(iter(sd.values())).device)                                        # :75-77 (.cuda())
    X, Y, Z = img.shape
    pred = np.zeros((X, Y, Z))
    pred_num = np.zeros((X, Y, Z))
    with torch.no_grad():
        for xl in window_starts(X, cube, step):                                    # :83-100
            for yl in window_starts(Y, cube, step):
                for zl in window_starts(Z, cube, step):
                    _p0, p = forward(sd, x[:, :, xl:xl + cube, yl:yl + cube, zl:zl + cube])   # :102-103
                    p = torch.sigmoid(p).cpu().numpy()[0, 0]                       # :104-105
                    pred[xl:xl + cube, yl:yl + cube, zl:zl + cube] += p            # :106
                    pred_num[xl:xl + cube, yl:yl + cube, zl:zl + cube] += 1        # :107
    pred = pred / pred_num                                                         # :109
    return pred, (pred >= 0.5)


# ---------------------------------------------------------------------------------------------------------------------
# post-processing - restated from prediction.py:13-37, 111-116 and util.py:58-75
# ---------------------------------------------------------------------------------------------------------------------
def double_threshold_iteration(pred, h_thresh, l_thresh):
    """prediction.py:13-37.  `gbin_pre = gbin` aliases the array, so the reference's while loop performs exactly ONE in-place
    raster sweep; that visiting-order dependence is part of the behaviour and is kept.  Pure-Python loops: small volumes only."""
    import numpy as np
    pred = np.array(np.asarray(pred) * 255, dtype=np.float64)
    h, w, z = pred.shape
    gbin = np.where(pred >= h_thresh * 255, 255, 0).astype(np.float64)
    neigb = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1) if (a, b, c) != (0, 0, 0)]
    for i in range(h):
        for j in range(w):
            for k in range(z):
                if gbin[i, j, k] == 0 and pred[i, j, k] < h_thresh * 255 and pred[i, j, k] >= l_thresh * 255:
                    for a, b, c in neigb:
                        if gbin[max(min(i + a, h - 1), 0), max(min(j + b, w - 1), 0), max(min(k + c, z - 1), 0)]:
                            gbin[i, j, k] = 255
                            break
    return gbin / 255


def zero_borders(pred, frac=0.15):
    """prediction.py:112-115."""
    pred = pred.copy()
    pred[0:int(frac * pred.shape[0]), :, :] = 0
    pred[int((1 - frac) * pred.shape[0]):, :, :] = 0
    pred[:, 0:int(frac * pred.shape[1]), :] = 0
    pred[:, int((1 - frac) * pred.shape[1]):, :] = 0
    return pred


def maximum_3d(region01, fill_holes=True):
    """util.py:58-75 with scipy.ndimage.label (full 3x3x3 structure) in place of cc3d (not installed here): both number the
    components in raster order of their first voxel, which fixes the tie rule of the reversed stable sort."""
    import numpy as np
    from scipy import ndimage
    label, num = ndimage.label(np.asarray(region01) != 0, structure=np.ones((3, 3, 3)))
    if num == 0:
        return np.zeros(region01.shape, dtype=bool)
    areas = np.bincount(label.ravel(), minlength=num + 1)
    num_list = list(range(1, num + 1))
    order = sorted(num_list, key=lambda x: areas[x])[::-1]
    best = label == order[0]
    z = region01.shape[2]
    if not best[:, :, z // 2].any() and not best[:, :, z // 3].any() and not best[:, :, z // 3 * 2].any() and num > 1:
        best = label == order[1]
    if fill_holes:
        best = ndimage.binary_fill_holes(best.astype(np.int8))
    return best
