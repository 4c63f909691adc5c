"""Drop-in replacement for the reference `SE_UNet.py` (Beryl2000/SE-UNet-AirSeg).

Same public surface as the reference module - `SE_UNet(in_channel=1, n_classes=1)`, the 117-tensor
`state_dict()` layout, `forward(x) -> (pred0, pred1)` raw fp32 logits, `SSEConv` / `SSEConv2` /
`CATConv` / `DropLayer` / `get_model` / `config` - but `SE_UNet.forward` and its backward run as
hand-written sm_100a kernels behind the C ABI of include/seunet_b200.h (libseunet_b200.so).

There is NO PyTorch/cuDNN/CPU fallback for the network: without the CUDA library or on a non-CUDA
input `forward` raises.  PyTorch is used for device memory, streams and autograd plumbing only.

Reference citations are to /root/reference/SE_UNet.py.
"""
import ctypes
import threading

import torch
import torch.nn as nn

from . import _lib

config = {}  # SE_UNet.py:7


# -------------------------------------------------------------------------------------------------
# Building blocks: parameter containers with the reference's attribute names, so that state_dict
# keys, default initialisation (same RNG draw order) and load_state_dict(strict=False) behave
# exactly like the reference classes.  The compute of the full network is NOT a composition of
# these modules' forwards - SE_UNet.forward hands the whole graph to the CUDA plan.
# -------------------------------------------------------------------------------------------------
class SSEConv(nn.Module):
    """Parameters of SE_UNet.py:9-35 (conv3x3x3 -> InstanceNorm -> LeakyReLU -> one sSE gate -> 1x1x1 side conv)."""
    num_gates = 1

    def __init__(self, in_channel=1, out_channel1=1, out_channel2=2, stride=1, kernel_size=3,
                 padding=1, dilation=1, down_sample=1, bias=True):
        self.in_channel = in_channel
        self.out_channel = out_channel1
        super().__init__()
        self.conv1 = nn.Conv3d(in_channel, out_channel1, kernel_size, stride=stride, padding=padding * dilation,
                               bias=bias, dilation=dilation)
        self.conv2 = nn.Conv3d(out_channel1, out_channel2, kernel_size=1, stride=1, padding=0, bias=bias)
        self.norm = nn.InstanceNorm3d(out_channel1)
        self.act = nn.LeakyReLU(inplace=True)
        self.up_sample = nn.Upsample(scale_factor=down_sample, mode='trilinear', align_corners=True)
        self.conv_se = nn.Conv3d(out_channel1, 1, kernel_size=1, stride=1, padding=0, bias=False)
        self.norm_se = nn.Sigmoid()

    def forward(self, x):
        raise NotImplementedError(
            "SSEConv/SSEConv2/CATConv are parameter containers in the B200 build; run them through SE_UNet.forward")


class SSEConv2(SSEConv):
    """Parameters of SE_UNet.py:51-82 (as SSEConv with a second sSE gate)."""
    num_gates = 2

    def __init__(self, in_channel=1, out_channel1=1, out_channel2=2, stride=1, kernel_size=3,
                 padding=1, dilation=1, down_sample=1, bias=True):
        super().__init__(in_channel, out_channel1, out_channel2, stride, kernel_size, padding, dilation,
                         down_sample, bias)
        self.conv_se2 = nn.Conv3d(out_channel1, 1, kernel_size=1, stride=1, padding=0, bias=False)
        self.norm_se2 = nn.Sigmoid()


class CATConv(nn.Module):
    """Parameters of SE_UNet.py:37-49 (1x1x1 conv, no bias -> InstanceNorm -> LeakyReLU)."""

    def __init__(self, in_channel=1, out_channel1=1):
        self.in_channel = in_channel
        self.out_channel = out_channel1
        super().__init__()
        self.conv1 = nn.Conv3d(in_channel, out_channel1, kernel_size=1, stride=1, padding=0, bias=False)
        self.norm = nn.InstanceNorm3d(out_channel1)
        self.act = nn.LeakyReLU(inplace=True)

    def forward(self, x):
        raise NotImplementedError(
            "SSEConv/SSEConv2/CATConv are parameter containers in the B200 build; run them through SE_UNet.forward")


class DropLayer(nn.Module):
    """SE_UNet.py:84-97.  `scale(batch, device)` returns the per-(sample, channel) factor the reference
    multiplies into the stacked side branches: r = rand(B,C,1,1,1) drawn on the CPU default generator
    (reference line 91), binarised at `threshold`, times C/(r.sum()+0.01) with the sum over the whole
    local batch.  The multiplication itself is folded into the head weights on the device."""

    def __init__(self, channel_num=1, thr=0.3):
        super().__init__()
        self.channel_num = channel_num
        self.threshold = thr

    def scale(self, batch, device):
        if self.training:
            r = torch.rand(batch, self.channel_num, 1, 1, 1)
            r = (r >= self.threshold).to(torch.float32)
            r = r * self.channel_num / (r.sum() + 0.01)
            return r.reshape(batch, self.channel_num).to(device, non_blocking=True)
        return torch.ones(batch, self.channel_num, device=device, dtype=torch.float32)

    def forward(self, x):
        if self.training:
            return x * self.scale(x.shape[0], x.device).reshape(x.shape[0], self.channel_num, 1, 1, 1)
        return x


# -------------------------------------------------------------------------------------------------
# CUDA plan cache (per module replica / device / shape)
# -------------------------------------------------------------------------------------------------
class _Plan:
    """One bound C-ABI plan: launch descriptors + caller-owned workspace and packed weight image."""

    def __init__(self, batch, D, H, W, in_ch, n_classes, mode, device):
        L = _lib.lib()
        self.key = (batch, D, H, W, in_ch, n_classes, mode, device.index)
        self.device = device
        h = ctypes.c_void_p()
        _lib.check(L.seunet_plan_create(ctypes.byref(h), batch, D, H, W, in_ch, n_classes, mode, device.index),
                   "seunet_plan_create")
        self.handle = h
        self.ws = torch.empty(L.seunet_plan_workspace_bytes(h), dtype=torch.uint8, device=device)
        self.wimg = torch.empty(L.seunet_plan_wimg_bytes(h), dtype=torch.uint8, device=device)
        _lib.check(L.seunet_plan_bind(h, _lib.ptr(self.ws), _lib.ptr(self.wimg), _lib.stream_ptr()), "seunet_plan_bind")
        self.nbytes = self.ws.numel() + self.wimg.numel()
        self.packed_gen = None  # weight generation (see SE_UNet._weights) the tensor-core image was packed from
        self.generation = 0     # bumped by every autograd forward (guards backward against workspace reuse)

    def pack(self, flat, gen):
        """Re-pack the tensor-core weight image when the parameters changed.  `gen` is the owner's monotonically
        increasing weight generation - NOT an address/version tag: a re-gathered flat buffer is a fresh tensor whose
        (data_ptr, _version) can repeat an older one when the caching allocator recycles the block."""
        if self.packed_gen != gen:
            _lib.check(_lib.lib().seunet_pack_weights(self.handle, _lib.ptr(flat), _lib.stream_ptr()), "seunet_pack_weights")
            self.packed_gen = gen

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().seunet_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_MAX_PLANS = 12              # bound plans kept per (module replica, device): LRU beyond this count ...
_MAX_PLAN_BYTES = 96 << 30   # ... or beyond this many bytes of workspaces (a 7-window inference plan holds ~8 GB)


class _SEUNetFunction(torch.autograd.Function):
    """autograd boundary: forward = seunet_forward, backward = seunet_backward (C ABI).  The plan's workspace holds the
    activations between the two calls, so a second forward on the same plan invalidates a pending backward."""

    @staticmethod
    def forward(ctx, module, plan, x, flat, drop0, drop1, *params):
        L = _lib.lib()
        B, _, D, H, W = x.shape
        pred0 = torch.empty((B, module.n_classes, D, H, W), dtype=torch.float32, device=x.device)
        pred1 = torch.empty_like(pred0)
        strides = (ctypes.c_int64 * 5)(*x.stride())
        _lib.check(L.seunet_forward(plan.handle, _lib.ptr(x), strides, None, _lib.ptr(flat), _lib.ptr(drop0),
                                    _lib.ptr(drop1), _lib.ptr(pred0), _lib.ptr(pred1), _lib.stream_ptr()), "seunet_forward")
        plan.generation += 1
        ctx.module, ctx.plan, ctx.generation = module, plan, plan.generation
        ctx.save_for_backward(x, flat, drop0, drop1)
        ctx.shapes = [p.shape for p in params]
        ctx.needs = [p.requires_grad for p in params]
        return pred0, pred1

    @staticmethod
    def backward(ctx, g0, g1):
        L = _lib.lib()
        x, flat, drop0, drop1 = ctx.saved_tensors
        plan = ctx.plan
        if plan.generation != ctx.generation:
            raise _lib.SeunetError("backward() after a newer forward pass reused this plan's activation workspace; "
                                   "call backward before the next forward of the same shape")
        shape = (x.shape[0], ctx.module.n_classes) + tuple(x.shape[2:])
        g0 = torch.zeros(shape, dtype=torch.float32, device=x.device) if g0 is None else g0.contiguous().float()
        g1 = torch.zeros(shape, dtype=torch.float32, device=x.device) if g1 is None else g1.contiguous().float()
        gflat = torch.empty_like(flat)
        strides = (ctypes.c_int64 * 5)(*x.stride())
        with torch.cuda.device(x.device):
            _lib.check(L.seunet_backward(plan.handle, _lib.ptr(x), strides, None, _lib.ptr(flat), _lib.ptr(drop0),
                                         _lib.ptr(drop1), _lib.ptr(g0), _lib.ptr(g1), _lib.ptr(gflat), _lib.stream_ptr()),
                       "seunet_backward")
        grads, off = [], 0
        for shp, need in zip(ctx.shapes, ctx.needs):
            n = shp.numel()
            grads.append(gflat[off:off + n].view(shp) if need else None)
            off += n
        # dc62 is dead code in the reference graph (SE_UNet.py:230): its .grad stays None there too
        grads[ctx.module._dead_param_index] = None
        return (None, None, None, None, None, None, *grads)


def _aliases_flat(params, flat):
    """True when every tensor of `params` is the view of `flat` at its state_dict offset (DataParallelTrainer layout)."""
    if flat.device != params[0].device:
        return False
    base, off = flat.data_ptr(), 0
    for p in params:
        if p.dtype != torch.float32 or not p.is_contiguous() or p.data_ptr() != base + 4 * off:
            return False
        off += p.numel()
    return off == flat.numel()


class SE_UNet(nn.Module):
    """SE_UNet.py:99-238 with the forward/backward replaced by the sm_100a plan."""

    def __init__(self, in_channel=1, n_classes=1):
        self.in_channel = in_channel
        self.n_classes = n_classes
        self.batchnorm = False
        self.bias = True
        self.out_channel2 = 2
        self.sigmoid_output = 0
        super().__init__()
        # creation order == SE_UNet.py:108-153 (state_dict order and RNG draw order)
        self.ec1 = SSEConv(self.in_channel, 8, self.out_channel2, bias=self.bias)
        self.ec2 = SSEConv(8, 16, self.out_channel2, bias=self.bias)
        self.ec3 = SSEConv(16, 32, self.out_channel2, bias=self.bias, dilation=2)
        self.ec33 = CATConv(56, 32)
        self.x33 = CATConv(self.in_channel, 32)

        self.ec4 = SSEConv2(32, 32, self.out_channel2, bias=self.bias, down_sample=2)
        self.ec5 = SSEConv2(32, 32, self.out_channel2, bias=self.bias, dilation=2, down_sample=2)
        self.ec6 = SSEConv2(32, 64, self.out_channel2, bias=self.bias, dilation=2, down_sample=2)
        self.ec63 = CATConv(128, 64)
        self.x63 = CATConv(self.in_channel, 64)

        self.ec7 = SSEConv2(64, 64, self.out_channel2, bias=self.bias, down_sample=4)
        self.ec8 = SSEConv2(64, 64, self.out_channel2, bias=self.bias, dilation=2, down_sample=4)
        self.ec9 = SSEConv2(64, 64, self.out_channel2, bias=self.bias, dilation=2, down_sample=4)
        self.ec93 = CATConv(192, 64)
        self.x93 = CATConv(self.in_channel, 64)

        self.ec10 = SSEConv2(64, 64, self.out_channel2, bias=self.bias, down_sample=8)
        self.ec11 = SSEConv2(64, 64, self.out_channel2, bias=self.bias, down_sample=8)
        self.ec12 = SSEConv2(64, 64, self.out_channel2, bias=self.bias, down_sample=8)
        self.ec123 = CATConv(192, 64)

        self.pool0 = nn.MaxPool3d(kernel_size=[2, 2, 2], stride=[2, 2, 2], return_indices=False)
        self.pool1 = nn.MaxPool3d(kernel_size=[2, 2, 2], stride=[2, 2, 2], return_indices=False)
        self.pool2 = nn.MaxPool3d(kernel_size=[2, 2, 2], stride=[2, 2, 2], return_indices=False)

        self.up_sample0 = nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True)
        self.up_sample1 = nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True)
        self.up_sample2 = nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True)

        self.dc1 = SSEConv2(128, 64, self.out_channel2, bias=self.bias, down_sample=4)
        self.dc2 = SSEConv2(64, 64, self.out_channel2, bias=self.bias, down_sample=4)
        self.dc22 = CATConv(128, 64)
        self.dc3 = SSEConv2(128, 64, self.out_channel2, bias=self.bias, down_sample=2)
        self.dc4 = SSEConv2(64, 32, self.out_channel2, bias=self.bias, down_sample=2)
        self.dc42 = CATConv(96, 32)
        self.dc5 = SSEConv(64, 32, self.out_channel2, bias=self.bias, down_sample=1)
        self.dc6 = SSEConv(32, 16, self.out_channel2, bias=self.bias, down_sample=1)
        self.dc62 = CATConv(48, 16)

        self.dc0_0 = nn.Conv3d(24, n_classes, kernel_size=1, stride=1, padding=0, bias=self.bias)
        self.dc0_1 = nn.Conv3d(12, n_classes, kernel_size=1, stride=1, padding=0, bias=self.bias)
        self.dropout1 = DropLayer(channel_num=24, thr=0.3)
        self.dropout2 = DropLayer(channel_num=12, thr=0.3)

        # (owner module path, attribute) of every parameter in state_dict order; resolved with getattr
        # at call time so it also works on nn.DataParallel replicas (whose weights are plain tensors).
        self._param_paths = [tuple(n.rsplit('.', 1)) for n, _ in self.named_parameters()]
        self._dead_param_index = [i for i, (m, a) in enumerate(self._param_paths) if m == 'dc62.conv1'][0]
        object.__setattr__(self, '_rt', None)
        object.__setattr__(self, '_replica_rts', {})   # device index -> runtime of the DataParallel replica on it

    # ------------------------------------------------------------------------------------------
    # runtime state (not part of state_dict, not replicated: created lazily per replica/device)
    # ------------------------------------------------------------------------------------------
    class _Runtime:
        def __init__(self):
            self.plans = {}
            self.order = []
            self.flat = None       # flat fp32 copy of the 117 tensors (or the trainer-owned master buffer, see `shared`)
            self.flat_sig = None   # what `flat` was gathered from: ((data_ptr, _version) per tensor, ext_gen)
            self.gen = 0           # weight generation: +1 whenever the CONTENTS of `flat` may have changed
            self.ext_gen = 0       # bumped by whoever updates the parameters behind autograd's back (C-ABI AdamW)
            self.shared = None     # DataParallelTrainer's master buffer; the module's Parameters are views of it
            self.lock = threading.Lock()

    def _runtime(self):
        """Per-module runtime (plans, flat weights).  nn.DataParallel builds NEW replica objects on every forward, so a
        replica looks its runtime up by device in a registry shared with the master module: plans and workspaces are
        then created once per (device, shape), not once per forward (train.py:197, test.py:91, prediction.py:63)."""
        if self.__dict__.get('_is_replica'):
            idx = self._param_tensors()[0].device.index
            reg = self.__dict__['_replica_rts']
            rt = reg.get(idx)
            if rt is None:
                rt = reg.setdefault(idx, SE_UNet._Runtime())
            return rt
        rt = self.__dict__.get('_rt')
        if rt is None:
            rt = SE_UNet._Runtime()
            object.__setattr__(self, '_rt', rt)
        return rt

    def __getstate__(self):
        st = super().__getstate__() if hasattr(super(), '__getstate__') else self.__dict__.copy()
        st = dict(st)
        st['_rt'] = None
        st['_replica_rts'] = {}
        return st

    def __setstate__(self, st):
        super().__setstate__(st)
        self.__dict__.setdefault('_rt', None)
        self.__dict__.setdefault('_replica_rts', {})

    def _replicate_for_data_parallel(self):
        rep = super()._replicate_for_data_parallel()   # shallow __dict__ copy: `_replica_rts` stays the master's dict
        object.__setattr__(rep, '_rt', None)
        object.__setattr__(rep, '_is_replica', True)
        return rep

    def _params_changed(self):
        """Tell the module that parameter CONTENTS changed without their version counters moving (in-place update through
        the C ABI).  The next forward re-gathers the flat copy and re-packs the weight images."""
        self._runtime().ext_gen += 1

    def _param_tensors(self):
        out = []
        for mod_path, attr in self._param_paths:
            m = self
            for part in mod_path.split('.'):
                m = getattr(m, part)
            out.append(getattr(m, attr))
        return out

    def _weights(self, params=None):
        """(flat, gen): flat fp32 buffer of the 117 tensors in state_dict order on the compute device, and its weight
        generation.  Re-gathered when a parameter was replaced, moved or modified (version counters), or when
        `_params_changed()` was called; `gen` then advances, which is what `_Plan.pack` compares."""
        rt = self._runtime()
        if params is None:
            params = self._param_tensors()
        sig = (tuple((p.data_ptr(), p._version) for p in params), rt.ext_gen)
        if rt.flat is None or rt.flat_sig != sig or rt.flat.device != params[0].device:
            shared = rt.shared
            if shared is not None and _aliases_flat(params, shared):
                rt.flat = shared    # trainer-owned master buffer: the Parameters ARE views of it, nothing to gather
            else:
                rt.shared = None
                with torch.no_grad():
                    rt.flat = torch.cat([p.detach().reshape(-1).to(torch.float32) for p in params])
            rt.flat_sig = sig
            rt.gen += 1
        return rt.flat, rt.gen

    def _flat_params(self, params):
        return self._weights(params)[0]

    def _plan(self, batch, D, H, W, mode, device, slot=0):
        """Bound plan for this shape; `slot` selects independent workspaces (concurrent streams)."""
        rt = self._runtime()
        key = (batch, D, H, W, self.in_channel, self.n_classes, mode, device.index, slot)
        plan = rt.plans.get(key)
        if plan is None:
            while rt.order and (len(rt.order) >= _MAX_PLANS or
                                sum(pl.nbytes for pl in rt.plans.values()) > _MAX_PLAN_BYTES):
                rt.plans.pop(rt.order.pop(0), None)
            plan = _Plan(batch, D, H, W, self.in_channel, self.n_classes, mode, device)
            rt.plans[key] = plan
        if key in rt.order:
            rt.order.remove(key)
        rt.order.append(key)
        return plan

    def forward(self, x):
        """SE_UNet.py:181-238.  x: (B, in_channel, D, H, W) fp32 CUDA tensor (any strides), D/H/W multiples
        of 8.  Returns (pred0, pred1): raw fp32 logits of the encoder- and decoder-side heads."""
        if not x.is_cuda:
            raise _lib.SeunetError("SE_UNet (B200 build) runs on CUDA tensors only; there is no CPU path")
        if x.dim() != 5 or x.shape[1] != self.in_channel:
            raise ValueError(f"expected input of shape (B, {self.in_channel}, D, H, W), got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            x = x.float()
        params = self._param_tensors()
        if params[0].device != x.device:
            raise RuntimeError(f"module parameters are on {params[0].device}, input on {x.device}")
        B, _, D, H, W = x.shape
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        with torch.cuda.device(x.device):
            flat, gen = self._weights(params)
            plan = self._plan(B, D, H, W, 1 if need_grad else 0, x.device)
            plan.pack(flat, gen)
            drop0 = self.dropout1.scale(B, x.device)   # same RNG draw order as SE_UNet.py:232-233
            drop1 = self.dropout2.scale(B, x.device)
            if need_grad:
                return _SEUNetFunction.apply(self, plan, x, flat, drop0, drop1, *params)
            L = _lib.lib()
            pred0 = torch.empty((B, self.n_classes, D, H, W), dtype=torch.float32, device=x.device)
            pred1 = torch.empty_like(pred0)
            strides = (ctypes.c_int64 * 5)(*x.stride())
            _lib.check(L.seunet_forward(plan.handle, _lib.ptr(x), strides, None, _lib.ptr(flat), _lib.ptr(drop0),
                                        _lib.ptr(drop1), _lib.ptr(pred0), _lib.ptr(pred1), _lib.stream_ptr()),
                       "seunet_forward")
            return pred0, pred1


def get_model():
    """SE_UNet.py:240-242."""
    net = SE_UNet(in_channel=2)
    return config, net
