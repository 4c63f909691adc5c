"""Data-parallel training step for SE_UNet on B200: one process per GPU (torch.distributed / NCCL over NVLink), the
B200-native replacement of the reference's single-process `torch.nn.DataParallel` loop (train.py:428-440, 234-247, 592-603).

Per step and rank:
  forward (C ABI, training plan)  ->  loss partial sums (fused kernel)  ->  all-reduce of 16 fp64 sums [C4]
  ->  loss gradient from the GLOBAL sums  ->  backward (C ABI)  ->  all-reduce(SUM) of the flat fp32 gradient [C3]
  ->  fused AdamW on the flat parameter buffer  ->  re-pack the tensor-core weight image (next forward).

graph=True replays everything up to and including the gradient all-reduce (~250 launches, both NCCL exchanges) as ONE CUDA
graph per input shape: the inputs are copied into trainer-owned static buffers, the DropLayer draws stay on the host and are
copied in before the replay, AdamW (whose bias correction depends on the step number) runs eagerly behind it.  At one patch
per rank - the 8-GPU end of the data-parallel curve - the launch gaps between ~250 short kernels are 6 % of the step
(tools/graph_step.py: 6.12 -> 5.76 ms at 1 x 128^3, 38.8 -> 38.6 ms at 8 x 128^3).

The loss is a ratio of batch-global sums (train.py:51-76 evaluated on the gathered batch), so partial sums - not
per-rank losses - are exchanged, and gradients are SUMMED, not averaged: N ranks x B/N patches reproduce the single-rank
step on the concatenated batch.  (DropLayer's normaliser is per local batch in the reference too - SE_UNet.py:94 under
DataParallel - so train-mode numerics depend on the GPU count there as well.)
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib

ADAMW_DEFAULTS = dict(lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)   # torch.optim.AdamW(lr=1e-4), train.py:188


def loss_from_sums(stage, sums):
    """Scalar stage loss from the [2][8] global sums (host-side mirror of seunet_loss_grad, for logging/tests)."""
    e, d = sums[0], sums[1]
    dice = lambda s: 1.0 - (2.0 * s[0] + 1.0) / (s[1] + s[2] + 1.0)
    gul = lambda s: 1.0 - (s[3] + 1.0) / (s[4] + 1.0)
    atr = lambda s: 1.0 - (s[5] + 1.0) / (s[6] + 1.0)
    if stage == 1:
        return dice(d) + dice(e)
    loss = gul(d) + 0.5 * gul(e)
    if stage == 3:
        loss = loss + 0.5 * (atr(e) + atr(d))
    return loss


class DataParallelTrainer:
    def __init__(self, model, stage=2, graph=False, **adamw):
        """stage: 1 | 2 | 3 (train.py's curriculum: Dice; LIB-weighted union loss; + skeleton term).  graph: replay the step as
        one CUDA graph per input shape (see the module docstring).  adamw: lr / betas / eps / weight_decay overrides."""
        self.model, self.stage = model, stage
        self.graph_mode = bool(graph)
        self._graphs = {}
        self.hp = dict(ADAMW_DEFAULTS, **adamw)
        self.step_count = 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        params = model._param_tensors()
        self.device = params[0].device
        if self.device.type != "cuda":
            raise _lib.SeunetError("DataParallelTrainer needs the model on a CUDA device")
        self._adopt(params)
        self.grads = torch.zeros_like(self.flat)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.sums = torch.zeros(16, dtype=torch.float64, device=self.device)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        L = _lib.lib()
        ic, nc = model.in_channel, model.n_classes
        self.skip_off = L.seunet_param_offset(ic, nc, b"dc62.conv1.weight")
        self.skip_len = 48 * 16
        self._buf = None

    def _adopt(self, params):
        """One flat fp32 master copy; the module's Parameters become views of it so state_dict()/checkpoints stay live and
        the fused AdamW updates them in place.  The buffer is registered with the module's runtime (`shared`): eval-mode
        forwards and the sliding-window predictor then read the SAME memory, and `model._params_changed()` after every step
        advances the weight generation that guards the packed tensor-core images."""
        with torch.no_grad():
            flat = torch.cat([p.detach().reshape(-1).float() for p in params]).contiguous()
            off = 0
            for p in params:
                n = p.numel()
                p.data = flat[off:off + n].view(p.shape)
                off += n
        self.flat = flat
        rt = self.model._runtime()
        rt.shared = flat
        rt.flat = None      # force the next _weights() to re-resolve

    def _check_input(self, name, t, shape):
        if t is None:
            return None
        if not torch.is_tensor(t) or t.device != self.device:
            raise ValueError(f"{name}: expected a tensor on {self.device}, got {getattr(t, 'device', type(t))}")
        if tuple(t.shape) != shape:
            raise ValueError(f"{name}: expected shape {shape}, got {tuple(t.shape)}")
        return t.to(torch.float32).contiguous()   # no-ops for the fp32 contiguous tensors the loaders produce

    def _buffers(self, shape):
        if self._buf is None or self._buf[0] != shape:
            B, D, H, W = shape
            mk = lambda: torch.empty((B, 1, D, H, W), dtype=torch.float32, device=self.device)
            self._buf = (shape, mk(), mk(), mk(), mk(), torch.zeros(B * 4, dtype=torch.float64, device=self.device))
        return self._buf[1:]

    def step(self, x, label, weight=None, skel=None):
        """One optimisation step on this rank's shard of the batch; returns the (global) loss as a 1-element device tensor
        (no host synchronisation).  x: (B, in_ch, D, H, W) fp32 CUDA; label/weight/skel: (B, 1, D, H, W) fp32 CUDA."""
        L = _lib.lib()
        m = self.model
        if not torch.is_tensor(x) or x.dim() != 5 or x.shape[1] != m.in_channel or x.device != self.device:
            raise ValueError(f"x: expected (B, {m.in_channel}, D, H, W) on {self.device}, got "
                             f"{tuple(x.shape) if torch.is_tensor(x) else type(x)}")
        if x.dtype != torch.float32:
            x = x.float()
        B, _, D, H, W = x.shape
        label = self._check_input("label", label, (B, 1, D, H, W))
        weight = self._check_input("weight", weight, (B, 1, D, H, W))
        skel = self._check_input("skel", skel, (B, 1, D, H, W))
        if label is None or (self.stage >= 2 and weight is None) or (self.stage == 3 and skel is None):
            raise ValueError(f"stage {self.stage} needs label" + (", weight" if self.stage >= 2 else "") +
                             (", skel" if self.stage == 3 else ""))
        with torch.cuda.device(self.device), torch.no_grad():
            params = m._param_tensors()
            flat, wgen = m._weights(params)     # notices load_state_dict / external edits through the version counters
            if flat is not self.flat:           # parameters were re-created or moved: adopt the new tensors
                self._adopt(params)
                self.m.zero_(); self.v.zero_(); self.step_count = 0
                self._graphs.clear()            # captured graphs point into the old buffers
                flat, wgen = m._weights(params)
            plan = m._plan(B, D, H, W, 1, self.device)
            drop0 = m.dropout1.scale(B, self.device)
            drop1 = m.dropout2.scale(B, self.device)
            if self.graph_mode:
                self._step_graph(plan, wgen, x, label, weight, skel, drop0, drop1)
            else:
                self._enqueue(plan, wgen, x, label, weight, skel, drop0, drop1)
            self.step_count += 1
            hp = self.hp
            p_ = _lib.ptr
            _lib.check(L.seunet_adamw_step(p_(self.flat), p_(self.grads), p_(self.m), p_(self.v), self.flat.numel(), hp["lr"],
                                           hp["betas"][0], hp["betas"][1], hp["eps"], hp["weight_decay"], self.step_count, 1.0,
                                           self.skip_off, self.skip_len, _lib.stream_ptr()), "seunet_adamw_step")
            m._params_changed()   # the in-place C-ABI update is invisible to the tensors' version counters
        self.per_sample_gul = self._buffers((B, D, H, W))[4]
        return self.loss

    def _enqueue(self, plan, wgen, x, label, weight, skel, drop0, drop1):
        """Everything of a step up to the summed gradient, enqueued on the current stream (capturable: no allocation, no host
        synchronisation; the first call per shape allocates the prediction / gradient buffers)."""
        L = _lib.lib()
        p_ = _lib.ptr
        st = _lib.stream_ptr()
        B, _, D, H, W = x.shape
        plan.pack(self.flat, wgen)
        plan.generation += 1
        pe, pd, ge, gd, per_sample = self._buffers((B, D, H, W))
        strides = (ctypes.c_int64 * 5)(*x.stride())
        _lib.check(L.seunet_forward(plan.handle, p_(x), strides, None, p_(self.flat), p_(drop0), p_(drop1), p_(pe), p_(pd), st),
                   "seunet_forward")
        V = D * H * W
        _lib.check(L.seunet_loss_sums(self.stage, p_(pe), p_(pd), p_(label), p_(weight), p_(skel), B, V, p_(self.sums),
                                      p_(per_sample), st), "seunet_loss_sums")
        if self.world > 1:
            dist.all_reduce(self.sums, op=dist.ReduceOp.SUM)          # C4: batch-global loss sums
        _lib.check(L.seunet_loss_grad(self.stage, p_(pe), p_(pd), p_(label), p_(weight), p_(skel), B * V, p_(self.sums),
                                      p_(ge), p_(gd), p_(self.loss), st), "seunet_loss_grad")
        _lib.check(L.seunet_backward(plan.handle, p_(x), strides, None, p_(self.flat), p_(drop0), p_(drop1), p_(ge), p_(gd),
                                     p_(self.grads), st), "seunet_backward")
        if self.world > 1:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM)         # C3: one flat 6.08 MB bucket, SUM (not mean)

    def _step_graph(self, plan, wgen, x, label, weight, skel, drop0, drop1):
        """graph=True: copy this step's inputs into the static buffers of their shape, then replay the captured step.  The
        first step of a shape runs eagerly on the static buffers (allocations, per-device kernel attributes, side streams),
        the second one is captured."""
        key = (tuple(x.shape), self.flat.data_ptr(), plan.handle.value if hasattr(plan.handle, "value") else id(plan))
        g = self._graphs.get(key)
        if g is None:
            mk = lambda t: None if t is None else torch.empty_like(t, memory_format=torch.contiguous_format)
            g = dict(x=mk(x), label=mk(label), weight=mk(weight), skel=mk(skel), drop0=mk(drop0), drop1=mk(drop1), graph=None, warm=0,
                     plan=plan)             # (the reference keeps the plan's workspace alive if the module's plan cache evicts it)
            while len(self._graphs) >= 4:   # a graph pins its static inputs and its plan's workspace: keep the last few shapes
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = g
        for name, src in (("x", x), ("label", label), ("weight", weight), ("skel", skel), ("drop0", drop0), ("drop1", drop1)):
            if src is not None:
                g[name].copy_(src, non_blocking=True)
        args = (plan, wgen, g["x"], g["label"], g["weight"], g["skel"], g["drop0"], g["drop1"])
        if g["graph"] is not None:
            g["graph"].replay()
            plan.packed_gen = wgen          # the replay re-packed the weight image from the current parameters
            plan.generation += 1
        elif g["warm"] == 0:
            self._enqueue(*args)
            g["warm"] = 1
        else:
            graph = torch.cuda.CUDAGraph()
            plan.packed_gen = None          # the capture must contain the re-pack: the weights change every step
            with torch.cuda.graph(graph):
                self._enqueue(*args)
            graph.replay()                  # (capturing does not execute)
            plan.packed_gen = wgen
            g["graph"] = graph
