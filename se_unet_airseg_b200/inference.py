"""GPU sliding-window inference driver, the B200-native counterpart of the loop in
the reference's prediction.py:65-111:

    img - 1024 -> two HU windows -> 128^3 windows at stride 64 (last window clamped) -> SE_UNet forward
    (eval mode) -> sigmoid -> overlap mean -> >= 0.5

Differences in HOW (not WHAT): the windows of one resident volume are batched into one forward (the C ABI
takes per-sample offsets into the volume, no gather copy), probabilities are accumulated on the device in
32-bit fixed point (order-independent, see csrc/window.cu) and the window-count volume is analytic; only the
final mask crosses PCIe.

Multi-GPU (BASELINE config 4): under torchrun the windows of ONE volume are sharded by patch -
every rank takes a contiguous range of the window list (contiguous slabs along the first axis), accumulates
into its own partial volume; the partial planes then go point-to-point (NCCL send/recv over NVLink/NVSwitch, all
pairs at once) to the rank that owns them, every rank finalizes its 1/N of the planes and rank 0 collects the
uint8 mask (`predict_sharded` / `predict_device_sharded`, `exchange_partials`).  Because the accumulation is
integer, the N-rank mask is bit-identical to the 1-rank mask.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def window_starts(length, cube=128, step=64):
    """Window origins along one axis, exactly prediction.py:80-100 (the last window is clamped)."""
    if length < cube:
        raise ValueError(f"axis of length {length} is shorter than the window ({cube})")
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


def split_batches(n, batch):
    """Sizes of the forward batches for n windows: ceil(n / batch) batches of (almost) equal size - at most two distinct
    sizes, so at most two plan shapes per stream slot."""
    if n <= 0:
        return []
    k = -(-n // batch)
    return [n // k + (1 if i < n % k else 0) for i in range(k)]


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of the window list owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_ranges(length, world):
    """Planes along the first axis that each rank FINALIZES in the sharded mode: an even contiguous partition."""
    return [(length * q // world, length * (q + 1) // world) for q in range(world)]


def _peer(group, q):
    return dist.get_global_rank(group, q) if group is not None else q


def exchange_partials(acc, slabs, owners, rank, world, group=None, staging=None):
    """The exchange step of patch-sharded inference.  acc: this rank's partial fixed-point volume (X, Y, Z) int32, non-zero
    only inside slabs[rank] = the planes its windows touched.  Every rank q owns the planes owners[q]; rank r sends q the
    intersection of its slab with q's planes (contiguous memory: the first axis is the outermost) and adds what it receives
    to its own planes - point-to-point over NVLink/NVSwitch, all pairs at once (one NCCL group).  On return
    acc[owners[rank]] holds the sums over ALL ranks (integer adds: the result does not depend on the order); the rest of
    acc is stale.  Compared with one SUM reduce of the whole volume to rank 0 this moves each partial plane once, to the
    rank that needs it, instead of funnelling world x 420 MB through a reduction tree."""
    x0, x1 = owners[rank]
    ops, recvs = [], []
    for q in range(world):
        if q == rank:
            continue
        a, b = max(slabs[rank][0], owners[q][0]), min(slabs[rank][1], owners[q][1])
        if a < b:
            ops.append(dist.P2POp(dist.isend, acc[a:b], _peer(group, q), group))
        a, b = max(slabs[q][0], x0), min(slabs[q][1], x1)
        if a < b:
            buf = staging.get((q, a, b)) if staging is not None else None
            if buf is None:
                buf = torch.empty_like(acc[a:b])
                if staging is not None:
                    staging[(q, a, b)] = buf
            recvs.append((a, b, buf))
            ops.append(dist.P2POp(dist.irecv, buf, _peer(group, q), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for a, b, buf in recvs:
        acc[a:b] += buf


def gather_planes(buf, owners, rank, world, group=None, root=0):
    """Rank `root` receives buf[owners[q]] from every other rank q (straight into its own buf), the others send."""
    ops = []
    if rank == root:
        for q in range(world):
            a, b = owners[q]
            if q != root and a < b:
                ops.append(dist.P2POp(dist.irecv, buf[a:b], _peer(group, q), group))
    else:
        a, b = owners[rank]
        if a < b:
            ops.append(dist.P2POp(dist.isend, buf[a:b], _peer(group, root), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def coverage_counts(length, starts, cube):
    c = np.zeros(length, dtype=np.int32)
    for s in starts:
        c[s:s + cube] += 1
    return c


class SlidingWindowPredictor:
    """Runs `model` (se_unet_airseg_b200.SE_UNet on a CUDA device, eval mode like prediction.py:64) over a CT volume."""

    def __init__(self, model, cube=128, step=64, batch=7, threshold=0.5, streams=3, fuse_head=True):
        """streams > 1: consecutive window batches run on different CUDA streams with their own plan workspaces, so the
        HBM-bound passes of one batch overlap the tensor-bound convolutions of the other (both fit on an SM together).
        fuse_head: one `seunet_forward_window` call per batch - prediction.py:103 discards the first output p0, so its whole
        head is not computed, and sigmoid(p) goes from the head kernel straight into the accumulator volume.  False: the
        two-call path (seunet_forward, seunet_window_accumulate); both give bit-identical volumes."""
        self.model = model
        self.fuse_head = fuse_head
        self.cube, self.step, self.batch, self.threshold = cube, step, batch, threshold
        self.nstreams = max(1, streams)
        self._streams = None
        self._geom = None

    def _geometry(self, shape, device):
        if self._geom is not None and self._geom[0] == (tuple(shape), device):
            return self._geom[1]
        X, Y, Z = shape
        sx, sy, sz = (window_starts(n, self.cube, self.step) for n in (X, Y, Z))
        wins = [(a, b, c) for a in sx for b in sy for c in sz]  # same nesting order as prediction.py:83-100
        counts = np.concatenate([coverage_counts(X, sx, self.cube), coverage_counts(Y, sy, self.cube),
                                 coverage_counts(Z, sz, self.cube)]).astype(np.int32)
        cmax = int(coverage_counts(X, sx, self.cube).max()) * int(coverage_counts(Y, sy, self.cube).max()) * \
            int(coverage_counts(Z, sz, self.cube).max())
        acc_log2 = min(26, 30 - int(np.ceil(np.log2(cmax + 1))))      # cmax * 2^acc_log2 < 2^31 (NCCL sums it as int32)
        if acc_log2 < 16:
            raise ValueError(f"window grid overlaps {cmax} times per voxel: too dense for the fixed-point accumulator")
        g = dict(wins=wins, counts=torch.from_numpy(counts).to(device), acc_log2=acc_log2,
                 acc=torch.empty((X, Y, Z), dtype=torch.int32, device=device),
                 mask=torch.empty((X, Y, Z), dtype=torch.uint8, device=device),
                 x2=torch.empty((1, 2, X, Y, Z), dtype=torch.float32, device=device))
        self._geom = ((tuple(shape), device), g)
        return g

    @torch.no_grad()
    def predict_device(self, img_dev, hu_offset=-1024.0, return_prob=False, reuse_output=False, _slab_events=None,
                       _shard=None, _host_out=None):
        """img_dev: (X, Y, Z) int16 or fp32 CUDA tensor holding the stored CT values (HU + 1024, prediction.py:68-69).
        Returns the uint8 mask (X, Y, Z) on the device (and the mean probability if return_prob).
        The results are fresh tensors unless reuse_output=True, which hands out the predictor's own accumulator / mask
        buffers: they are OVERWRITTEN by the next call on a volume of the same shape (zero-copy mode for streaming use).
        _slab_events (internal, used by predict()): [(x_end, event)] - the two-HU-window input has already been produced slab
        by slab on a copy stream; a window batch waits only for the slabs it reads.
        _shard (internal, used by the *_sharded entry points): (rank, world, group) - run only this rank's range of the
        window list, exchange the partial planes with their owner ranks (exchange_partials), finalize the owned planes and
        collect the mask on rank 0; ranks != 0 return None.
        _host_out (internal, used by predict()): pinned uint8 host tensor - planes are divided, thresholded and copied to the
        host as soon as the last window that touches them has been issued (the window list is ordered along the first axis),
        so only the last slab's D2H copy is left after the last forward."""
        L = _lib.lib()
        m = self.model
        if m.in_channel != 2:
            raise ValueError("sliding-window CT inference needs the two-HU-window model (in_channel=2)")
        if m.training:
            raise RuntimeError("SlidingWindowPredictor mirrors prediction.py: call model.eval() first")
        dev = img_dev.device
        X, Y, Z = img_dev.shape
        g = self._geometry((X, Y, Z), dev)
        st = _lib.stream_ptr()
        dtype = {torch.int16: 0, torch.float32: 1}[img_dev.dtype]
        if _slab_events is None:
            img_dev = img_dev.contiguous()
            _lib.check(L.seunet_hu_windows(_lib.ptr(img_dev), dtype, X * Y * Z, float(hu_offset), _lib.ptr(g["x2"]), st),
                       "seunet_hu_windows")
        sharded = _shard is not None and _shard[1] > 1
        if sharded:
            rank, world, group = _shard
            slabs = [self.shard_planes((X, Y, Z), r, world) for r in range(world)]
            owners = owner_ranges(X, world)
            own0, own1 = owners[rank]
            z0, z1 = (min(slabs[rank][0], own0), max(slabs[rank][1], own1)) if slabs[rank][1] > slabs[rank][0] else (own0, own1)
            g["acc"][z0:z1].zero_()          # only the planes this rank accumulates into or finalizes
        else:
            g["acc"].zero_()
        x2 = g["x2"]
        sN, sC, sD, sH, sW = x2.stride()
        cube = self.cube
        wins = g["wins"]
        if _shard is not None:
            lo, hi = shard_range(len(wins), _shard[0], _shard[1])
            wins = wins[lo:hi]
        params = m._param_tensors()
        flat, wgen = m._weights(params)
        main = torch.cuda.current_stream(dev)
        if self._streams is None or self._streams[0].device != dev:
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(self.nstreams)] if self.nstreams > 1 else [main]
        streams = self._streams
        if self.nstreams > 1:
            ready = torch.cuda.Event()
            ready.record(main)
            for s_ in streams:
                s_.wait_event(ready)
        early = None
        if _host_out is not None and not sharded:
            out_stream = getattr(self, "_out_stream", None)
            if out_stream is None or out_stream.device != dev:
                out_stream = torch.cuda.Stream(device=dev)
                self._out_stream = out_stream
            early = dict(x=0, stream=out_stream, host=_host_out)
        i = 0
        for k, b in enumerate(split_batches(len(wins), self.batch)):
            slot = k % len(streams)
            cs = streams[slot]
            if _slab_events is not None:
                need = max(w[0] for w in wins[i:i + b]) + cube
                for x_end, ev in _slab_events:          # every slab below the highest plane this batch reads
                    cs.wait_event(ev)
                    if x_end >= need:
                        break
            with torch.cuda.stream(cs):
                stp = ctypes.c_void_p(cs.cuda_stream)
                plan = m._plan(b, cube, cube, cube, 0, dev, slot=slot)
                plan.pack(flat, wgen)
                ones0, ones1, pred0, pred1 = self._buffers(plan, b, dev)
                offs = (ctypes.c_int64 * b)(*[w[0] * sD + w[1] * sH + w[2] * sW for w in wins[i:i + b]])
                strides = (ctypes.c_int64 * 5)(0, sC, sD, sH, sW)
                starts = (ctypes.c_int * (3 * b))(*[v for w in wins[i:i + b] for v in w])
                if self.fuse_head:
                    _lib.check(L.seunet_forward_window(plan.handle, _lib.ptr(x2), strides, offs, _lib.ptr(flat), _lib.ptr(ones0),
                                                       _lib.ptr(ones1), starts, _lib.ptr(g["acc"]), X, Y, Z, g["acc_log2"], stp),
                               "seunet_forward_window")
                else:
                    _lib.check(L.seunet_forward(plan.handle, _lib.ptr(x2), strides, offs, _lib.ptr(flat), _lib.ptr(ones0),
                                                _lib.ptr(ones1), _lib.ptr(pred0), _lib.ptr(pred1), stp), "seunet_forward")
                    _lib.check(L.seunet_window_accumulate(_lib.ptr(pred1), starts, b, cube, cube, cube, _lib.ptr(g["acc"]),
                                                          X, Y, Z, 1, g["acc_log2"], stp), "seunet_window_accumulate")
            i += b
            if early is not None:
                # planes below the first-axis origin of the next window are complete once everything issued so far has run
                x_done = wins[i][0] if i < len(wins) else X
                if x_done > early["x"]:
                    self._finalize_slab(g, early, x_done, streams[:min(k + 1, len(streams))], return_prob)
        if self.nstreams > 1:
            for s_ in streams:
                done = torch.cuda.Event()
                done.record(s_)
                main.wait_event(done)
        if early is not None:
            main.wait_stream(early["stream"])
            prob = g["acc"].view(torch.float32)
            if reuse_output:
                return (g["mask"], prob) if return_prob else g["mask"]
            return (g["mask"].clone(), prob.clone()) if return_prob else g["mask"].clone()
        if sharded:
            # the exchange step of patch sharding (NCCL point-to-point over NVLink/NVSwitch): partial planes go to the rank
            # that finalizes them; every rank divides / thresholds its own planes; rank 0 collects the uint8 mask planes
            exchange_partials(g["acc"], slabs, owners, rank, world, group, g.setdefault("staging", {}))
            key = ("counts_own", rank, world)
            if key not in g:
                cx = g["counts"][own0:own1]
                g[key] = torch.cat([cx, g["counts"][X:]]).contiguous()
            if own1 > own0:
                plane = Y * Z
                _lib.check(L.seunet_window_finalize(ctypes.c_void_p(g["acc"].data_ptr() + own0 * plane * 4), _lib.ptr(g[key]),
                                                    own1 - own0, Y, Z, float(self.threshold),
                                                    ctypes.c_void_p(g["mask"].data_ptr() + own0 * plane),
                                                    1 if return_prob else 0, g["acc_log2"], st), "seunet_window_finalize")
            gather_planes(g["mask"], owners, rank, world, group)
            if return_prob:
                gather_planes(g["acc"], owners, rank, world, group)      # fp32 means in the accumulator's 4-byte slots
            if rank != 0:
                return None
            prob = g["acc"].view(torch.float32)
            if reuse_output:
                return (g["mask"], prob) if return_prob else g["mask"]
            return (g["mask"].clone(), prob.clone()) if return_prob else g["mask"].clone()
        _lib.check(L.seunet_window_finalize(_lib.ptr(g["acc"]), _lib.ptr(g["counts"]), X, Y, Z, float(self.threshold),
                                            _lib.ptr(g["mask"]), 1 if return_prob else 0, g["acc_log2"], st),
                   "seunet_window_finalize")
        prob = g["acc"].view(torch.float32)     # finalize(write_mean) replaced the fixed-point sums by the fp32 means
        if reuse_output:
            return (g["mask"], prob) if return_prob else g["mask"]
        return (g["mask"].clone(), prob.clone()) if return_prob else g["mask"].clone()

    def _finalize_slab(self, g, early, x_done, used_streams, write_mean):
        """Divide / threshold planes [early.x, x_done) on the output stream once the work issued so far on `used_streams` has
        finished, and start their D2H copy (prediction.py:109-110 applied slab by slab: the operations are per voxel)."""
        L = _lib.lib()
        x0, out_stream = early["x"], early["stream"]
        X, Y, Z = g["acc"].shape
        for s_ in used_streams:
            ev = torch.cuda.Event()
            ev.record(s_)
            out_stream.wait_event(ev)
        key = ("counts_slab", x0, x_done)
        if key not in g:
            g[key] = torch.cat([g["counts"][x0:x_done], g["counts"][X:]]).contiguous()
        plane = Y * Z
        with torch.cuda.stream(out_stream):
            _lib.check(L.seunet_window_finalize(ctypes.c_void_p(g["acc"].data_ptr() + x0 * plane * 4), _lib.ptr(g[key]),
                                                x_done - x0, Y, Z, float(self.threshold),
                                                ctypes.c_void_p(g["mask"].data_ptr() + x0 * plane), 1 if write_mean else 0,
                                                g["acc_log2"], ctypes.c_void_p(out_stream.cuda_stream)), "seunet_window_finalize")
            early["host"][x0:x_done].copy_(g["mask"][x0:x_done], non_blocking=True)
        early["x"] = x_done

    # ------------------------------------------------------------------------------------------
    # patch-sharded inference of one volume over the ranks of a process group (one process per GPU)
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _shard_info(group):
        if not dist.is_initialized():
            return (0, 1, group)
        return (dist.get_rank(group), dist.get_world_size(group), group)

    def shard_planes(self, shape, rank, world):
        """[x_lo, x_hi): planes along the first axis that `rank`'s windows read (what it needs on its device)."""
        sx, sy, sz = (window_starts(n, self.cube, self.step) for n in shape)
        wins = [(a, b, c) for a in sx for b in sy for c in sz]
        lo, hi = shard_range(len(wins), rank, world)
        if lo == hi:
            return 0, 0
        return min(w[0] for w in wins[lo:hi]), max(w[0] for w in wins[lo:hi]) + self.cube

    @torch.no_grad()
    def predict_device_sharded(self, img_dev, hu_offset=-1024.0, return_prob=False, reuse_output=False, group=None):
        """Collective call (every rank of `group`): img_dev is this rank's device copy of the stored CT volume - only the
        planes of `shard_planes()` are read.  Rank 0 returns the mask (and mean probability), the others None."""
        return self.predict_device(img_dev, hu_offset, return_prob, reuse_output, _shard=self._shard_info(group))

    @torch.no_grad()
    def predict_sharded(self, img_host, hu_offset=-1024.0, slab=64, reuse_output=False, group=None):
        """Collective end-to-end call: every rank passes the same host volume (e.g. the memory-mapped NIfTI the ranks of
        one node share); each rank copies only the planes its windows read, rank 0 returns the host mask."""
        return self.predict(img_host, hu_offset, slab, reuse_output, _shard=self._shard_info(group))

    @torch.no_grad()
    def predict_postprocessed_device(self, img_dev, hu_offset=-1024.0, h_thresh=0.5, l_thresh=0.4, border_frac=0.15):
        """The whole of prediction.py:78-116 on the device: sliding-window mean probability, double-threshold hysteresis,
        border crop, largest 26-connected component, hole filling.  Returns the final uint8 mask (X, Y, Z)."""
        from .postprocess import PostProcessor
        _, prob = self.predict_device(img_dev, hu_offset, return_prob=True, reuse_output=True)   # consumed before returning
        pp = getattr(self, "_post", None)
        if pp is None or (pp.D, pp.H, pp.W) != tuple(prob.shape) or pp.device != prob.device:
            pp = PostProcessor(tuple(prob.shape), prob.device)
            self._post = pp
        return pp(prob, h_thresh, l_thresh, border_frac)

    def _buffers(self, plan, b, dev):
        buf = getattr(plan, "_sw_buffers", None)
        if buf is None:
            c = self.cube
            buf = (torch.ones(b, 24, device=dev), torch.ones(b, 12, device=dev),
                   torch.empty((b, 1, c, c, c), dtype=torch.float32, device=dev),
                   torch.empty((b, 1, c, c, c), dtype=torch.float32, device=dev))
            plan._sw_buffers = buf
        return buf

    @torch.no_grad()
    def predict(self, img_host, hu_offset=-1024.0, slab=64, reuse_output=False, _shard=None):
        """End-to-end call a user makes: host volume (numpy int16/float32 or CPU tensor, ideally pinned) in, host uint8
        mask out.  The H2D copy and the HU windowing run slab by slab (along the first axis) on a copy stream, and every window
        batch waits only for the slabs it reads, so the forward passes start after the first ~128 planes have arrived;
        the mask goes back slab by slab as well: planes are finalized and copied to the host as soon as their last window
        has run, only the last slab's copy follows the last forward.
        Returns a fresh CPU tensor.  reuse_output=True returns the predictor's pinned staging buffer instead (no host
        copy); it is OVERWRITTEN by the next predict() of a same-shaped volume - only for callers that consume the mask
        before the next call (prediction.py does: it writes the NIfTI, then moves on)."""
        L = _lib.lib()
        t = torch.from_numpy(img_host) if isinstance(img_host, np.ndarray) else img_host
        dev = next(p for p in self.model._param_tensors()).device
        if t.dtype not in (torch.int16, torch.float32) or t.dim() != 3:
            raise ValueError("expected a 3-D int16 or float32 volume")
        t = t.contiguous()
        X, Y, Z = t.shape
        g = self._geometry((X, Y, Z), dev)
        dtype = {torch.int16: 0, torch.float32: 1}[t.dtype]
        stage = getattr(self, "_dev_img", None)
        if stage is None or stage.shape != t.shape or stage.dtype != t.dtype or stage.device != dev:
            stage = torch.empty(t.shape, dtype=t.dtype, device=dev)
            self._dev_img = stage
        copy_stream = getattr(self, "_copy_stream", None)
        if copy_stream is None or copy_stream.device != dev:
            copy_stream = torch.cuda.Stream(device=dev)
            self._copy_stream = copy_stream
        main = torch.cuda.current_stream(dev)
        start = torch.cuda.Event()
        start.record(main)                     # the previous volume's windows must be done with x2 before it is overwritten
        copy_stream.wait_event(start)
        events = []
        plane = Y * Z
        xa, xb = (0, X) if _shard is None else self.shard_planes((X, Y, Z), _shard[0], _shard[1])
        with torch.cuda.stream(copy_stream):
            for x0 in range(xa, xb, slab):
                x1 = min(xb, x0 + slab)
                stage[x0:x1].copy_(t[x0:x1], non_blocking=True)
                _lib.check(L.seunet_hu_windows_slab(_lib.ptr(stage[x0:x1]), dtype, (x1 - x0) * plane, X * plane, float(hu_offset),
                                                    ctypes.c_void_p(g["x2"].data_ptr() + x0 * plane * 4),
                                                    ctypes.c_void_p(copy_stream.cuda_stream)), "seunet_hu_windows_slab")
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                events.append((x1, ev))
        out = getattr(self, "_host_mask", None)
        single = _shard is None or _shard[1] <= 1
        if (single or _shard[0] == 0) and (out is None or tuple(out.shape) != (X, Y, Z)):
            out = torch.empty((X, Y, Z), dtype=torch.uint8, pin_memory=True)
            self._host_mask = out
        mask = self.predict_device(stage, hu_offset, reuse_output=True, _slab_events=events, _shard=_shard,
                                   _host_out=out if single else None)
        if mask is None:            # sharded call on a rank != 0
            torch.cuda.current_stream(dev).synchronize()
            return None
        if not single:
            out.copy_(mask, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()   # (single GPU: the slab copies were joined into this stream)
        return out if reuse_output else out.clone()
