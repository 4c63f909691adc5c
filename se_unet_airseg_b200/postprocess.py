"""GPU post-processing of the sliding-window probability volume.

Mirrors the tail of the reference's `network_prediction` (prediction.py:111-116):

    pred = double_threshold_iteration(pred, h_thresh=0.5, l_thresh=0.4)      # prediction.py:13-37
    pred[0:int(0.15*X)] = 0; pred[int(0.85*X):] = 0; same for axis 1         # prediction.py:112-115
    pred_img = maximum_3d(pred)                                              # util.py:58-75

through the C ABI (`seunet_postproc_*`).  No CPU fallback: the CUDA library is required.
"""
import ctypes

import torch

from . import _lib


class PostProcessor:
    """Device-side hysteresis threshold + border crop + largest 26-connected component + hole filling for one volume shape."""

    def __init__(self, shape, device, max_runs=None):
        self.D, self.H, self.W = (int(s) for s in shape)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SeunetError("PostProcessor needs a CUDA device: there is no CPU fallback")
        L = _lib.lib()
        worst = self.D * self.H * ((self.W + 1) // 2) + 1
        self.max_runs = int(min(worst, 32 * 1024 * 1024) if max_runs is None else max_runs)
        nbytes = L.seunet_postproc_scratch_bytes(self.D, self.H, self.W, self.max_runs)
        if nbytes == 0:
            _lib.check(1, "seunet_postproc_scratch_bytes")
        self.scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.info = torch.zeros(16, dtype=torch.int32, device=self.device)

    def _check_prob(self, prob):
        if prob.dtype != torch.float32 or not prob.is_contiguous() or tuple(prob.shape) != (self.D, self.H, self.W):
            raise ValueError(f"expected a contiguous float32 volume of shape {(self.D, self.H, self.W)}")
        if prob.device != self.device:
            raise ValueError("volume is on another device")

    def dti(self, prob, h_thresh=0.5, l_thresh=0.4, border_frac=None, want_mask=True):
        """double_threshold_iteration (+ optional border zeroing); returns the uint8 mask (or None) and keeps the bit-packed
        result on the device for `largest_component()`."""
        self._check_prob(prob)
        L = _lib.lib()
        out = torch.empty((self.D, self.H, self.W), dtype=torch.uint8, device=self.device) if want_mask else None
        with torch.cuda.device(self.device):
            _lib.check(L.seunet_postproc_dti(_lib.ptr(prob), self.D, self.H, self.W, float(h_thresh), float(l_thresh),
                                             -1.0 if border_frac is None else float(border_frac), _lib.ptr(out),
                                             _lib.ptr(self.scratch), self.max_runs, _lib.stream_ptr()), "seunet_postproc_dti")
        return out

    def largest_component(self, mask=None, fill_holes=True):
        """maximum_3d of `mask` (uint8, None = the result of the preceding `dti` call)."""
        L = _lib.lib()
        if mask is not None:
            if mask.dtype != torch.uint8 or not mask.is_contiguous() or tuple(mask.shape) != (self.D, self.H, self.W):
                raise ValueError("expected a contiguous uint8 mask of the processor's shape")
        out = torch.empty((self.D, self.H, self.W), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(L.seunet_postproc_largest_component(_lib.ptr(mask), self.D, self.H, self.W, 1 if fill_holes else 0,
                                                           _lib.ptr(out), _lib.ptr(self.info), _lib.ptr(self.scratch),
                                                           self.max_runs, _lib.stream_ptr()),
                       "seunet_postproc_largest_component")
        info = self.info.cpu()   # one small D2H; also the point where an overflow of the run table is reported
        if int(info[0]) != 0:
            raise _lib.SeunetError(f"post-processing saw more than max_runs={self.max_runs} row runs; pass a larger max_runs")
        self.last_info = {"runs": int(info[1]), "largest": int(info[2]), "second": int(info[3]), "used_second": bool(info[4])}
        return out

    def __call__(self, prob, h_thresh=0.5, l_thresh=0.4, border_frac=0.15):
        """prediction.py:111-116 on the device: mean probability volume -> final uint8 airway mask."""
        self.dti(prob, h_thresh, l_thresh, border_frac, want_mask=False)
        return self.largest_component(None, fill_holes=True)
