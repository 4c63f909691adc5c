"""GPU sliding-window driver (HU windows -> batched windows -> device-side overlap mean -> mask) against the
oracle's restatement of prediction.py:65-111 on a small synthetic CT volume (cube 32 / stride 16, ragged axes)."""
import numpy as np
import pytest
import torch

from oracle import seunet_oracle as oracle

pytestmark = pytest.mark.gpu


def _synthetic_ct(shape, seed):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(shape, generator=g) * 400.0 + 424.0
    return v.clamp_(0, 4095).round().to(torch.int16)


def test_hu_windows_bit_exact(cuda_lib):
    from se_unet_airseg_b200 import _lib
    img = _synthetic_ct((17, 33, 20), 5)
    ref = oracle.two_channel(img.to(torch.float64) - 1024).to(torch.float32)
    d = img.cuda()
    out = torch.empty(2, img.numel(), device="cuda")
    _lib.check(cuda_lib.seunet_hu_windows(_lib.ptr(d), 0, img.numel(), -1024.0, _lib.ptr(out), _lib.stream_ptr()), "hu")
    assert torch.equal(out.cpu().view(2, *img.shape), ref)


@pytest.mark.parametrize("batch", [1, 5])
def test_sliding_window_matches_oracle(batch):
    from se_unet_airseg_b200 import SE_UNet
    from se_unet_airseg_b200.inference import SlidingWindowPredictor, window_starts
    assert window_starts(512) == oracle.window_starts(512) and window_starts(400) == oracle.window_starts(400)
    assert len(window_starts(512)) == 7 and len(window_starts(400)) == 6   # 294 windows, SURVEY 3.3
    sd = oracle.init_params(2, 1, seed=777)
    # shift the decoder head bias so the thresholded mask is not degenerate with random weights
    m = SE_UNet(2, 1)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    img = _synthetic_ct((40, 56, 48), 11)
    prob_ref, mask_ref = oracle.predict_volume(sd, img, cube=32, step=16)
    sw = SlidingWindowPredictor(m, cube=32, step=16, batch=batch)
    mask, prob = sw.predict_device(img.cuda(), return_prob=True)
    prob = prob.cpu().numpy().astype(np.float64)
    err = np.abs(prob - prob_ref).max()
    print("max |dprob|", err, "mask agreement", (mask.cpu().numpy().astype(bool) == mask_ref).mean())
    assert err < 5e-3   # 2e-2 logit tolerance * max sigmoid slope 0.25
    conf = np.abs(prob_ref - 0.5) > 5e-3
    assert (mask.cpu().numpy().astype(bool) == mask_ref)[conf].all()
    host_mask = sw.predict(img.numpy())
    assert host_mask.dtype == torch.uint8 and torch.equal(host_mask, mask.cpu())
