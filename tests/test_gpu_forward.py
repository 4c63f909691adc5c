"""GPU parity of the drop-in SE_UNet.forward (C ABI seunet_forward) against the fp32 CPU oracle.

Tolerances are BASELINE.json's: logits within 2e-2 max-abs of the fp32 reference, >= 99.9 % agreement of
the thresholded masks (logit >= 0 <=> sigmoid >= 0.5, prediction.py:104-111).  With random-init weights
logits sit near 0, so mask agreement is also reported on voxels with |logit_ref| > 2e-2 (SURVEY 8d hazard).
"""
import os

import numpy as np
import pytest
import torch

from oracle import seunet_oracle as oracle

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-2
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _model(in_ch, seed=777, train=False):
    from se_unet_airseg_b200 import SE_UNet
    sd = oracle.init_params(in_ch, 1, seed=seed)
    m = SE_UNet(in_ch, 1)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    m = m.cuda()
    m.train(train)
    return m, sd


def _compare(p, r, name):
    err = (p - r).abs().max().item()
    agree = ((p >= 0) == (r >= 0)).float().mean().item()
    conf = r.abs() > LOGIT_TOL
    agree_conf = ((p >= 0) == (r >= 0))[conf].float().mean().item() if conf.any() else 1.0
    print(f"{name}: max|dlogit|={err:.3e} mask agreement={agree:.5f} (confident voxels: {agree_conf:.5f})")
    assert err <= LOGIT_TOL, f"{name}: logits differ by {err}"
    assert agree_conf >= 0.999
    return err, agree


@pytest.mark.parametrize("in_ch,shape", [(2, (1, 32, 32, 32)), (1, (1, 32, 32, 32)), (2, (2, 16, 32, 48)), (2, (1, 24, 40, 24))])
def test_forward_eval_matches_oracle(in_ch, shape):
    B, D, H, W = shape
    m, sd = _model(in_ch)
    g = torch.Generator().manual_seed(42)
    x = torch.rand(B, in_ch, D, H, W, generator=g)
    with torch.no_grad():
        r0, r1 = oracle.forward(sd, x)
        p0, p1 = m(x.cuda())
    assert p0.dtype == torch.float32 and p0.shape == (B, 1, D, H, W)
    _compare(p0.cpu(), r0, "pred0")
    _compare(p1.cpu(), r1, "pred1")


def test_forward_noncontiguous_input_view():
    """prediction.py:102 feeds x[:, :, xl:xr, yl:yr, zl:zr] - a strided view - straight into the model."""
    m, sd = _model(2)
    g = torch.Generator().manual_seed(7)
    big = torch.rand(1, 2, 40, 48, 40, generator=g)
    view = big[:, :, 4:36, 8:40, 3:35]
    with torch.no_grad():
        r0, r1 = oracle.forward(sd, view.contiguous())
        bigc = big.cuda()
        p0, p1 = m(bigc[:, :, 4:36, 8:40, 3:35])
    _compare(p0.cpu(), r0, "pred0")
    _compare(p1.cpu(), r1, "pred1")


def test_forward_train_mode_droplayer_matches_oracle():
    """DropLayer is live in train mode (also in the reference's validation/test loops): same CPU-generator
    draw order (dropout1 then dropout2, SE_UNet.py:232-233) => same mask => same logits."""
    m, sd = _model(2, train=True)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 2, 16, 16, 16, generator=g)
    torch.manual_seed(2024)
    d0 = oracle.drop_scale(2, 24)
    d1 = oracle.drop_scale(2, 12)
    with torch.no_grad():
        r0, r1 = oracle.forward(sd, x, d0, d1)
        torch.manual_seed(2024)
        p0, p1 = m(x.cuda())
    _compare(p0.cpu(), r0, "pred0")
    _compare(p1.cpu(), r1, "pred1")


def test_forward_matches_golden_reference_vectors():
    """tests/golden/forward_*.npz were produced by the UNMODIFIED reference module (oracle/make_golden.py)."""
    files = sorted(f for f in os.listdir(GOLDEN) if f.startswith("forward_") and f.endswith(".npz"))
    assert files, "golden vectors missing"
    for f in files:
        z = np.load(os.path.join(GOLDEN, f))
        in_ch = int(z["in_channel"])
        m, _ = _model(in_ch, seed=int(z["seed"]))
        x = torch.from_numpy(z["x"])
        with torch.no_grad():
            p0, p1 = m(x.cuda())
        _compare(p0.cpu(), torch.from_numpy(z["pred0"]), f + ":pred0")
        _compare(p1.cpu(), torch.from_numpy(z["pred1"]), f + ":pred1")


def test_forward_full_window_128_matches_oracle_and_is_batch_consistent():
    """BASELINE config size: one 128^3 window (2,097,152 voxels, 14-42 conv tiles per CTA) against the fp32 oracle, plus
    the size-independent property that a sample's logits do not depend on what else is in the batch (InstanceNorm is per
    sample; catches cross-sample leaks in the persistent multi-tile kernels)."""
    m, sd = _model(2)
    g = torch.Generator().manual_seed(128)
    x = torch.rand(2, 2, 128, 128, 128, generator=g)
    with torch.no_grad():
        r0, r1 = oracle.forward(sd, x[:1])
        xb = x.cuda()
        p0, p1 = m(xb)
        q0, q1 = m(xb[1:2])
    e0, a0 = _compare(p0[:1].cpu(), r0, "128^3 pred0")
    e1, a1 = _compare(p1[:1].cpu(), r1, "128^3 pred1")
    # Raw mask agreement with RANDOM-INIT weights is a worst case (SURVEY 8d hazard): the logits are N(~0, 0.04), so ~0.1 %
    # of 2M voxels lie within the 1e-3 logit error of the threshold.  Every voxel with |logit_ref| > 2e-2 agrees (checked
    # in _compare); the raw figure is printed above and held to 99.8 % here (it is >= 99.9 % on all smaller cases).
    assert a1 >= 0.998, f"raw thresholded-mask agreement {a1:.5f}"
    # same kernels, same data, different batch position: only the fp64-atomic summation order may differ
    assert (p1[1:2] - q1).abs().max().item() <= 1e-4 and (p0[1:2] - q0).abs().max().item() <= 1e-4


def test_forward_config5_window_160_matches_oracle():
    """BASELINE config 5 patch size (160^3: 20 x 10 in-plane tiles, 20^3 coarsest level - not a power of two anywhere)."""
    m, sd = _model(2)
    g = torch.Generator().manual_seed(160)
    x = torch.rand(1, 2, 160, 160, 160, generator=g)
    with torch.no_grad():
        r0, r1 = oracle.forward(sd, x)
        p0, p1 = m(x.cuda())
    _compare(p0.cpu(), r0, "160^3 pred0")
    _, a1 = _compare(p1.cpu(), r1, "160^3 pred1")
    assert a1 >= 0.998   # random-init worst case, see the 128^3 test above


def test_mask_agreement_with_briefly_trained_weights_at_full_window():
    """SURVEY 8d hazard, option (i): with random-init weights the logits sit at 0 and the 99.9 % mask bar measures rounding
    noise (see the 128^3 test above).  Here the network is trained for a few dozen steps on synthetic tubes with the B200
    trainer (forward + dice loss + backward + fused AdamW, all through the C ABI), and the thresholded masks of the CUDA forward and
    of the fp32 CPU oracle are compared on a full 128^3 window with THOSE weights: >= 99.9 % raw agreement, logits within 2e-2."""
    from se_unet_airseg_b200 import SE_UNet
    from se_unet_airseg_b200.trainer import DataParallelTrainer
    sd = oracle.init_params(2, 1, seed=2024)
    m = SE_UNet(2, 1)
    m.load_state_dict(sd)
    m = m.cuda().train()
    tr = DataParallelTrainer(m, stage=1, lr=2e-3)

    def sample(B, S, seed):
        g = torch.Generator().manual_seed(seed)
        label = torch.zeros(B, 1, S, S, S)
        for b in range(B):                                # a few axis-aligned and diagonal "airways"
            for _ in range(6):
                c = torch.randint(4, S - 4, (3,), generator=g)
                r = int(torch.randint(1, 3, (1,), generator=g))
                ax = int(torch.randint(0, 3, (1,), generator=g))
                sl = [slice(int(c[0]) - r, int(c[0]) + r + 1), slice(int(c[1]) - r, int(c[1]) + r + 1), slice(int(c[2]) - r, int(c[2]) + r + 1)]
                sl[ax] = slice(2, S - 2)
                label[b, 0][tuple(sl)] = 1.0
        hu = torch.where(label > 0, torch.full_like(label, -950.0), torch.full_like(label, -600.0))
        hu = hu + 60.0 * torch.randn(hu.shape, generator=g)
        x = oracle.two_channel(hu[:, 0].double()).transpose(0, 1).float()      # (B, 2, S, S, S)
        return x.contiguous(), label

    for it in range(40):
        x, label = sample(2, 64, 100 + it)
        loss = tr.step(x.cuda(), label.cuda())
    print("training loss after 40 steps:", float(loss))
    m.eval()
    x, label = sample(1, 128, 999)
    trained = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        r0, r1 = oracle.forward(trained, x)
        p0, p1 = m(x.cuda())
    e1, a1 = _compare(p1.cpu(), r1, "trained 128^3 pred1")
    e0, a0 = _compare(p0.cpu(), r0, "trained 128^3 pred0")
    fg = (r1 >= 0).float().mean().item()
    print(f"foreground fraction of the reference mask: {fg:.4f}")
    assert 0.0 < fg < 0.5, "training did not produce a non-degenerate mask"
    assert a1 >= 0.999 and a0 >= 0.999


def test_fused_cat_pass_matches_oracle(monkeypatch):
    """Inference plans compute the CAT 1x1x1 convs inside the apply pass of the block that completes their concat
    (csrc/pointwise3.cu, warp-level mma.sync); SEUNET_CAT_FUSION=0 keeps the tcgen05 1x1x1 conv.  Both must give the network."""
    from se_unet_airseg_b200 import SE_UNet
    sd = oracle.init_params(2, 1, seed=777)
    x = torch.rand(2, 2, 32, 40, 48, generator=torch.Generator().manual_seed(3))      # a shape no other test plans for
    with torch.no_grad():
        r0, r1 = oracle.forward(sd, x)
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("SEUNET_CAT_FUSION", flag)      # read when the plan is created
        m = SE_UNet(2, 1)
        m.load_state_dict(sd)
        m = m.cuda().eval()
        with torch.no_grad():
            p0, p1 = m(x.cuda())
        assert (p0.cpu() - r0).abs().max().item() <= 2e-2 and (p1.cpu() - r1).abs().max().item() <= 2e-2
        outs.append(p1)
    assert (outs[0] - outs[1]).abs().max().item() <= 2e-3       # same storage points, other accumulation order
