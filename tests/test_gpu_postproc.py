"""GPU post-processing (SURVEY 8f N4, C ABI seunet_postproc_*) against the numpy oracle and the reference's own outputs.

Bit-exact: the hysteresis sweep (prediction.py:13-37, including its raster-order dependence), the border crop
(prediction.py:112-115) and maximum_3d (util.py:58-75: largest 26-connected component, probe-slice fallback, hole filling).
"""
import os

import numpy as np
import pytest
import torch

from oracle import seunet_oracle as oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _prob(shape, seed):
    from scipy import ndimage
    rng = np.random.RandomState(seed)
    f = ndimage.gaussian_filter(rng.rand(*shape), 1.2)
    f = (f - f.min()) / (f.max() - f.min())
    p = (0.15 + 0.7 * f + 0.08 * (rng.rand(*shape) - 0.5)).astype(np.float32)
    flat = p.reshape(-1)
    flat[rng.randint(0, flat.size, 8)] = np.float32(0.5)   # exact threshold values
    flat[rng.randint(0, flat.size, 8)] = np.float32(0.4)
    return p


def _pp(shape, **kw):
    from se_unet_airseg_b200.postprocess import PostProcessor
    return PostProcessor(shape, "cuda:0", **kw)


def test_dti_matches_reference_golden_vectors():
    """tests/golden/postproc_dti_*.npz hold outputs of the reference's OWN double_threshold_iteration (oracle/make_golden.py)."""
    files = sorted(f for f in os.listdir(GOLDEN) if f.startswith("postproc_dti_"))
    assert files
    for f in files:
        z = np.load(os.path.join(GOLDEN, f))
        p = torch.from_numpy(z["prob"]).cuda()
        got = _pp(p.shape).dti(p).cpu().numpy()
        assert np.array_equal(got, z["dti"]), f


@pytest.mark.parametrize("shape,seed", [((8, 8, 8), 0), ((10, 12, 31), 1), ((7, 9, 33), 2), ((12, 6, 64), 3), ((6, 11, 100), 4), ((20, 24, 40), 5)])
def test_dti_matches_oracle(shape, seed):
    p = _prob(shape, seed)
    want = oracle.double_threshold_iteration(p, 0.5, 0.4)
    got = _pp(shape).dti(torch.from_numpy(p).cuda()).cpu().numpy()
    assert np.array_equal(got, want.astype(np.uint8))
    # the sweep only ever adds weak voxels to the strong set
    assert np.all(got >= (p >= 0.5)) and np.all(got <= (p >= 0.4))


def test_dti_is_order_dependent_like_the_reference():
    """A weak chain running AGAINST the raster order is not followed (single sweep): only the voxel next to the strong seed
    is set; a chain running WITH the raster order is followed to its end."""
    p = np.full((3, 3, 12), 0.1, np.float32)
    p[1, 1, :] = 0.45          # weak line along k
    p[1, 1, 6] = 0.9           # strong seed in the middle
    got = _pp(p.shape).dti(torch.from_numpy(p).cuda()).cpu().numpy()
    want = oracle.double_threshold_iteration(p, 0.5, 0.4)
    assert np.array_equal(got, want.astype(np.uint8))
    assert got[1, 1, 7:].all() and got[1, 1, 5] == 1 and not got[1, 1, :5].any()


def test_border_crop_matches_reference_rule():
    for shape in ((20, 20, 8), (13, 27, 9), (40, 33, 16)):
        p = np.full(shape, 0.9, np.float32)
        got = _pp(shape).dti(torch.from_numpy(p).cuda(), border_frac=0.15).cpu().numpy()
        want = oracle.zero_borders(np.ones(shape), 0.15)
        assert np.array_equal(got, want.astype(np.uint8)), shape


def _blobs(shape, seed, thr=0.55):
    from scipy import ndimage
    rng = np.random.RandomState(seed)
    f = ndimage.gaussian_filter(rng.rand(*shape), 1.5)
    f = (f - f.min()) / (f.max() - f.min())
    return (f > thr).astype(np.uint8)


@pytest.mark.parametrize("shape,seed,thr", [((16, 16, 16), 0, 0.55), ((12, 20, 45), 1, 0.6), ((24, 18, 70), 2, 0.5), ((9, 9, 130), 3, 0.58),
                                            ((32, 32, 32), 4, 0.62)])
def test_largest_component_matches_oracle(shape, seed, thr):
    m = _blobs(shape, seed, thr)
    pp = _pp(shape)
    for fill in (False, True):
        got = pp.largest_component(torch.from_numpy(m).cuda(), fill_holes=fill).cpu().numpy()
        want = oracle.maximum_3d(m, fill_holes=fill)
        assert np.array_equal(got.astype(bool), want), (shape, fill)


def test_largest_component_matches_reference_golden_vectors():
    """tests/golden/postproc_max3d.npz hold outputs of the reference's OWN maximum_3d (util.py:58-75; oracle/make_golden.py):
    area tie, probe-slice fallback, corner connectivity, hole filling, random blobs."""
    z = np.load(os.path.join(GOLDEN, "postproc_max3d.npz"))
    names = sorted(k[3:] for k in z.files if k.startswith("in."))
    assert len(names) >= 7
    for n in names:
        vol = z["in." + n].astype(np.uint8)
        pp = _pp(vol.shape)
        got = pp.largest_component(torch.from_numpy(vol).cuda()).cpu().numpy()
        assert np.array_equal(got, z["out." + n]), n
        if n == "probe":
            assert pp.last_info["used_second"]
        if n == "tie":
            assert pp.last_info["largest"] == pp.last_info["second"] == 12 and not pp.last_info["used_second"]


def test_largest_component_diagonal_connectivity_and_holes():
    m = np.zeros((10, 10, 40), np.uint8)
    for t in range(8):                      # a 26-connected diagonal staircase: one component
        m[1 + t, 1 + t, 16 + t] = 1
    m[2:9, 2:9, 2:9] = 1                    # a bigger hollow box elsewhere ...
    m[3:8, 3:8, 3:8] = 0                    # ... with a sealed cavity
    m[5, 5, 5] = 1                          # and a speck inside the cavity (its own component)
    pp = _pp(m.shape)
    got = pp.largest_component(torch.from_numpy(m).cuda()).cpu().numpy()
    want = oracle.maximum_3d(m)
    assert np.array_equal(got.astype(bool), want)
    # the box does not touch the probe slices k = 20, 13, 26 but the staircase does -> reference falls back to the SECOND largest
    assert pp.last_info["used_second"] and got[1, 1, 16] == 1 and got[4, 4, 4] == 0


def test_empty_and_single_component():
    shape = (8, 8, 40)
    pp = _pp(shape)
    assert not pp.largest_component(torch.zeros(shape, dtype=torch.uint8, device="cuda")).any()
    m = np.zeros(shape, np.uint8); m[2:5, 2:5, 18:23] = 1
    got = pp.largest_component(torch.from_numpy(m).cuda()).cpu().numpy()
    assert np.array_equal(got, m)


def test_run_table_overflow_is_reported():
    from se_unet_airseg_b200 import _lib
    shape = (8, 8, 64)
    m = np.zeros(shape, np.uint8); m[:, :, ::2] = 1    # 32 runs per row
    pp = _pp(shape, max_runs=100)
    with pytest.raises(_lib.SeunetError):
        pp.largest_component(torch.from_numpy(m).cuda())


def test_pipeline_matches_oracle_composition():
    """prediction.py:111-116 end to end on a moderate volume."""
    shape = (40, 44, 70)
    p = _prob(shape, 11)
    want = oracle.maximum_3d(oracle.zero_borders(oracle.double_threshold_iteration(p, 0.5, 0.4), 0.15))
    got = _pp(shape)(torch.from_numpy(p).cuda()).cpu().numpy()
    assert np.array_equal(got.astype(bool), want)


def test_full_size_volume_properties_and_scipy_components():
    """BASELINE config-2 size (512x512x400).  The pure-Python sweep cannot run here, so the sweep is checked through
    size-independent properties (strong <= result <= strong|weak; idempotent under a second component pass) and the
    component stage against scipy on the same mask."""
    from scipy import ndimage
    D, H, W = 512, 512, 400
    g = torch.Generator(device="cuda").manual_seed(5)
    coarse = torch.rand(1, 1, 32, 32, 25, generator=g, device="cuda")
    p = torch.nn.functional.interpolate(coarse, size=(D, H, W), mode="trilinear", align_corners=True)[0, 0].contiguous()
    p = (0.1 + 0.75 * p).float()
    pp = _pp((D, H, W))
    dti = pp.dti(p, border_frac=0.15)
    strong, weak = p >= 0.5, p >= 0.4
    inner = torch.zeros_like(strong); inner[int(0.15 * D):int(0.85 * D), int(0.15 * H):int(0.85 * H)] = True
    assert bool(((dti > 0) <= (weak & inner)).all()) and bool(((strong & inner) <= (dti > 0)).all())
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); pp.dti(p, border_frac=0.15, want_mask=False); e1.record()
    out = pp.largest_component(None)
    e2.record(); torch.cuda.synchronize()
    print(f"512x512x400 post-processing on the GPU: hysteresis sweep {e0.elapsed_time(e1):.1f} ms, components + holes {e1.elapsed_time(e2):.1f} ms")
    info = dict(pp.last_info)
    dti_np = dti.cpu().numpy()
    want = oracle.maximum_3d(dti_np)
    assert np.array_equal(out.cpu().numpy().astype(bool), want)
    label, num = ndimage.label(dti_np, structure=np.ones((3, 3, 3)))
    assert info["largest"] == int(np.bincount(label.ravel())[1:].max())
    again = pp.largest_component(out)
    assert torch.equal(again, out)


def test_predictor_postprocessed_matches_oracle_chain():
    """prediction.py:78-116 in one device call vs the oracle's sliding-window restatement followed by the numpy post-processing."""
    from se_unet_airseg_b200 import SE_UNet
    from se_unet_airseg_b200.inference import SlidingWindowPredictor
    sd = oracle.init_params(2, 1, seed=777)
    m = SE_UNet(2, 1); m.load_state_dict(sd); m = m.cuda().eval()
    rng = np.random.RandomState(3)
    img = np.clip(rng.normal(424, 400, (40, 48, 40)), 0, 4095).astype(np.int16)
    sw = SlidingWindowPredictor(m, cube=32, step=16, batch=3, streams=1)
    mask_dev, prob_dev = sw.predict_device(torch.from_numpy(img).cuda(), return_prob=True)
    prob = prob_dev.cpu().numpy().copy()
    got = sw.predict_postprocessed_device(torch.from_numpy(img).cuda()).cpu().numpy()
    # the post-processing is bit-exact given the probability volume (the forward itself is covered by test_gpu_sliding_window.py)
    want = oracle.maximum_3d(oracle.zero_borders(oracle.double_threshold_iteration(prob, 0.5, 0.4), 0.15))
    assert np.array_equal(got.astype(bool), want)
