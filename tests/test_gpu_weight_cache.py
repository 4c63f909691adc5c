"""Regression tests for the weight caches between the host module, the trainer and the packed tensor-core images:
a train / validate / train / validate loop (train.py:403-512: `validation()` after every epoch) must validate against the
CURRENT weights, `load_state_dict` after the trainer was built must reach the kernels, and the sliding-window predictor
must not hand out aliases of one result buffer."""
import numpy as np
import pytest
import torch

from oracle import seunet_oracle as oracle

pytestmark = pytest.mark.gpu


def _batch(seed, B=2, S=16):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 2, S, S, S, generator=g).cuda()
    label = (torch.rand(B, 1, S, S, S, generator=g) > 0.9).float().cuda()
    weight = torch.where(label > 0, torch.full_like(label, 1.7), torch.ones_like(label))
    return x, label, weight


def _fresh_eval(state_dict, x):
    from se_unet_airseg_b200 import SE_UNet
    m = SE_UNet(2, 1)
    m.load_state_dict({k: v.detach().cpu().clone() for k, v in state_dict.items()})
    m = m.cuda().eval()
    with torch.no_grad():
        return m(x)


def test_train_eval_train_eval_sees_current_weights(cuda_lib):
    from se_unet_airseg_b200 import SE_UNet
    from se_unet_airseg_b200.trainer import DataParallelTrainer
    m = SE_UNet(2, 1)
    m.load_state_dict(oracle.init_params(2, 1, seed=41))
    m = m.cuda()
    tr = DataParallelTrainer(m, stage=2, lr=1e-2)      # large steps: stale weights would be far outside the tolerance
    x, label, weight = _batch(1)
    xv = torch.rand(1, 2, 16, 16, 16, generator=torch.Generator().manual_seed(2)).cuda()
    prev = None
    for rnd in range(3):
        m.train()
        for _ in range(2):
            tr.step(x, label, weight)
        m.eval()
        with torch.no_grad():
            got = m(xv)
        want = _fresh_eval(m.state_dict(), xv)
        for a, b in zip(got, want):
            assert torch.equal(a, b), f"round {rnd}: eval forward used stale weights (max diff {(a - b).abs().max().item():.3e})"
        if prev is not None:
            assert not torch.equal(prev, got[1]), "weights did not change between rounds"
        prev = got[1].clone()


def test_load_state_dict_after_trainer_construction_reaches_the_kernels(cuda_lib):
    from se_unet_airseg_b200 import SE_UNet
    from se_unet_airseg_b200.trainer import DataParallelTrainer
    sd_a, sd_b = oracle.init_params(2, 1, seed=51), oracle.init_params(2, 1, seed=52)
    x, label, weight = _batch(3)

    def first_step_grads(build):
        m = SE_UNet(2, 1)
        m.load_state_dict(sd_a)
        m = m.cuda().eval()            # eval: no DropLayer draw, both runs see the same graph
        tr = DataParallelTrainer(m, stage=2)
        build(m, tr)
        loss = tr.step(x, label, weight).item()
        return loss, tr.grads.clone()

    def late_load(m, tr):
        tr.step(x, label, weight)      # packs the images of sd_a, then the checkpoint arrives (train.py:393-395 idiom)
        m.load_state_dict(sd_b)
        tr.m.zero_(); tr.v.zero_(); tr.step_count = 0

    def early_load(m, tr):
        m.load_state_dict(sd_b)

    la, ga = first_step_grads(late_load)
    lb, gb = first_step_grads(early_load)
    assert abs(la - lb) <= 1e-6 * max(1.0, abs(lb)), (la, lb)
    assert (ga - gb).norm().item() <= 1e-5 * gb.norm().item()


def test_trainer_master_buffer_is_shared_with_the_module(cuda_lib):
    """The module's eval path reads the trainer's flat master copy itself (no second 6 MB gather per validation call)."""
    from se_unet_airseg_b200 import SE_UNet
    from se_unet_airseg_b200.trainer import DataParallelTrainer
    m = SE_UNet(2, 1).cuda()
    tr = DataParallelTrainer(m, stage=1)
    flat, gen0 = m._weights()
    assert flat is tr.flat
    x, label, _ = _batch(4)
    tr.step(x, label)
    flat2, gen1 = m._weights()
    assert flat2 is tr.flat and gen1 > gen0
    # moving the module re-creates the parameters: the trainer adopts the new tensors instead of updating orphans
    m.to("cuda")
    m.ec1.conv1.weight.data = m.ec1.conv1.weight.data.clone()
    tr.step(x, label)
    assert m._weights()[0] is tr.flat
    assert m.ec1.conv1.weight.data_ptr() == tr.flat.data_ptr()


def test_predictor_results_are_not_aliases(cuda_lib):
    from se_unet_airseg_b200 import SE_UNet
    from se_unet_airseg_b200.inference import SlidingWindowPredictor
    m = SE_UNet(2, 1)
    m.load_state_dict(oracle.init_params(2, 1, seed=61))
    m = m.cuda().eval()
    sw = SlidingWindowPredictor(m, cube=32, step=16, batch=4, streams=2)
    rng = np.random.RandomState(0)
    vols = [rng.randint(0, 2400, size=(48, 32, 40)).astype(np.int16) for _ in range(2)]
    masks = [sw.predict(v) for v in vols]
    assert masks[0].data_ptr() != masks[1].data_ptr()
    again = sw.predict(vols[0])
    assert torch.equal(masks[0], again), "first result was overwritten by the second call"
    dmasks = [sw.predict_device(torch.from_numpy(v).cuda()) for v in vols]
    assert dmasks[0].data_ptr() != dmasks[1].data_ptr()
    assert torch.equal(dmasks[0].cpu(), masks[0]) and torch.equal(dmasks[1].cpu(), masks[1])
    # the zero-copy mode is explicit
    r1 = sw.predict(vols[0], reuse_output=True)
    r2 = sw.predict(vols[1], reuse_output=True)
    assert r1.data_ptr() == r2.data_ptr()
