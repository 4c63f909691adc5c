"""Multi-GPU parity (needs >= 2 visible B200s; skipped otherwise):

* the reference's own multi-GPU mode, single-process `torch.nn.DataParallel` over several devices (train.py:197/396/577,
  test.py:91, prediction.py:63): forward + backward through the wrapper == the single-device result on the whole batch;
* the B200-native replacement, one process per GPU with NCCL (`DataParallelTrainer`): a 2-rank step == the 1-rank step on
  the concatenated batch (loss partial sums exchanged, gradients SUMMED - SURVEY 8e);
* patch-sharded sliding-window inference of ONE volume over 2 ranks == the 1-rank mask (prediction.py:80-110).
"""
import os
import subprocess
import sys

import pytest
import torch

from oracle import seunet_oracle as oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs2 = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")


@needs2
def test_nn_dataparallel_two_devices_matches_single_device(cuda_lib):
    from se_unet_airseg_b200 import SE_UNet
    sd = oracle.init_params(2, 1, seed=71)
    g = torch.Generator().manual_seed(8)
    x = torch.rand(4, 2, 32, 32, 32, generator=g)
    label = (torch.rand(4, 1, 32, 32, 32, generator=g) > 0.9).float()

    def run(wrap):
        m = SE_UNet(2, 1)
        m.load_state_dict(sd)
        m = m.cuda(0).eval()          # eval: DropLayer's normaliser is per replica in the reference too (SE_UNet.py:94)
        net = torch.nn.DataParallel(m, device_ids=[0, 1]) if wrap else m
        outs = []
        for it in range(2):           # twice: the second pass must reuse the per-device plans
            m.zero_grad()
            p0, p1 = net(x.cuda(0))
            assert p0.device.index == 0 and p0.shape == (4, 1, 32, 32, 32)
            loss = oracle.dice_loss(torch.sigmoid(p1), label.cuda(0)) + oracle.dice_loss(torch.sigmoid(p0), label.cuda(0))
            loss.backward()
            outs.append((p0.detach().clone(), p1.detach().clone(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}))
        return m, outs

    m_dp, o_dp = run(True)
    m_1, o_1 = run(False)
    for (a0, a1, ga), (b0, b1, gb) in zip(o_dp, o_1):
        assert torch.equal(a0, b0) and torch.equal(a1, b1)       # per-sample statistics: batch split changes nothing
        # dc62 is dead code (SE_UNet.py:230): no gradient on a single device; under multi-device DataParallel autograd's
        # Broadcast.backward materialises ZEROS for the replicas' unused tensors - in the reference exactly the same way
        assert "dc62.conv1.weight" not in gb
        if "dc62.conv1.weight" in ga:
            assert ga.pop("dc62.conv1.weight").abs().max().item() == 0.0
        assert ga.keys() == gb.keys()
        for n in ga:
            den = max(gb[n].norm().item(), 1e-12)
            # weight gradients are sums over the batch: the two replicas' partial sums are added on device 0 in another order
            assert (ga[n] - gb[n]).norm().item() <= 2e-3 * den + 1e-9, n
    # plans were cached per replica device, not rebuilt per forward
    reg = m_dp._replica_rts
    assert sorted(reg.keys()) == [0, 1]
    assert all(len(rt.plans) == 1 for rt in reg.values())


def _torchrun(script, nproc, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, script), *args]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


@needs2
def test_two_rank_nccl_trainer_step_equals_single_rank_step():
    out = _torchrun("tools/dp_check.py", 2)
    assert "DP CHECK OK" in out, out[-2000:]


@needs2
def test_two_rank_patch_sharded_inference_equals_single_rank_mask():
    out = _torchrun("tools/sharded_infer_check.py", 2)
    assert "SHARDED INFER OK" in out, out[-2000:]
