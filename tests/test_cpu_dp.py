"""CPU world_size-2 (gloo) test of the data-parallel step's HOST logic: exchanging the loss PARTIAL SUMS and
SUMMING gradients across ranks reproduces the single-process step on the concatenated batch.  The per-rank math is done
by the fp32 oracle here (no GPU on this box); the exchange pattern is the one se_unet_airseg_b200/trainer.py uses."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import seunet_oracle as oracle
from se_unet_airseg_b200.trainer import loss_from_sums


def _sums(stage, pe, pd, label, weight, skel):
    out = torch.zeros(2, 8, dtype=torch.float64)
    for h, z in enumerate((pe, pd)):
        p = torch.sigmoid(z).double()
        t, w, s = label.double(), weight.double(), skel.double()
        out[h, 0], out[h, 1], out[h, 2] = (p * t).sum(), p.sum(), t.sum()
        out[h, 3], out[h, 4] = (w * (p + 1e-4) ** 0.7 * t).sum(), (w * (0.2 * p + 0.8 * t)).sum()
        out[h, 5], out[h, 6] = (w * p * s * s).sum(), (w * (p * s + s)).sum()
    return out


def _loss_from_tensor_sums(stage, s):
    e, d = s[0], s[1]
    dice = lambda v: 1.0 - (2.0 * v[0] + 1.0) / (v[1] + v[2] + 1.0)
    gul = lambda v: 1.0 - (v[3] + 1.0) / (v[4] + 1.0)
    atr = lambda v: 1.0 - (v[5] + 1.0) / (v[6] + 1.0)
    if stage == 1:
        return dice(d) + dice(e)
    loss = gul(d) + 0.5 * gul(e)
    return loss + 0.5 * (atr(e) + atr(d)) if stage == 3 else loss


def _data(B):
    g = torch.Generator().manual_seed(123)
    x = torch.rand(B, 2, 16, 16, 16, generator=g)
    label = (torch.rand(B, 1, 16, 16, 16, generator=g) > 0.8).float()
    weight = torch.where(label > 0, torch.rand(B, 1, 16, 16, 16, generator=g) + 0.5, torch.ones(B, 1, 16, 16, 16))
    skel = label * (torch.rand(B, 1, 16, 16, 16, generator=g) > 0.5).float()
    return x, label, weight, skel


def _worker(rank, world, port, stage, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    x, label, weight, skel = (t.chunk(world)[rank] for t in _data(4))
    sd = {k: v.requires_grad_(True) for k, v in oracle.init_params(2, 1, seed=9).items()}
    pe, pd = oracle.forward(sd, x)
    local = _sums(stage, pe.detach(), pd.detach(), label, weight, skel)
    glob = local.clone()
    dist.all_reduce(glob, op=dist.ReduceOp.SUM)                      # C4: partial sums, not per-rank losses
    # loss as a function of this rank's logits with the OTHER ranks' contributions held constant
    mine = _sums_graph(stage, pe, pd, label, weight, skel)
    loss = _loss_from_tensor_sums(stage, mine + (glob - local))
    loss.backward()
    flat = torch.cat([(v.grad if v.grad is not None else torch.zeros_like(v)).flatten() for v in sd.values()])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                      # C3: SUM, not mean
    if rank == 0:
        torch.save({"loss": loss.item(), "grad": flat, "host_loss": loss_from_sums(stage, glob.tolist())}, out)
    dist.destroy_process_group()


def _sums_graph(stage, pe, pd, label, weight, skel):
    rows = []
    for z in (pe, pd):
        p = torch.sigmoid(z).double()
        t, w, s = label.double(), weight.double(), skel.double()
        rows.append(torch.stack([(p * t).sum(), p.sum(), t.sum(), (w * (p + 1e-4) ** 0.7 * t).sum(), (w * (0.2 * p + 0.8 * t)).sum(),
                                 (w * p * s * s).sum(), (w * (p * s + s)).sum(), torch.zeros((), dtype=torch.float64)]))
    return torch.stack(rows)


def test_two_rank_dp_step_equals_single_rank_on_concatenated_batch(tmp_path):
    stage = 3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, port, stage, out), nprocs=2, join=True)
    got = torch.load(out)
    x, label, weight, skel = _data(4)
    sd = {k: v.requires_grad_(True) for k, v in oracle.init_params(2, 1, seed=9).items()}
    pe, pd = oracle.forward(sd, x)
    loss = oracle.stage_loss(stage, pe, pd, label, weight, skel)
    loss.backward()
    flat = torch.cat([(v.grad if v.grad is not None else torch.zeros_like(v)).flatten() for v in sd.values()])
    assert abs(got["loss"] - loss.item()) <= 1e-6 and abs(got["host_loss"] - loss.item()) <= 1e-6
    assert (got["grad"] - flat).norm().item() <= 1e-4 * flat.norm().item()
