"""GPU parity of the fused loss kernels (train.py:51-76 + stage combinations), the fused AdamW step and the
single-rank DataParallelTrainer step against torch autograd / torch.optim.AdamW on the same tensors."""
import pytest
import torch

from oracle import seunet_oracle as oracle

pytestmark = pytest.mark.gpu


def _targets(shape, seed):
    g = torch.Generator().manual_seed(seed)
    label = (torch.rand(shape, generator=g) > 0.9).float()
    weight = torch.where(label > 0, torch.rand(shape, generator=g) * 2.0 + 0.5, torch.ones(shape))
    skel = label * (torch.rand(shape, generator=g) > 0.5).float()
    return label, weight, skel


@pytest.mark.parametrize("stage", [1, 2, 3])
def test_loss_kernels_match_oracle_autograd(cuda_lib, stage):
    from se_unet_airseg_b200 import _lib
    from se_unet_airseg_b200.trainer import loss_from_sums
    L, p_ = cuda_lib, _lib.ptr
    shape = (3, 1, 16, 24, 16)
    g = torch.Generator().manual_seed(stage)
    pe = (torch.randn(shape, generator=g) * 2).requires_grad_(True)
    pd = (torch.randn(shape, generator=g) * 2).requires_grad_(True)
    label, weight, skel = _targets(shape, 10 + stage)
    ref = oracle.stage_loss(stage, pe.double(), pd.double(), label.double(), weight.double(), skel.double())
    ge_ref, gd_ref = torch.autograd.grad(ref, (pe, pd))
    dev = "cuda"
    t = [v.detach().to(dev).contiguous() for v in (pe, pd, label, weight, skel)]
    sums = torch.zeros(16, dtype=torch.float64, device=dev)
    per = torch.zeros(shape[0] * 4, dtype=torch.float64, device=dev)
    ge, gd = torch.empty_like(t[0]), torch.empty_like(t[0])
    loss = torch.zeros(1, device=dev)
    st = _lib.stream_ptr()
    V = shape[2] * shape[3] * shape[4]
    _lib.check(L.seunet_loss_sums(stage, p_(t[0]), p_(t[1]), p_(t[2]), p_(t[3]), p_(t[4]), shape[0], V, p_(sums), p_(per), st), "sums")
    _lib.check(L.seunet_loss_grad(stage, p_(t[0]), p_(t[1]), p_(t[2]), p_(t[3]), p_(t[4]), shape[0] * V, p_(sums), p_(ge), p_(gd),
                                  p_(loss), st), "grad")
    torch.cuda.synchronize()
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    assert abs(loss_from_sums(stage, sums.cpu().view(2, 8).tolist()) - ref.item()) <= 1e-5
    for a, r in ((ge, ge_ref), (gd, gd_ref)):
        assert (a.cpu() - r).norm().item() <= 1e-4 * r.norm().item()
    if stage >= 2:   # per-sample GUL of the decoder head, as used for hard-mining file names (train.py:249-253)
        ps = per.cpu().view(shape[0], 2, 2)
        for b in range(shape[0]):
            want = oracle.general_union_loss_lib(torch.sigmoid(pd[b:b + 1]).double(), label[b:b + 1].double(), weight[b:b + 1].double())
            got = 1 - (ps[b, 1, 0] + 1) / (ps[b, 1, 1] + 1)
            assert abs(got.item() - want.item()) <= 1e-5


def test_adamw_kernel_matches_torch_adamw(cuda_lib):
    from se_unet_airseg_b200 import _lib
    L, p_ = cuda_lib, _lib.ptr
    g = torch.Generator().manual_seed(0)
    n = 10007
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-4)
    p = p0.clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        gr = torch.randn(n, generator=g) * 10 ** (-step)
        ref.grad = gr.clone()
        opt.step()
        _lib.check(L.seunet_adamw_step(p_(p), p_(gr.cuda()), p_(m), p_(v), n, 1e-4, 0.9, 0.999, 1e-8, 1e-2, step, 1.0, 100, 50,
                                       _lib.stream_ptr()), "adamw")
    torch.cuda.synchronize()
    out = p.cpu()
    keep = torch.ones(n, dtype=torch.bool)
    keep[100:150] = False
    assert torch.allclose(out[keep], ref.detach()[keep], rtol=0, atol=2e-7)
    assert torch.equal(out[~keep], p0[~keep])       # skipped range (dc62) is never touched


@pytest.mark.parametrize("stage", [1, 3])
def test_trainer_step_equals_autograd_plus_torch_adamw(stage):
    """Fused step (C-ABI forward/loss/backward/AdamW on flat buffers) == nn.Module autograd path + torch.optim.AdamW."""
    from se_unet_airseg_b200 import SE_UNet
    from se_unet_airseg_b200.trainer import DataParallelTrainer
    sd = oracle.init_params(2, 1, seed=31)
    shape = (2, 1, 16, 16, 16)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 2, 16, 16, 16, generator=g).cuda()
    label, weight, skel = (t.cuda() for t in _targets(shape, 6))
    ma, mb = SE_UNet(2, 1), SE_UNet(2, 1)
    ma.load_state_dict(sd); mb.load_state_dict(sd)
    ma, mb = ma.cuda().train(), mb.cuda().train()
    opt = torch.optim.AdamW(ma.parameters(), lr=1e-4)
    tr = DataParallelTrainer(mb, stage=stage)
    for it in range(2):
        torch.manual_seed(100 + it)
        pe, pd = ma(x)
        loss_a = oracle.stage_loss(stage, pe, pd, label, weight, skel)
        opt.zero_grad()
        loss_a.backward()
        torch.manual_seed(100 + it)
        loss_b = tr.step(x, label, weight, skel)
        assert abs(loss_a.item() - loss_b.item()) <= 1e-5
        if it == 0:      # same kernels behind both paths (dpred differs at 1e-7: torch vs fused loss gradient)
            off = 0
            for n, p in ma.named_parameters():
                k = p.numel()
                gb = tr.grads[off:off + k].view(p.shape)
                off += k
                if p.grad is None:
                    assert n == "dc62.conv1.weight"
                    continue
                assert (p.grad - gb).norm().item() <= 1e-2 * max(gb.norm().item(), 1e-12), n   # fp16 dY rounding decorrelates at ~1e-3
        opt.step()
    # Adam turns any non-zero gradient into a step of ~lr, so entries whose gradient is at the fp32 noise level (1e-8) move
    # by up to lr in either path; everything else must coincide.
    diffs = torch.cat([(a - b).abs().flatten() for (_, a), (_, b) in zip(ma.named_parameters(), mb.named_parameters())])
    assert diffs.max().item() <= 4.1e-4            # <= 2 steps * 2 * lr
    assert diffs.mean().item() <= 5e-6
    assert (diffs > 2e-5).float().mean().item() <= 0.05
    assert torch.equal(mb.dc62.conv1.weight.cpu(), sd["dc62.conv1.weight"])


def test_reference_training_loop_runs_unchanged_on_the_drop_in_module():
    """The reference's own training idiom (train.py:186-188, 323, 597-603): nn.DataParallel wrapper, torch.optim.AdamW on
    model.parameters(), torch loss on the two outputs, loss.backward(), optimizer.step(), model.module.state_dict().
    With one visible GPU DataParallel.forward short-circuits to the module."""
    from se_unet_airseg_b200 import SE_UNet
    torch.manual_seed(5)
    model = SE_UNet(in_channel=2, n_classes=1).cuda()
    model = torch.nn.DataParallel(model, device_ids=[0])
    optimizer = torch.optim.AdamW(model.parameters(), lr=1e-3)
    model.train()
    g = torch.Generator().manual_seed(9)
    x = torch.rand(2, 2, 32, 32, 32, generator=g).cuda()
    label = torch.zeros(2, 1, 32, 32, 32)
    label[:, :, 8:24, 14:18, 14:18] = 1.0
    label = label.cuda()
    x[:, :, 8:24, 14:18, 14:18] *= 0.2          # the "airway" is darker

    def dice_loss(pred, target):                 # train.py:51-57
        smooth = 1.0
        iflat, tflat = pred.view(-1), target.view(-1)
        inter = (iflat * tflat).sum()
        return 1 - (2.0 * inter + smooth) / (iflat.sum() + tflat.sum() + smooth)

    before = {k: v.detach().clone() for k, v in model.module.state_dict().items()}
    losses = []
    for _ in range(12):
        p_en, p_de = model(x)
        loss = dice_loss(torch.sigmoid(p_de), label) + dice_loss(torch.sigmoid(p_en), label)   # train.py:597-599
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        losses.append(loss.item())
    print("reference-style loop losses:", [round(v, 4) for v in losses])
    assert losses[-1] < losses[0] - 0.02, "loss did not decrease"
    after = model.module.state_dict()
    assert len(after) == 117 and list(after.keys()) == list(before.keys())
    changed = [k for k in after if not torch.equal(after[k], before[k])]
    frozen = [k for k in after if k not in changed]
    assert all(k.startswith("dc62") for k in frozen), f"parameters without updates: {frozen[:5]}"   # dc62 is dead code in the graph
    # the saved checkpoint loads back into a fresh module exactly like train.py/prediction.py do
    fresh = SE_UNet(in_channel=2, n_classes=1)
    fresh.load_state_dict({k: v.cpu() for k, v in after.items()}, strict=False)
    fresh = fresh.cuda().eval()
    model.eval()
    with torch.no_grad():
        a0, a1 = model(x)
        b0, b1 = fresh(x)
    assert torch.equal(a1, b1) and torch.equal(a0, b0)


def test_graph_mode_step_equals_eager_step():
    """DataParallelTrainer(graph=True) replays forward + losses + backward as one CUDA graph: same kernels on static copies of
    the inputs, so the trajectory equals the eager one up to the order of the floating-point atomics; new inputs, new DropLayer
    draws and the changing AdamW step number must all reach the replayed step (train.py:428-440 feeds a new batch every step)."""
    from se_unet_airseg_b200 import SE_UNet
    from se_unet_airseg_b200.trainer import DataParallelTrainer
    sd = oracle.init_params(2, 1, seed=41)
    shape = (2, 1, 16, 24, 16)
    batches = []
    for it in range(5):
        g = torch.Generator().manual_seed(50 + it)
        x = torch.rand(2, 2, 16, 24, 16, generator=g).cuda()
        batches.append((x,) + tuple(t.cuda() for t in _targets(shape, 60 + it)))
    runs = []
    for graph in (False, True):
        m = SE_UNet(2, 1)
        m.load_state_dict(sd)
        m = m.cuda().train()
        tr = DataParallelTrainer(m, stage=3, graph=graph)
        losses = []
        for it, b in enumerate(batches):
            torch.manual_seed(200 + it)           # the DropLayer draws come from the CPU generator
            losses.append(tr.step(*b).item())
        runs.append((losses, tr.flat.clone(), tr))
    (la, pa, _), (lb, pb, trb) = runs
    assert trb._graphs and all(g["graph"] is not None for g in trb._graphs.values())    # steps 3.. were replays
    assert all(abs(a - b) <= 2e-5 for a, b in zip(la, lb)), (la, lb)
    assert len(set(round(v, 6) for v in lb)) == len(lb)                               # every step saw its own batch
    d = (pa - pb).abs()
    assert d.max().item() <= 5 * 2 * 1e-4 + 1e-6 and d.mean().item() <= 1e-5           # see the AdamW remark in the test above
