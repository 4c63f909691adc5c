"""GPU parity of the backward pass (C ABI seunet_backward through the drop-in module's autograd.Function).

Tolerance (BASELINE.json): parameter gradients within 1e-2 relative error.  Excluded exactly as SURVEY 8c says:
dc62.conv1.weight (dead code in the reference, grad None) and the 18 SSE conv1.bias tensors whose true gradient is
identically zero (the bias cancels in the non-affine InstanceNorm; autograd yields ~1e-10 noise).

What the 1e-2 bar is measured against: the gradient is a DISCONTINUOUS function of the forward activations (LeakyReLU'
jumps from 0.01 to 1 at 0, 28 times in depth).  Any forward that stores activations in 16 bits moves a ~1e-3 fraction of
them across the kink relative to the fp32 forward, which changes conv-weight gradients by 2-15 % in L2 no matter how exact
the backward is (tools: the same numbers come out of the fp32 oracle with fp16 rounding emulated and an EXACT autograd
backward - see DESIGN.md "Numerics of the backward pass").  The backward kernels are therefore held to 1e-2 against
autograd of the oracle evaluated AT THE SAME FORWARD STATE (oracle.INJECT: the plan's stored conv inputs, raw conv outputs
and InstanceNorm statistics replace the oracle's own, gradients flow through the oracle's graph), i.e. the exact fp32
gradient of the function the CUDA forward actually computed.  Against the plain fp32 reference (and the reference's own
golden gradients) the smooth parameters - heads, side branches - are held to 1e-2 as well, and every kink-sensitive tensor
to 1.5 x ITS OWN measured floor: tests/golden/kink_floor.json, written by tools/kink_floor.py, holds the per-tensor error of
an EXACT backward under emulated fp16 storage for exactly these inputs (so a backward regression cannot hide under a
blanket tolerance)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import seunet_oracle as oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = 1e-2
KINK_MARGIN = 1.5  # allowed multiple of the per-tensor storage-rounding floor (tools/kink_floor.py), never below REL_TOL
with open(os.path.join(GOLDEN, "kink_floor.json")) as _fh:
    KINK_FLOOR = json.load(_fh)["cases"]


def _bound(case, name):
    """Tolerance of one tensor against a DIFFERENT (fp32) forward state: max(1e-2, 1.5 x measured floor of this tensor)."""
    from se_unet_airseg_b200 import _lib
    storage = "fp16" if _lib.lib().seunet_act_dtype() == 0 else "bf16"
    return max(REL_TOL, KINK_MARGIN * KINK_FLOOR[case][storage]["per_tensor"][name])


def _model(in_ch, seed, train):
    from se_unet_airseg_b200 import SE_UNet
    sd = oracle.init_params(in_ch, 1, seed=seed)
    m = SE_UNet(in_ch, 1)
    m.load_state_dict(sd)
    m = m.cuda()
    m.train(train)
    return m, sd


def _check_grads(named_grads, ref, label):
    worst = 0.0
    for name, g in named_grads.items():
        if name == "dc62.conv1.weight":
            assert g is None, "dc62 is dead code in the reference graph: its grad must stay None"
            continue
        assert g is not None, f"{name}: missing gradient"
        g = g.detach().cpu().double()
        r = ref[name]
        if name.endswith("conv1.bias"):
            assert g.abs().max().item() <= 1e-6, f"{name}: conv1.bias gradient must be ~0"
            continue
        rel = (g - r).norm().item() / max(r.norm().item(), 1e-30)
        worst = max(worst, rel)
        assert rel <= REL_TOL, f"{label} {name}: relative gradient error {rel:.3e}"
    print(f"{label}: worst relative gradient error {worst:.3e}")


@pytest.mark.parametrize("stage", [1, 2, 3])
def test_backward_matches_reference_golden(stage):
    """tests/golden/train_stage*.npz: train-mode forward + stage loss + backward of the UNMODIFIED reference module."""
    z = np.load(os.path.join(GOLDEN, f"train_stage{stage}_c2_16.npz"))
    m, _ = _model(int(z["in_channel"]), int(z["seed"]), train=True)
    x = torch.from_numpy(z["x"]).cuda()
    label, weight, skel = (torch.from_numpy(z[k]).cuda() for k in ("label", "weight", "skel"))
    torch.manual_seed(int(z["torch_seed"]))      # same CPU-generator DropLayer draws as the reference run
    pe, pd = m(x)
    assert (pe.detach().cpu() - torch.from_numpy(z["pred0"])).abs().max().item() <= 2e-2
    assert (pd.detach().cpu() - torch.from_numpy(z["pred1"])).abs().max().item() <= 2e-2
    loss = oracle.stage_loss(stage, pe, pd, label, weight, skel)
    assert abs(loss.item() - float(z["loss"])) <= 2e-3
    loss.backward()
    worst, tight = 0.0, 0.0
    for name, p in m.named_parameters():
        if name == "dc62.conv1.weight":
            assert p.grad is None
            continue
        g = p.grad.detach().cpu().double()
        if name.endswith("conv1.bias"):
            assert g.abs().max().item() <= 1e-6
            continue
        ref_norm = float(z["gnorm." + name])
        if "grad." + name in z.files:
            r = torch.from_numpy(z["grad." + name]).double()
            rel = (g - r).norm().item() / max(r.norm().item(), 1e-30)
        else:
            rel = abs(g.norm().item() - ref_norm) / max(ref_norm, 1e-30)   # large tensors: only the norm is stored
        worst = max(worst, rel)
        smooth = name.startswith("dc0_") or ".conv2." in name
        bound = REL_TOL if smooth else _bound(f"golden:{stage}", name)
        tight = max(tight, rel / bound)
        assert rel <= bound, f"stage {stage} {name}: relative gradient error {rel:.3e} > {bound:.3e} (1.5 x storage floor)"
    print(f"stage {stage}: worst relative gradient error vs the fp32 reference's own gradients {worst:.3e}; "
          f"tightest tensor at {tight:.2f} of its bound")


# (1, 24, 24, 40): the coarsest level has 3 x 3 x 5 = 45 voxels - not a multiple of a warp's voxel group in any backward kernel
@pytest.mark.parametrize("in_ch,shape,train", [(2, (1, 16, 24, 16), False), (1, (2, 16, 16, 16), True), (2, (1, 32, 32, 32), False),
                                               (2, (1, 24, 24, 40), False)])
def test_backward_matches_oracle_autograd_at_same_forward_state(cuda_lib, in_ch, shape, train):
    """Every parameter tensor within 1e-2 of the exact fp32 gradient at the forward state the CUDA path computed."""
    from _plan_state import forward_state
    B, D, H, W = shape
    m, sd = _model(in_ch, 4242, train=train)
    g = torch.Generator().manual_seed(9)
    x = torch.rand(B, in_ch, D, H, W, generator=g)
    label = (torch.rand(B, 1, D, H, W, generator=g) > 0.9).float()
    weight = torch.where(label > 0, torch.rand(B, 1, D, H, W, generator=g) * 2 + 0.5, torch.ones(B, 1, D, H, W))
    torch.manual_seed(77)
    d0 = oracle.drop_scale(B, 24) if train else None
    d1 = oracle.drop_scale(B, 12) if train else None
    torch.manual_seed(77)
    p0, p1 = m(x.cuda())
    oracle.stage_loss(2, p0, p1, label.cuda(), weight.cuda()).backward()
    torch.cuda.synchronize()
    plan = m._plan(B, D, H, W, 1, torch.device("cuda", 0))
    sdt = torch.float16 if cuda_lib.seunet_act_dtype() == 0 else torch.bfloat16
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    oracle.INJECT = forward_state(cuda_lib, plan, shape, in_ch, sd, sdt)
    try:
        r0, r1 = oracle.forward(sdr, x, d0, d1)
        assert (r0 - p0.detach().cpu()).abs().max().item() < 1e-3 and (r1 - p1.detach().cpu()).abs().max().item() < 1e-3
        oracle.stage_loss(2, r0, r1, label, weight).backward()
    finally:
        oracle.INJECT = None
    ref = {k: (v.grad.double() if v.grad is not None else None) for k, v in sdr.items()}
    _check_grads({n: p.grad for n, p in m.named_parameters()}, ref, f"in_ch={in_ch} {shape} train={train}")


def test_backward_vs_plain_fp32_reference_reports_inherent_error():
    """Informational + sanity bound against the UNINJECTED fp32 oracle: smooth parameters within 1e-2, kink-sensitive ones
    within the inherent level; the flattened gradient must still point the same way (cosine > 0.98)."""
    in_ch, (B, D, H, W) = 2, (1, 32, 32, 32)
    m, sd = _model(in_ch, 4242, train=False)
    g = torch.Generator().manual_seed(9)
    x = torch.rand(B, in_ch, D, H, W, generator=g)
    label = (torch.rand(B, 1, D, H, W, generator=g) > 0.9).float()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    r0, r1 = oracle.forward(sdr, x)
    oracle.stage_loss(1, r0, r1, label).backward()
    p0, p1 = m(x.cuda())
    oracle.stage_loss(1, p0, p1, label.cuda()).backward()
    ours, refs, worst, tight = [], [], 0.0, 0.0
    for n, p in m.named_parameters():
        if p.grad is None or n.endswith("conv1.bias"):
            continue
        a, r = p.grad.cpu().double().flatten(), sdr[n].grad.double().flatten()
        rel = (a - r).norm().item() / max(r.norm().item(), 1e-30)
        smooth = n.startswith("dc0_") or ".conv2." in n
        bound = REL_TOL if smooth else _bound("plain32", n)
        tight = max(tight, rel / bound)
        assert rel <= bound, f"{n}: {rel:.3e} > {bound:.3e} (1.5 x storage floor)"
        worst = max(worst, rel)
        ours.append(a); refs.append(r)
    a, r = torch.cat(ours), torch.cat(refs)
    cos = (a @ r / (a.norm() * r.norm())).item()
    glob = ((a - r).norm() / r.norm()).item()
    from se_unet_airseg_b200 import _lib
    fl = KINK_FLOOR["plain32"]["fp16" if _lib.lib().seunet_act_dtype() == 0 else "bf16"]
    print(f"vs plain fp32 reference: worst per-tensor rel err {worst:.3e} (tightest tensor at {tight:.2f} of its bound), "
          f"global rel err {glob:.3e} (floor {fl['global_rel']:.3e}), cosine {cos:.5f} (floor {fl['cosine']:.5f})")
    assert glob <= KINK_MARGIN * fl["global_rel"]
    assert 1.0 - cos <= KINK_MARGIN * (1.0 - fl["cosine"])


def test_backward_guard_against_workspace_reuse():
    from se_unet_airseg_b200._lib import SeunetError
    m, _ = _model(2, 1, train=False)
    x = torch.rand(1, 2, 16, 16, 16, device="cuda")
    p0, _ = m(x)
    q0, _ = m(x)          # second forward of the same shape overwrites the saved activations
    with pytest.raises(SeunetError):
        p0.sum().backward()
    q0.sum().backward()


def test_backward_full_size_is_linear_in_the_output_gradient():
    """BASELINE training patch size (128^3).  CPU autograd of the oracle takes ~25 s per patch, so the full-size backward is
    checked through a size-independent property: with the forward state fixed, the backward is a LINEAR map of the output
    gradients, grads(a + b) = grads(a) + grads(b) (tensor-core dgrad/wgrad chains, the per-layer power-of-two dY scaling,
    InstanceNorm/gate/pool adjoints and the head fold all have to commute with addition), and it is not identically zero."""
    m, _ = _model(2, 4242, train=False)
    g = torch.Generator().manual_seed(77)
    x = torch.rand(1, 2, 128, 128, 128, generator=g).cuda()
    ga0, ga1 = (torch.randn(1, 1, 128, 128, 128, generator=g).cuda() * 1e-3 for _ in range(2))
    gb0, gb1 = (torch.randn(1, 1, 128, 128, 128, generator=g).cuda() * 1e-3 for _ in range(2))
    params = [p for n, p in m.named_parameters() if not n.startswith("dc62") and not n.endswith("conv1.bias")]

    def grads(d0, d1):
        p0, p1 = m(x)
        gs = torch.autograd.grad([p0, p1], params, [d0, d1], allow_unused=True)
        return torch.cat([gg.reshape(-1) for gg in gs if gg is not None]).double()

    a, b, ab = grads(ga0, ga1), grads(gb0, gb1), grads(ga0 + gb0, ga1 + gb1)
    assert a.norm().item() > 0 and b.norm().item() > 0
    rel = ((a + b) - ab).norm().item() / ab.norm().item()
    print(f"128^3 backward linearity: |g(a)+g(b)-g(a+b)| / |g(a+b)| = {rel:.3e}")
    assert rel <= REL_TOL
