"""Generates tests/golden/*.npz from the UNMODIFIED reference module (/root/reference/SE_UNet.py).

Run in the build container (the reference checkout does not travel to the GPU box):
    python oracle/make_golden.py
The reference ships no golden vectors of its own; these files pin both the oracle
(tests/test_oracle_golden.py, CPU) and the CUDA path (tests/test_gpu_*.py) to the reference's own outputs.

Train mode: the reference DropLayer hard-codes `.cuda()` (SE_UNet.py:91), which fails on a CPU-only host;
for the train-mode vectors that single call is neutralised (Tensor.cuda -> identity) - the math of
SE_UNet.py:90-95 runs unmodified.  Losses are extracted from train.py with `ast` (train.py imports
packages that are not installed here); only the three FunctionDef nodes are executed.
"""
import ast
import importlib.util
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("SEUNET_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import seunet_oracle as oracle  # noqa: E402


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_SE_UNet", os.path.join(REF, "SE_UNet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_losses():
    src = open(os.path.join(REF, "train.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("dice_loss", "general_union_loss_lib", "atr_loss"):
            exec(compile(ast.Module([node], []), "train.py", "exec"), ns)
    return ns


def synth_targets(shape, seed):
    g = torch.Generator().manual_seed(seed)
    label = (torch.rand(shape, generator=g) > 0.9).float()
    weight = torch.where(label > 0, torch.rand(shape, generator=g) * 2.0 + 0.5, torch.ones(shape))
    skel = label * (torch.rand(shape, generator=g) > 0.5).float()
    return label, weight, skel


def grad_summary(model):
    out = {}
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach()
        out["gsum." + n] = np.float64(g.double().sum().item())
        out["gnorm." + n] = np.float64(g.double().norm().item())
        if g.numel() <= 4096:
            out["grad." + n] = g.numpy().copy()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    losses = load_reference_losses()
    cases = [("forward_c2_16", 2, (1, 16, 16, 16), 777), ("forward_c1_16x24x16", 1, (2, 16, 24, 16), 778)]
    for name, ic, (B, D, H, W), seed in cases:
        sd = oracle.init_params(ic, 1, seed=seed)
        m = ref.SE_UNet(ic, 1)
        m.load_state_dict(sd, strict=True)
        m.eval()
        x = torch.rand(B, ic, D, H, W, generator=torch.Generator().manual_seed(seed + 1))
        with torch.no_grad():
            p0, p1 = m(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.numpy(), pred0=p0.numpy(), pred1=p1.numpy(),
                            in_channel=ic, seed=seed)
        print(name, "pred0 std", p0.std().item(), "pred1 std", p1.std().item())

    # train-mode forward + backward for the three stage losses (train.py:597-599, 432-435, 238-243)
    ic, (B, D, H, W), seed = 2, (2, 16, 16, 16), 779
    sd = oracle.init_params(ic, 1, seed=seed)
    x = torch.rand(B, ic, D, H, W, generator=torch.Generator().manual_seed(seed + 1))
    label, weight, skel = synth_targets((B, 1, D, H, W), seed + 2)
    orig_cuda = torch.Tensor.cuda
    for stage in (1, 2, 3):
        m = ref.SE_UNet(ic, 1)
        m.load_state_dict(sd, strict=True)
        m.train()
        torch.manual_seed(1000 + stage)
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            pe, pd = m(x)
        finally:
            torch.Tensor.cuda = orig_cuda
        se, sdg = torch.sigmoid(pe), torch.sigmoid(pd)
        if stage == 1:
            loss = losses["dice_loss"](sdg, label) + losses["dice_loss"](se, label)
        else:
            loss = losses["general_union_loss_lib"](sdg, label, weight) * 1 + losses["general_union_loss_lib"](se, label, weight) * 0.5
            if stage == 3:
                loss = loss + 0.5 * (losses["atr_loss"](se, label, skel, weight) + losses["atr_loss"](sdg, label, skel, weight))
        loss.backward()
        data = dict(x=x.numpy(), label=label.numpy(), weight=weight.numpy(), skel=skel.numpy(), pred0=pe.detach().numpy(),
                    pred1=pd.detach().numpy(), loss=np.float64(loss.item()), in_channel=ic, seed=seed, stage=stage,
                    torch_seed=1000 + stage)
        data.update(grad_summary(m))
        np.savez_compressed(os.path.join(OUT, f"train_stage{stage}_c2_16.npz"), **data)
        print("stage", stage, "loss", loss.item())


def load_reference_dti():
    """prediction.py imports pyvista/skimage/stl (not installed here): only the FunctionDef of double_threshold_iteration
    is executed, with numpy as its sole global."""
    src = open(os.path.join(REF, "prediction.py")).read()
    ns = {"np": np}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == "double_threshold_iteration":
            exec(compile(ast.Module([node], []), "prediction.py", "exec"), ns)
    return ns["double_threshold_iteration"]


def synth_prob(shape, seed):
    """Smooth random probability field with strong cores, weak halos and isolated weak specks (float32, like the mean of
    sigmoid outputs), plus exact threshold values."""
    rng = np.random.RandomState(seed)
    from scipy import ndimage
    f = ndimage.gaussian_filter(rng.rand(*shape), 1.2)
    f = (f - f.min()) / (f.max() - f.min())
    p = (0.15 + 0.7 * f + 0.08 * (rng.rand(*shape) - 0.5)).astype(np.float32)
    flat = p.reshape(-1)
    flat[rng.randint(0, flat.size, 8)] = np.float32(0.5)
    flat[rng.randint(0, flat.size, 8)] = np.float32(0.4)
    return p


def main_postproc():
    dti = load_reference_dti()
    for shape, seed in (((12, 14, 20), 1), ((9, 16, 37), 2), ((16, 10, 70), 3)):
        p = synth_prob(shape, seed)
        ref = dti(p.copy(), h_thresh=0.5, l_thresh=0.4)
        mine = oracle.double_threshold_iteration(p, 0.5, 0.4)
        assert np.array_equal(ref, mine), "oracle restatement differs from the reference's double_threshold_iteration"
        name = "postproc_dti_%dx%dx%d.npz" % shape
        np.savez_compressed(os.path.join(OUT, name), prob=p, dti=ref.astype(np.uint8), seed=seed)
        print(name, "set voxels", int(ref.sum()), "strong", int((p >= 0.5).sum()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "postproc":
        main_postproc()
    else:
        main()
        main_postproc()
