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


# ---------------------------------------------------------------------------------------------------------------------
# maximum_3d (util.py:58-75).  The function's TEXT is the reference's own (extracted with ast, executed unmodified); the two
# third-party calls it makes are not installed in this container (connected-components-3d 3.x `cc3d.connected_components`,
# scikit-image `measure.regionprops`), so they are supplied by the two small stand-ins below, which restate the published
# behaviour of exactly the features the call sites use and are validated against an independent implementation
# (scipy.ndimage.label with the full 3x3x3 structure) before any fixture is written:
#   cc3d.connected_components(binary, connectivity=26): 26-connected components of the non-zero voxels, labels 1..N numbered
#     in raster order (C order for a C-contiguous array) of each component's first voxel;
#   measure.regionprops(label): one entry per label in ascending label order, `.area` = voxel count.
# ---------------------------------------------------------------------------------------------------------------------
class _CC3DStandIn:
    @staticmethod
    def connected_components(region, connectivity=26):
        assert connectivity == 26
        a = np.asarray(region) != 0
        lab = np.zeros(a.shape, dtype=np.uint32)
        nxt = 0
        D, H, W = a.shape
        for i in range(D):
            for j in range(H):
                for k in range(W):
                    if a[i, j, k] and lab[i, j, k] == 0:        # first voxel of a new component in raster order
                        nxt += 1
                        lab[i, j, k] = nxt
                        stack = [(i, j, k)]
                        while stack:
                            x, y, z = stack.pop()
                            for u in range(max(x - 1, 0), min(x + 2, D)):
                                for v in range(max(y - 1, 0), min(y + 2, H)):
                                    for w in range(max(z - 1, 0), min(z + 2, W)):
                                        if a[u, v, w] and lab[u, v, w] == 0:
                                            lab[u, v, w] = nxt
                                            stack.append((u, v, w))
        return lab


class _RegionPropsStandIn:
    class _R:
        def __init__(self, area):
            self.area = area

    @staticmethod
    def regionprops(label):
        cnt = np.bincount(np.asarray(label).ravel())
        return [_RegionPropsStandIn._R(int(c)) for c in cnt[1:] if c > 0]


def load_reference_maximum_3d():
    from scipy.ndimage import binary_fill_holes
    src = open(os.path.join(REF, "util.py")).read()
    ns = {"np": np, "cc3d": _CC3DStandIn, "measure": _RegionPropsStandIn, "binary_fill_holes": binary_fill_holes}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == "maximum_3d":
            exec(compile(ast.Module([node], []), "util.py", "exec"), ns)
    return ns["maximum_3d"]


def max3d_cases():
    """(name, binary int8 volume): random blobs, an exact AREA TIE (the reversed stable sort picks the later label), the
    probe-slice fallback (largest component misses the slices W//2, W//3, W//3*2 of the last axis), diagonal (26- but not
    6-/18-) connectivity, enclosed holes."""
    from scipy import ndimage
    out = []
    for seed, shape, thr in ((1, (14, 16, 30), 0.58), (2, (10, 12, 41), 0.62), (3, (20, 9, 24), 0.55)):
        rng = np.random.RandomState(seed)
        f = ndimage.gaussian_filter(rng.rand(*shape), 1.0)
        f = (f - f.min()) / (f.max() - f.min())
        out.append((f"blobs{seed}", (f > thr).astype(np.int8)))
    tie = np.zeros((8, 8, 12), dtype=np.int8)
    tie[1:3, 1:3, 5:8] = 1            # 12 voxels, label 1 (first in raster order), touches slice 6 = 12 // 2
    tie[5:7, 5:7, 4:7] = 1            # 12 voxels, label 2 -> wins the tie
    tie[4, 0, 0] = 1                  # a 1-voxel component
    out.append(("tie", tie))
    probe = np.zeros((9, 9, 30), dtype=np.int8)
    probe[1:8, 1:8, 0:4] = 1          # largest (196 voxels) but away from slices 15, 10, 20
    probe[3:6, 3:6, 9:22] = 1         # second largest, crosses all three probe slices
    probe[0, 0, 28] = 1
    out.append(("probe", probe))
    diag = np.zeros((7, 7, 9), dtype=np.int8)
    for t in range(6):
        diag[t, t, t + 1] = 1         # corner-connected chain: one component only under 26-connectivity
    diag[5:7, 0:2, 3:6] = 1
    out.append(("diag", diag))
    hole = np.zeros((9, 9, 9), dtype=np.int8)
    hole[1:8, 1:8, 1:8] = 1
    hole[3:6, 3:6, 3:6] = 0           # enclosed cavity: filled by binary_fill_holes
    hole[4, 4, 4] = 1                 # an island inside the cavity (a separate, smaller component)
    out.append(("hole", hole))
    return out


def main_max3d():
    from scipy import ndimage
    # validate the stand-in against an independent implementation, numbering included
    for seed in range(6):
        rng = np.random.RandomState(100 + seed)
        a = (rng.rand(9, 11, 13) > 0.72).astype(np.int8)
        mine = _CC3DStandIn.connected_components(a, connectivity=26)
        ref, n = ndimage.label(a, structure=np.ones((3, 3, 3)))
        assert n == mine.max() and np.array_equal(mine.astype(np.int64), ref.astype(np.int64)), "cc3d stand-in is not faithful"
        areas = [r.area for r in _RegionPropsStandIn.regionprops(mine)]
        assert areas == [int((ref == i).sum()) for i in range(1, n + 1)]
    fn = load_reference_maximum_3d()
    data = {}
    for name, vol in max3d_cases():
        ref = np.asarray(fn(vol.copy())).astype(np.uint8)
        mine = oracle.maximum_3d(vol).astype(np.uint8)
        assert np.array_equal(ref, mine), f"oracle.maximum_3d differs from the reference's function on case {name}"
        data["in." + name] = vol
        data["out." + name] = ref
        print("max3d", name, vol.shape, "in", int(vol.sum()), "out", int(ref.sum()))
    np.savez_compressed(os.path.join(OUT, "postproc_max3d.npz"), **data)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "postproc":
        main_postproc()
    elif len(sys.argv) > 1 and sys.argv[1] == "max3d":
        main_max3d()
    else:
        main()
        main_postproc()
        main_max3d()
