"""GPU parity of the tcgen05 weight-gradient kernel (C ABI seunet_conv_wgrad) against torch's conv3d weight gradient
(fp64 on CPU, operands rounded to the storage types first).  Replaces autograd of nn.Conv3d (SE_UNet.py:15/42/57)."""
import pytest
import torch
import torch.nn.functional as F

from test_gpu_conv import _store_dtype, to_chunk_planes

pytestmark = pytest.mark.gpu

CASES = [
    # Cin, Cout, k, dil, (N, D, H, W)
    (64, 32, 3, 1, (1, 8, 16, 8)),
    (64, 32, 3, 1, (2, 8, 32, 16)),
    (128, 64, 3, 1, (1, 8, 16, 16)),
    (16, 32, 3, 2, (1, 8, 16, 8)),
    (32, 64, 3, 2, (1, 8, 16, 8)),
    (32, 16, 3, 1, (1, 8, 16, 8)),
    (8, 16, 3, 1, (1, 8, 16, 8)),
    (2, 8, 3, 1, (1, 8, 16, 8)),
    (56, 32, 1, 0, (1, 8, 16, 8)),
    (192, 64, 1, 0, (2, 4, 16, 8)),
    (96, 32, 1, 0, (1, 8, 16, 8)),
    (32, 32, 3, 1, (1, 6, 20, 12)),      # ragged tiles
    (64, 64, 3, 2, (1, 5, 24, 12)),
]


@pytest.mark.parametrize("Cin,Cout,k,dil,shape", CASES)
def test_wgrad_matches_autograd(cuda_lib, Cin, Cout, k, dil, shape):
    from se_unet_airseg_b200 import _lib
    L = cuda_lib
    sdt = _store_dtype(L)
    g = torch.Generator().manual_seed(4321 + Cin + Cout + k + dil)
    N, D, H, W = shape
    x = torch.randn(N, Cin, D, H, W, generator=g)
    dy = torch.randn(N, Cout, D, H, W, generator=g)
    xq = x.to(sdt).double().requires_grad_(False)
    dyq = dy.to(sdt).double()
    w = torch.zeros(Cout, Cin, k, k, k, dtype=torch.float64, requires_grad=True)
    y = F.conv3d(xq, w, padding=dil if k == 3 else 0, dilation=max(dil, 1))
    (ref,) = torch.autograd.grad(y, w, dyq)
    dev = torch.device("cuda", 0)
    COUT = 16 if Cout <= 16 else (32 if Cout <= 32 else 64)
    xc = 1 if Cin <= 8 else ((Cin + 15) // 16) * 2
    xin = to_chunk_planes(x.to(dev), xc + 3, 2, sdt)            # slice inside a wider buffer
    dyin = to_chunk_planes(dy.to(dev), COUT // 8 + 1, 1, sdt)
    scratch = torch.empty(L.seunet_wgrad_scratch_bytes(Cin, Cout, k), dtype=torch.uint8, device=dev)
    dw = torch.full((Cout, Cin, k, k, k), float("nan"), device=dev)
    st = _lib.stream_ptr()
    _lib.check(L.seunet_debug_poison_smem(st), "poison")
    _lib.check(L.seunet_conv_wgrad(_lib.ptr(xin), xc + 3, 2, _lib.ptr(dyin), COUT // 8 + 1, 1, N, D, H, W, Cin, Cout, k, dil,
                                   _lib.ptr(scratch), _lib.ptr(dw), st), "wgrad")
    torch.cuda.synchronize()
    err = (dw.cpu().double() - ref).abs().max().item()
    tol = 1e-4 * ref.abs().max().item() + 1e-4
    assert err <= tol, f"max abs err {err} (ref max {ref.abs().max().item()})"
