"""GPU parity of the tcgen05 implicit-GEMM convolution (C ABI seunet_conv_fprop) against F.conv3d (fp32, CPU).

Replaces nn.Conv3d call sites SE_UNet.py:15/57 (3x3x3, dilation 1|2) and SE_UNet.py:42 (1x1x1).
Tolerance: operands are rounded to the 16-bit storage type before both computations, accumulation is fp32
on both sides, so the only difference is the final rounding of the output to the storage type (fp16: 2^-11
relative, bf16: 2^-8) plus summation order.
"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _store_dtype(L):
    return torch.float16 if L.seunet_act_dtype() == 0 else torch.bfloat16


def to_chunk_planes(x, chunks, off, dtype):
    """(N,C,D,H,W) fp32 -> chunk planes [N][chunks][D][H][W][8] of `dtype` with x occupying planes [off, off+ceil(C/8))."""
    N, C, D, H, W = x.shape
    k = (C + 7) // 8
    buf = torch.zeros(N, chunks, D, H, W, 8, dtype=dtype, device=x.device)
    xp = torch.zeros(N, k * 8, D, H, W, dtype=x.dtype, device=x.device)
    xp[:, :C] = x
    buf[:, off:off + k] = xp.view(N, k, 8, D, H, W).permute(0, 1, 3, 4, 5, 2).to(dtype)
    return buf


def from_chunk_planes(buf, C):
    N, k, D, H, W, _ = buf.shape
    return buf.permute(0, 1, 5, 2, 3, 4).reshape(N, k * 8, D, H, W)[:, :C].float()


def _run_conv(L, x, w, ksize, dil, in_chunks=None, in_off=0, transpose_flip=0, bf16=0, accum_into=None, grad_out=0):
    from se_unet_airseg_b200 import _lib
    N, Cin, D, H, W = x.shape
    Cout = w.shape[1] if transpose_flip else w.shape[0]
    cin_k = w.shape[0] if transpose_flip else w.shape[1]
    assert cin_k == Cin
    dev = torch.device("cuda", 0)
    COUT = 16 if Cout <= 16 else (32 if Cout <= 32 else 64)
    sdt = torch.bfloat16 if bf16 else _store_dtype(L)
    if in_chunks is None:
        in_chunks = 1 if (Cin <= 8 and ksize == 3) else ((Cin + 15) // 16) * 2
    st = _lib.stream_ptr()
    xin = to_chunk_planes(x.to(dev), in_chunks, in_off, sdt)
    oc = (Cout + 7) // 8
    odt = torch.float32 if grad_out else sdt
    if accum_into is not None:
        out = to_chunk_planes(accum_into.to(dev), oc, 0, odt)
    else:
        out = torch.full((N, oc, D, H, W, 8), float("nan"), dtype=odt, device=dev)
    stats = torch.zeros(N * COUT * 2, dtype=torch.float64, device=dev)
    scratch = torch.empty(L.seunet_conv_scratch_bytes(Cin, Cout, ksize, dil), dtype=torch.uint8, device=dev)
    wd = w.to(dev).contiguous()
    _lib.check(L.seunet_debug_poison_smem(st), "poison_smem")   # stale shared memory must never reach the accumulators
    _lib.check(L.seunet_conv_fprop(_lib.ptr(xin), in_chunks, in_off, _lib.ptr(wd), N, D, H, W, Cin, Cout, ksize, dil,
                                   _lib.ptr(out), _lib.ptr(stats), _lib.ptr(scratch), transpose_flip, bf16, grad_out,
                                   1 if accum_into is not None else 0, st), "conv_fprop")
    torch.cuda.synchronize()
    return from_chunk_planes(out, Cout).cpu(), stats.cpu().view(N, COUT, 2)


CASES = [
    # Cin, Cout, k, dil, (N, D, H, W)
    (16, 32, 3, 1, (1, 8, 16, 8)),
    (16, 32, 3, 2, (1, 8, 16, 16)),     # ec3
    (32, 32, 3, 1, (2, 8, 16, 8)),      # ec4, batch 2
    (32, 64, 3, 2, (1, 8, 16, 8)),      # ec6
    (64, 32, 3, 1, (1, 16, 32, 16)),    # dc5 / dc4
    (64, 64, 3, 1, (1, 8, 16, 16)),     # ec7.. (weight ring)
    (64, 64, 3, 2, (1, 8, 16, 8)),      # ec8/ec9 (weight ring + dilation)
    (128, 64, 3, 1, (1, 8, 16, 8)),     # dc1 / dc3 (weight ring, 8 chunks)
    (32, 16, 3, 1, (1, 16, 16, 8)),     # dc6
    (8, 16, 3, 1, (1, 8, 16, 8)),       # ec2 (two taps per K=16 step)
    (2, 8, 3, 1, (1, 8, 16, 8)),        # ec1 (channels padded 2->8, Cout padded 8->16)
    (1, 8, 3, 1, (1, 8, 16, 8)),        # ec1 with in_channel=1
    (56, 32, 1, 0, (1, 8, 16, 8)),      # ec33 (56 of 64 channels used)
    (128, 64, 1, 0, (1, 8, 16, 8)),     # ec63 / dc22
    (192, 64, 1, 0, (1, 8, 16, 8)),     # ec93 / ec123
    (96, 32, 1, 0, (1, 8, 16, 8)),      # dc42
    # ragged shapes: partial tiles in every dimension (H not multiple of 16, W not multiple of 8, D partial)
    (32, 32, 3, 1, (1, 6, 20, 12)),
    (32, 32, 3, 2, (1, 5, 20, 12)),
    (64, 64, 3, 1, (1, 3, 8, 8)),
    (64, 32, 1, 0, (1, 5, 24, 12)),
]


@pytest.mark.parametrize("Cin,Cout,k,dil,shape", CASES)
def test_conv_fprop_matches_conv3d(cuda_lib, Cin, Cout, k, dil, shape):
    L = cuda_lib
    sdt = _store_dtype(L)
    g = torch.Generator().manual_seed(1234 + Cin * 7 + Cout * 3 + k + dil)
    N, D, H, W = shape
    x = torch.randn(N, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, k, generator=g) / (Cin * k ** 3) ** 0.5
    xq, wq = x.to(sdt).float(), w.to(sdt).float()
    ref = F.conv3d(xq.double(), wq.double(), padding=dil if k == 3 else 0, dilation=max(dil, 1)).float()
    y, stats = _run_conv(L, x, w, k, dil)
    eps = 2.0 ** -11 if sdt == torch.float16 else 2.0 ** -8
    tol = eps * ref.abs().max().item() * 1.5 + 1e-5
    err = (y - ref).abs().max().item()
    assert err <= tol, f"max abs err {err} > {tol}"
    V = D * H * W
    s_ref = ref.sum(dim=(2, 3, 4))
    q_ref = (ref * ref).sum(dim=(2, 3, 4))
    assert torch.allclose(stats[:, :Cout, 0].float(), s_ref, rtol=1e-4, atol=1e-3 * V ** 0.5)
    assert torch.allclose(stats[:, :Cout, 1].float(), q_ref, rtol=1e-4, atol=1e-3)


def test_conv_reads_channel_slice_of_concat_buffer(cuda_lib):
    """Virtual concat (SE_UNet.py:186 etc.): the conv input is a chunk slice of a wider buffer."""
    L = cuda_lib
    sdt = _store_dtype(L)
    g = torch.Generator().manual_seed(99)
    x = torch.randn(1, 16, 8, 16, 8, generator=g)
    w = torch.randn(32, 16, 3, 3, 3, generator=g) / 20.0
    ref = F.conv3d(x.to(sdt).double(), w.to(sdt).double(), padding=1).float()
    y, _ = _run_conv(L, x, w, 3, 1, in_chunks=8, in_off=5)
    assert (y - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item() * 1.5


def test_conv_dgrad_accumulates_into_gradient_buffer(cuda_lib):
    """Backward use: mirrored/transposed weights, fp32 gradient-format output, out += result; 40 output channels exercise
    the channel-plane mask (COUT padded to 64, only 5 planes written)."""
    L = cuda_lib
    sdt = _store_dtype(L)
    g = torch.Generator().manual_seed(8)
    w = torch.randn(32, 40, 3, 3, 3, generator=g) / 30.0          # forward conv 40 -> 32
    dy = torch.randn(2, 32, 8, 16, 8, generator=g)
    base = torch.randn(2, 40, 8, 16, 8, generator=g)
    q = lambda t: t.to(sdt).double()
    ref = (base.double() + F.conv_transpose3d(q(dy), q(w), padding=1)).float()
    y, _ = _run_conv(L, dy, w, 3, 1, transpose_flip=1, accum_into=base, grad_out=1)
    assert (y - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


def test_conv_dgrad_operator(cuda_lib):
    """transpose_flip=1 turns the same kernel into the data-gradient operator of conv3d."""
    L = cuda_lib
    sdt = _store_dtype(L)
    g = torch.Generator().manual_seed(5)
    Cin, Cout = 32, 64   # forward conv: Cin -> Cout ; dgrad maps dy (Cout ch) -> dx (Cin ch)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / 30.0
    dy = torch.randn(1, Cout, 8, 16, 8, generator=g)
    ref = F.conv_transpose3d(dy.to(sdt).double(), w.to(sdt).double(), padding=2, dilation=2).float()
    # C-ABI: "Cin" is the operator's input width (forward Cout), "Cout" its output width (forward Cin)
    from se_unet_airseg_b200 import _lib  # noqa: F401
    y, _ = _run_conv(L, dy, w, 3, 2, transpose_flip=1)
    assert (y - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item() * 1.5
