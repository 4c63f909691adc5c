"""Dev check: tcgen05 conv vs torch GPU conv3d (fp32, TF32 off) at sizes with many tiles per CTA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from se_unet_airseg_b200 import _lib
from test_gpu_conv import _run_conv, _store_dtype
L = _lib.lib()
sdt = _store_dtype(L)
for (cin, cout, k, dil, shape) in [(64, 32, 3, 1, (1, 64, 64, 64)), (16, 32, 3, 2, (1, 64, 64, 64)), (8, 16, 3, 1, (1, 64, 64, 64)),
                                   (128, 64, 3, 1, (1, 32, 64, 64)), (56, 32, 1, 0, (1, 64, 64, 64)), (32, 16, 3, 1, (2, 64, 64, 64)),
                                   (64, 64, 3, 2, (1, 48, 48, 48))]:
    g = torch.Generator().manual_seed(1)
    N, D, H, W = shape
    x = torch.randn(N, cin, D, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, k, generator=g) / (cin * k ** 3) ** 0.5
    ref = F.conv3d(x.to(sdt).float().cuda(), w.to(sdt).float().cuda(), padding=dil if k == 3 else 0, dilation=max(dil, 1)).cpu()
    for rep in range(2):
        y, stats = _run_conv(L, x, w, k, dil)
        err = (y - ref).abs()
        print(cin, cout, k, dil, shape, "rep", rep, "max err", err.max().item(), "bad frac", (err > 1e-2).float().mean().item(), "nan", torch.isnan(y).any().item())
