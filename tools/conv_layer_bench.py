"""Per-layer timing of the tcgen05 conv (C ABI seunet_conv_fprop) at network shapes. Dev tool.
usage: python tools/conv_layer_bench.py [B] [S] [layer-name-filter]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se_unet_airseg_b200 import _lib
L = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
flt = sys.argv[3] if len(sys.argv) > 3 else ""
LAYERS = [("ec1", 2, 8, 3, 1, 0), ("ec2", 8, 16, 3, 1, 0), ("ec3", 16, 32, 3, 2, 0), ("ec33", 56, 32, 1, 0, 0),
          ("ec4", 32, 32, 3, 1, 1), ("ec5", 32, 32, 3, 2, 1), ("ec6", 32, 64, 3, 2, 1), ("ec63", 128, 64, 1, 0, 1),
          ("ec7", 64, 64, 3, 1, 2), ("ec8", 64, 64, 3, 2, 2), ("ec93", 192, 64, 1, 0, 2), ("ec10", 64, 64, 3, 1, 3),
          ("dc1", 128, 64, 3, 1, 2), ("dc22", 128, 64, 1, 0, 2), ("dc3", 128, 64, 3, 1, 1), ("dc4", 64, 32, 3, 1, 1),
          ("dc42", 96, 32, 1, 0, 1), ("dc5", 64, 32, 3, 1, 0), ("dc6", 32, 16, 3, 1, 0)]
dev = torch.device("cuda", 0)
sdt = torch.float16 if L.seunet_act_dtype() == 0 else torch.bfloat16
st = _lib.stream_ptr()
tot = 0.0
for name, cin, cout, k, dil, lvl in LAYERS:
    if flt and name not in flt.split(","):
        continue
    s = S >> lvl
    COUT = 16 if cout <= 16 else (32 if cout <= 32 else 64)
    chunks = 1 if (cin <= 8 and k == 3) else ((cin + 15) // 16) * 2
    V = s ** 3
    xin = (torch.randn(B * chunks * V * 8, device=dev) * 0.5).to(sdt)
    w = torch.randn(cout, cin, k, k, k, device=dev) / (cin * k ** 3) ** 0.5
    out = torch.empty(B * (COUT // 8) * V * 8, dtype=sdt, device=dev)
    stats = torch.zeros(B * COUT * 2, dtype=torch.float64, device=dev)
    scratch = torch.empty(L.seunet_conv_scratch_bytes(cin, cout, k, dil), dtype=torch.uint8, device=dev)
    run = lambda: _lib.check(L.seunet_conv_fprop(_lib.ptr(xin), chunks, 0, _lib.ptr(w), B, s, s, s, cin, cout, k, dil,
                                                 _lib.ptr(out), _lib.ptr(stats), _lib.ptr(scratch), 0, 0, 0, 0, st), "conv")
    for _ in range(int(os.environ.get("CONV_BENCH_WARMUP", "2"))):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = int(os.environ.get("CONV_BENCH_ITERS", "5"))
    e0.record()
    for _ in range(it):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / it * 1e3
    fl = 2.0 * B * V * cin * cout * k ** 3
    by = B * V * 2.0 * (chunks * 8 + COUT)
    tot += us
    print(f"{name:6s} {cin:3d}->{cout:2d} k{k} d{dil} {s:3d}^3: {us:8.1f} us  {fl/us/1e6:7.1f} TFLOP/s  {by/us/1e3:7.1f} GB/s")
print(f"total {tot:.1f} us")
