"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/seunet_b200.h declares, the ctypes table
mirrors the header, and the host logic that needs no device behaves (parameter table, error paths)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "seunet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(seunet_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cuda_lib):
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(cuda_lib, n), f"{n} declared in include/seunet_b200.h but not exported"


def test_ctypes_table_mirrors_header():
    from se_unet_airseg_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_functions()


def test_param_table_matches_module_state_dict(cuda_lib):
    from se_unet_airseg_b200 import SE_UNet
    for ic in (1, 2):
        m = SE_UNet(ic, 1)
        sd = m.state_dict()
        assert cuda_lib.seunet_param_tensors(ic, 1) == len(sd) == 117
        off = 0
        for i, (k, v) in enumerate(sd.items()):
            assert cuda_lib.seunet_param_name(ic, 1, i).decode() == k
            assert cuda_lib.seunet_param_numel(ic, 1, i) == v.numel()
            assert cuda_lib.seunet_param_offset(ic, 1, k.encode()) == off
            off += v.numel()
        assert cuda_lib.seunet_param_count(ic, 1) == off


def test_errors_are_reported_not_swallowed(cuda_lib):
    assert cuda_lib.seunet_param_count(7, 1) == -1
    assert b"in_channel" in cuda_lib.seunet_last_error()
    h = ctypes.c_void_p()
    assert cuda_lib.seunet_plan_create(ctypes.byref(h), 1, 30, 32, 32, 2, 1, 0, 0) != 0      # 30 is not a multiple of 8
    assert b"multiples of 8" in cuda_lib.seunet_last_error()


def test_module_refuses_cpu_tensors_and_matches_reference_surface():
    from se_unet_airseg_b200 import SE_UNet, get_model, config, SSEConv, SSEConv2, CATConv, DropLayer
    from se_unet_airseg_b200._lib import SeunetError
    cfg, net = get_model()
    assert cfg is config and net.in_channel == 2 and net.n_classes == 1
    for attr, val in (("batchnorm", False), ("bias", True), ("out_channel2", 2), ("sigmoid_output", 0)):
        assert getattr(net, attr) == val
    assert isinstance(net.ec1, SSEConv) and isinstance(net.ec4, SSEConv2) and isinstance(net.ec33, CATConv)
    assert isinstance(net.dropout1, DropLayer) and net.dropout1.channel_num == 24 and net.dropout2.threshold == 0.3
    with pytest.raises(SeunetError):
        net(torch.zeros(1, 2, 16, 16, 16))          # no CPU fallback for the hot path
    # DropLayer factor semantics (SE_UNet.py:89-97) on the host
    net.train()
    torch.manual_seed(3)
    r = net.dropout1.scale(4, torch.device("cpu"))
    torch.manual_seed(3)
    u = torch.rand(4, 24, 1, 1, 1)
    keep = (u >= 0.3).float()
    assert torch.allclose(r, (keep * 24 / (keep.sum() + 0.01)).reshape(4, 24))
    net.eval()
    assert torch.equal(net.dropout1.scale(2, torch.device("cpu")), torch.ones(2, 24))


def test_root_shim_exposes_reference_import_name():
    import SE_UNet as shim
    from se_unet_airseg_b200 import SE_UNet as cls
    assert shim.SE_UNet is cls and hasattr(shim, "get_model") and shim.config == {}
