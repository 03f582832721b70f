"""The C-ABI library loads on a CPU-only box and exports every symbol include/fsgs_raster.h declares
(no compute calls here); argument validation that needs no device; the drop-in module names resolve."""
import ctypes
import os
import re

import pytest
import torch

from fsgs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "fsgs_raster.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fsgs_[a-z0-9_]+)\s*\(", hdr)) - {"fsgs_alloc_fn"})


def test_every_declared_symbol_is_exported():
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/fsgs_raster.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) <= set(declared)


def test_host_only_helpers():
    L = _lib.lib()
    assert L.fsgs_abi_version() == 1
    assert L.fsgs_geom_bytes(1000) >= 1000 * 48 + 1000
    assert L.fsgs_binning_bytes(10) >= 10 * 56
    assert L.fsgs_grad_scratch_bytes(1000) >= 1000 * 48
    io = _lib.img_offsets(1280, 1024)
    assert io["n_contrib"] >= 1280 * 1024 * 4 and io["tile_offset"] > io["tile_count"]
    assert L.fsgs_img_bytes(1280, 1024) > io["counters"]
    assert _lib.binning_offsets(100)["records"] == 0 and _lib.binning_offsets(100)["keys"] >= 4800
    assert L.fsgs_error_string(-4).decode().startswith("device is not compute capability 10")
    assert "k_composite_bwd" in _lib.kernel_names() and len(_lib.kernel_names()) >= 16 and "k_preprocess_pose_bwd" in _lib.kernel_names()


def test_invalid_arguments_are_rejected_before_any_cuda_call():
    L = _lib.lib()
    st = _lib.Settings(0, 0, 1.0, 1.0, 1.0, 0, 0, 0, 0)        # zero-sized image
    rc = L.fsgs_mark_visible(-1, None, None, None, None, None)
    assert rc == -1
    nr = ctypes.c_int64(0)
    cb = _lib.ALLOC_FN(lambda u, n: None)
    rc = L.fsgs_rasterize_forward(ctypes.byref(st), 0, *([None] * 11), cb, None, cb, None, cb, None, None, None, None,
                                  ctypes.byref(nr), None, None)
    assert rc == -1
    st = _lib.Settings(64, 64, 1.0, 1.0, 1.0, 7, 0, 0, 0)      # unsupported SH degree
    rc = L.fsgs_render_backward(ctypes.byref(st), 1, 0, *([None] * 16), 1, 1, *([None] * 9))
    assert rc == -1
    # frozen-model forward: NULL / mis-aligned arguments are rejected before any device call
    st_ok = _lib.Settings(64, 64, 1.0, 1.0, 1.0, 3, 16, 0, 0)
    assert L.fsgs_frozen_bytes(1000) == 64000 and L.fsgs_frozen_bytes(0) == 64
    assert L.fsgs_freeze_model(ctypes.byref(st_ok), 5, None, None, None, None, None, None, None, None, None) == -1
    one_ = ctypes.c_void_p(256)
    assert L.fsgs_freeze_model(ctypes.byref(st_ok), 5, one_, one_, one_, one_, one_, one_, one_, ctypes.c_void_p(264), None) == -1   # rows not 16-byte aligned
    assert L.fsgs_render_forward_frozen(ctypes.byref(st_ok), 5, one_, None, one_, one_, one_, cb, None, cb, None, cb, None,
                                        one_, one_, None, None, None, None) == -1                                     # rows NULL
    assert L.fsgs_fixed_bin_capacity(-1) == -1 and L.fsgs_fixed_bin_capacity(0) >= 0
    # frame-parallel exchange: rank / world / pointer checks come first
    ptrs = (ctypes.c_void_p * 2)(256, 512)
    assert L.fsgs_exchange_rows_scatter(None, None, 2, 0, 0, 16, 0, None) == -1                  # no buffers at all
    assert L.fsgs_exchange_rows_scatter(one_, None, 2, 0, 0, 16, 4, None) == -1                  # scatter needs peer pointers
    assert L.fsgs_exchange_rows_scatter(None, ptrs, 2, 2, 0, 16, 0, None) == -1                  # rank outside the world
    assert L.fsgs_exchange_rows_scatter(None, ptrs, 1, 0, 0, 16, 0, None) == 0                   # one rank: nothing to do
    bad = (ctypes.c_void_p * 2)(256, 520)                                                        # second buffer mis-aligned
    assert L.fsgs_compact_grad_expand_peers(ctypes.byref(st_ok), 512, 0, 512, one_, one_, bad, 2, 0, 0, 0,
                                            one_, one_, one_, one_, one_, one_, None) == -1
    assert L.fsgs_compact_grad_expand_peers(ctypes.byref(st_ok), 512, 100, 412, one_, one_, ptrs, 2, 0, 0, 0,
                                            one_, one_, one_, one_, one_, one_, None) == -1      # first % 256 != 0
    assert L.fsgs_compact_grad_expand_peers(ctypes.byref(st_ok), 512, 0, 512, one_, one_, ptrs, 1, 0, 0, 0,
                                            one_, one_, one_, one_, one_, one_, None) == -1      # needs >= 2 ranks
    # fused image loss: argument checks come before any device call
    one = ctypes.c_void_p(256)                                 # a non-NULL placeholder; never dereferenced
    assert L.fsgs_rgb_loss_forward(0, 8, 8, one, one, None, None, 0, 0.2, None, one, one, None) == -1      # C = 0
    assert L.fsgs_rgb_loss_forward(3, 8, 8, None, one, None, None, 0, 0.2, None, one, one, None) == -1     # img NULL
    assert L.fsgs_rgb_loss_forward(3, 8, 8, one, one, one, one, 0, 0.2, None, one, one, None) == -1        # two masks
    assert L.fsgs_rgb_loss_forward(3, 8, 8, one, one, one, None, 7, 0.2, None, one, one, None) == -1       # bad stride
    assert L.fsgs_rgb_loss_forward(3, 8, 8, one, one, None, None, 0, 0.2, None, None, one, None) == -1     # no scratch
    assert L.fsgs_rgb_loss_backward(3, 8, 8, one, one, None, None, 0, 0.2, None, None, one, None) == -1    # no maps
    assert L.fsgs_rgb_loss_scratch_bytes(3, 1024, 1280) >= 3 * 64 * 40 * 16


def test_no_cpu_fallback():
    from fsgs_b200 import GaussianRasterizationSettings, GaussianRasterizer
    rs = GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.ones(3), 1.0, torch.eye(4)[None], torch.eye(4)[None], 0,
                                       torch.zeros(3), False, False)
    z = lambda *s: torch.zeros(*s)
    with pytest.raises(_lib.FsgsError, match="no CPU fallback"):
        GaussianRasterizer(rs)(means3D=z(2, 3), means2D=z(2, 3), opacities=z(2, 1), colors_precomp=z(2, 3),
                               scales=z(2, 3), rotations=z(2, 4))


def test_drop_in_module_names():
    import diff_gaussian_rasterization as dgr
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401
    from simple_knn._C import distCUDA2
    assert GaussianRasterizationSettings._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg",
                                                     "scale_modifier", "viewmatrix", "projmatrix", "sh_degree", "campos",
                                                     "prefiltered", "debug")
    assert {"rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"} <= set(dir(dgr._C))
    pts = torch.randn(500, 3, generator=torch.Generator().manual_seed(0))
    d2 = torch.cdist(pts.double(), pts.double()) ** 2
    d2.fill_diagonal_(float("inf"))
    ref = d2.topk(3, largest=False).values.mean(1).float()
    assert torch.allclose(distCUDA2(pts), ref, rtol=1e-5, atol=1e-7)
