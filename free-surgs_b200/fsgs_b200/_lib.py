"""ctypes binding of ``libfsgs_raster.so`` (the C ABI declared in ``include/fsgs_raster.h``).

There is NO fallback: if the shared library is missing, or the device is not a B200-class
(sm_100) GPU, every entry point raises.  PyTorch is only used by the callers for device memory
and streams; nothing torch-typed crosses this boundary.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSGS_RASTER_LIB") or os.path.join(_HERE, "libfsgs_raster.so")   # override: A/B builds
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")

FLAG_NO_TMA = 1
FLAG_NO_TILE_CULL = 2
FLAG_NO_OPTIMISTIC = 8
FLAG_SORT_NETWORK = 16
FLAG_FIXED_CAPACITY = 32
FLAG_NO_POSE_ONLY = 64
FLAG_SORT_WINDOW_LARGE = 128
FLAG_NO_BINS = 256
FLAG_UPSTREAM_STYLE = 512


class FsgsError(RuntimeError):
    pass


class Settings(ctypes.Structure):
    _fields_ = [("image_height", ctypes.c_int32), ("image_width", ctypes.c_int32),
                ("tanfovx", ctypes.c_float), ("tanfovy", ctypes.c_float),
                ("scale_modifier", ctypes.c_float), ("sh_degree", ctypes.c_int32),
                ("n_coeffs", ctypes.c_int32), ("debug", ctypes.c_int32), ("flags", ctypes.c_int32)]


class BackwardOpts(ctypes.Structure):
    """fsgs_backward_opts (include/fsgs_raster.h)."""
    _fields_ = [("xyz_gradient_accum", ctypes.c_void_p), ("denom", ctypes.c_void_p), ("compact", ctypes.c_void_p),
                ("first", ctypes.c_int32), ("count", ctypes.c_int32), ("skip_composite", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


class RenderExtras(ctypes.Structure):
    """fsgs_render_extras: optional derived outputs of render() (device pointers, NULL = skip)."""
    _fields_ = [("uncertainty", ctypes.c_void_p), ("presence_mask", ctypes.c_void_p), ("nan_mask", ctypes.c_void_p),
                ("visibility", ctypes.c_void_p), ("max_radii2D", ctypes.c_void_p)]


ALLOC_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

_vp, _i32, _i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
_SIGNATURES = {
    # name: (restype, argtypes)
    "fsgs_abi_version": (ctypes.c_int, []),
    "fsgs_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "fsgs_kernel_names": (ctypes.c_char_p, []),
    "fsgs_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "fsgs_profile_collect": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i64), ctypes.c_int]),
    "fsgs_geom_bytes": (ctypes.c_size_t, [_i32]),
    "fsgs_img_bytes": (ctypes.c_size_t, [_i32, _i32]),
    "fsgs_binning_bytes": (ctypes.c_size_t, [_i64]),
    "fsgs_grad_scratch_bytes": (ctypes.c_size_t, [_i32]),
    "fsgs_geom_record_offset": (ctypes.c_size_t, [_i32]),
    "fsgs_img_offsets": (None, [_i32, _i32, ctypes.POINTER(ctypes.c_size_t)]),
    "fsgs_binning_offsets": (None, [_i64, ctypes.POINTER(ctypes.c_size_t)]),
    "fsgs_rasterize_forward": (ctypes.c_int, [ctypes.POINTER(Settings), _i32] + [_vp] * 11 +
                               [ALLOC_FN, _vp, ALLOC_FN, _vp, ALLOC_FN, _vp] + [_vp] * 3 +
                               [ctypes.POINTER(_i64), ctypes.POINTER(_i64), _vp]),
    "fsgs_rasterize_backward": (ctypes.c_int, [ctypes.POINTER(Settings), _i32, _i64] + [_vp] * 26),
    "fsgs_mark_visible": (ctypes.c_int, [_i32, _vp, _vp, _vp, _vp, _vp]),
    "fsgs_pose_forward": (ctypes.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "fsgs_pose_backward": (ctypes.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "fsgs_render_forward": (ctypes.c_int, [ctypes.POINTER(Settings), _i32] + [_vp] * 11 +
                            [ALLOC_FN, _vp, ALLOC_FN, _vp, ALLOC_FN, _vp] + [_vp] * 2 +
                            [ctypes.POINTER(_i64), ctypes.POINTER(_i64), _vp]),
    "fsgs_render_forward_ex": (ctypes.c_int, [ctypes.POINTER(Settings), _i32] + [_vp] * 11 +
                               [ALLOC_FN, _vp, ALLOC_FN, _vp, ALLOC_FN, _vp] + [_vp] * 2 +
                               [ctypes.POINTER(_i64), ctypes.POINTER(_i64), _vp, _vp]),
    "fsgs_render_backward": (ctypes.c_int, [ctypes.POINTER(Settings), _i32, _i64] + [_vp] * 16 +
                             [_i32, _i32] + [_vp] * 9),
    "fsgs_render_backward_ex": (ctypes.c_int, [ctypes.POINTER(Settings), _i32, _i64] + [_vp] * 19 +
                                [_i32, _i32] + [_vp] * 10),
    "fsgs_render_backward_v2": (ctypes.c_int, [ctypes.POINTER(Settings), _i32, _i64] + [_vp] * 19 +
                                [_i32, _i32] + [_vp] * 9 + [ctypes.POINTER(BackwardOpts), _vp]),
    "fsgs_compact_grad_expand": (ctypes.c_int, [ctypes.POINTER(Settings), _i32, _i32, _i32] + [_vp] * 10),
    "fsgs_compact_grad_expand_peers": (ctypes.c_int, [ctypes.POINTER(Settings), _i32, _i32, _i32, _vp, _vp,
                                                      ctypes.POINTER(_vp), _i32, _i32, _i64, _i64] + [_vp] * 7),
    "fsgs_exchange_rows_scatter": (ctypes.c_int, [_vp, ctypes.POINTER(_vp), _i32, _i32, _i64, _i64, _i64, _vp]),
    "fsgs_exchange_rows": (ctypes.c_int, [_vp, ctypes.POINTER(_vp), _i32, _i32, _i64, _i64, _vp]),
    "fsgs_set_instance_capacity": (ctypes.c_int, [_i32, _i64]),
    "fsgs_fixed_bin_capacity": (_i64, [_i32]),
    "fsgs_frozen_bytes": (ctypes.c_size_t, [_i32]),
    "fsgs_freeze_model": (ctypes.c_int, [ctypes.POINTER(Settings), _i32] + [_vp] * 9),
    "fsgs_render_forward_frozen": (ctypes.c_int, [ctypes.POINTER(Settings), _i32] + [_vp] * 5 +
                                   [ALLOC_FN, _vp, ALLOC_FN, _vp, ALLOC_FN, _vp] + [_vp] * 2 +
                                   [ctypes.POINTER(_i64), ctypes.POINTER(_i64), _vp, _vp]),
    "fsgs_watchdog_flag": (ctypes.c_int, [_i32, _i32]),
    "fsgs_sh_grad_expand": (ctypes.c_int, [ctypes.POINTER(Settings), _i32] + [_vp] * 6),
    "fsgs_rgb_loss_scratch_bytes": (ctypes.c_size_t, [_i32, _i32, _i32]),
    "fsgs_rgb_loss_forward": (ctypes.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, ctypes.c_float, _vp, _vp, _vp, _vp]),
    "fsgs_pearson_scratch_bytes": (ctypes.c_size_t, []),
    "fsgs_pearson_forward": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fsgs_pearson_backward": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fsgs_local_pearson_scratch_bytes": (ctypes.c_size_t, [_i32]),
    "fsgs_local_pearson_forward": (ctypes.c_int, [_i32, _i32, _i32, _i32] + [_vp] * 8),
    "fsgs_local_pearson_backward": (ctypes.c_int, [_i32, _i32, _i32, _i32] + [_vp] * 9),
    "fsgs_rgb_loss_backward": (ctypes.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, ctypes.c_float, _vp, _vp, _vp, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lock = threading.Lock()
_lib = None


def build(force: bool = False, extra: str = "") -> str:
    """Compile the library in-tree with nvcc for sm_100a (works without a GPU)."""
    cmd = ["make", "-C", CSRC]
    if force:
        cmd.append("-B")
    if extra:
        cmd.append(f"EXTRA={extra}")
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """Load (once) and type the shared library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise FsgsError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                                    f"g.build()'` (or `make -C {CSRC}`); there is no CPU / PyTorch fallback")
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(L, name)
                    fn.restype, fn.argtypes = res, args
                if L.fsgs_abi_version() != 1:
                    raise FsgsError("libfsgs_raster.so ABI version mismatch; rebuild")
                _lib = L
    return _lib


def check(code: int) -> None:
    if code != 0:
        msg = lib().fsgs_error_string(code)
        raise FsgsError(f"fsgs_raster error {code}: {msg.decode() if msg else '?'}")


def img_offsets(W: int, H: int):
    out = (ctypes.c_size_t * 6)()
    lib().fsgs_img_offsets(W, H, out)
    return dict(zip(("final_T", "n_contrib", "tile_count", "tile_offset", "cursor", "counters"), map(int, out)))


def binning_offsets(R: int):
    out = (ctypes.c_size_t * 2)()
    lib().fsgs_binning_offsets(R, out)
    return {"keys": int(out[0]), "records": int(out[1])}


def profile_enable(on: bool) -> None:
    check(lib().fsgs_profile_enable(1 if on else 0))


def profile_collect():
    """-> {kernel name: (total ms, launches)} since profile_enable(True)."""
    names = kernel_names()
    ms = (ctypes.c_double * len(names))()
    cnt = (_i64 * len(names))()
    check(lib().fsgs_profile_collect(ms, cnt, len(names)))
    return {n: (float(ms[i]), int(cnt[i])) for i, n in enumerate(names)}


def kernel_names():
    return lib().fsgs_kernel_names().decode().split(",")
