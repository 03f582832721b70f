"""Drop-in ``diff_gaussian_rasterization`` API on top of the sm_100a C-ABI library.

Mirrors the third-party package Free-SurGS imports (``requirements.txt:26``; call sites
``gaussian_renderer/__init__.py:15,68,69,131``, ``scene/pose_optimizer.py:5,619-632``,
``scene/gaussian_model.py:18``): same class names, same 12-field settings tuple, same keyword
call convention, same 3-tuple return ``(color[3,H,W], radii[P] int32, depth[1,H,W])``, same error
messages, same gradient slots.  All arithmetic runs in ``libfsgs_raster.so``; there is no
PyTorch or CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# process-wide switches used by tests / profiling (never needed in production)
_FLAGS = {"flags": 0}


def set_debug_flags(no_tma: bool = False, no_tile_cull: bool = False,
                    no_optimistic: bool = False, sort_network: bool = False, no_pose_only: bool = False,
                    sort_window_large: bool = False, no_bins: bool = False, upstream_style: bool = False) -> None:
    _FLAGS["flags"] = ((_lib.FLAG_NO_TMA if no_tma else 0) | (_lib.FLAG_NO_TILE_CULL if no_tile_cull else 0)
                       | (_lib.FLAG_NO_OPTIMISTIC if no_optimistic else 0)
                       | (_lib.FLAG_SORT_NETWORK if sort_network else 0)
                       | (_lib.FLAG_NO_POSE_ONLY if no_pose_only else 0)
                       | (_lib.FLAG_SORT_WINDOW_LARGE if sort_window_large else 0)
                       | (_lib.FLAG_NO_BINS if no_bins else 0)
                       # baseline for bench.py (GaussianRasterizer path only): the published rasteriser's compositor
                       # structure on the reference's full 3-sigma rectangles
                       | ((_lib.FLAG_UPSTREAM_STYLE | _lib.FLAG_NO_TILE_CULL) if upstream_style else 0))


class _WorkspacePool:
    """Grow-only cache of the library's scratch buffers (geometry / binning / image state / gradient
    accumulator), keyed by (device, stream, kind).  Leased at forward, handed back when the autograd
    node dies (after backward) -- so the ~300 MB of per-frame scratch never goes through the framework
    allocator's split/merge machinery, which otherwise fragments under the mixed lifetimes and ends
    up in a cudaMalloc per frame.  Reuse is stream-ordered: a buffer is only handed to work enqueued
    on the stream that used it last."""
    MAX_PER_KEY = 6

    def __init__(self):
        import threading
        self._lock = threading.RLock()     # re-entrant: a lease may be garbage-collected inside acquire()
        self._free = {}

    def acquire(self, key, nbytes: int, device, headroom: float = 1.0) -> torch.Tensor:
        if torch.cuda.is_current_stream_capturing():
            # CUDA-graph capture: the buffer must belong to the graph (its private memory pool keeps it alive
            # and at a fixed address for every replay); never a pooled tensor another call could be handed
            return torch.empty(max(int(nbytes * headroom), 256), dtype=torch.uint8, device=device)
        with self._lock:
            lst = self._free.get(key)
            if lst:
                best = None
                for i, t in enumerate(lst):
                    if t.numel() >= nbytes and (best is None or t.numel() < lst[best].numel()):
                        best = i
                if best is not None:
                    return lst.pop(best)
        want = max(int(nbytes * headroom), 256)
        want = (want + (1 << 20) - 1) >> 20 << 20 if want > (1 << 20) else want
        return torch.empty(want, dtype=torch.uint8, device=device)

    def release(self, key, t: torch.Tensor) -> None:
        with self._lock:
            lst = self._free.setdefault(key, [])
            lst.append(t)
            if len(lst) > self.MAX_PER_KEY:
                lst.sort(key=lambda x: x.numel())
                lst.pop(0)

    def clear(self) -> None:
        with self._lock:
            self._free.clear()

    def drop_stream(self, device_index, stream_handle) -> None:
        """Forget the buffers cached for one (device, stream): for streams that will not be used again."""
        with self._lock:
            for key in [k for k in self._free if k[0] == device_index and k[1] == stream_handle]:
                del self._free[key]


_POOL = _WorkspacePool()


class _Lease:
    """Buffers checked out of the pool for one forward; returned when this object is collected."""

    def __init__(self):
        self.items = []
        self.pooled = not torch.cuda.is_current_stream_capturing()   # graph-owned buffers never enter the pool

    def add(self, key, t):
        self.items.append((key, t))

    def release(self):
        items, self.items = self.items, []
        if not self.pooled:
            return
        for key, t in items:
            _POOL.release(key, t)

    def __del__(self):
        try:
            self.release()
        except Exception:  # interpreter shutdown
            pass


class _Arena:
    """Serves the library's allocation callbacks from the workspace pool."""
    HEADROOM = {"binning": 1.25}

    def __init__(self, device, stream: Optional[int] = None):
        self.device = device
        self.stream = torch.cuda.current_stream(device).cuda_stream if stream is None else stream
        self.tensors = {}
        self.lease = _Lease()
        self._cbs = []

    def key(self, name):
        return (self.device.index, self.stream, name)

    def take(self, name: str, nbytes: int) -> torch.Tensor:
        t = _POOL.acquire(self.key(name), int(nbytes), self.device, self.HEADROOM.get(name, 1.0))
        self.lease.add(self.key(name), t)
        self.tensors[name] = t
        return t

    def callback(self, name: str):
        def cb(_user, nbytes):
            return self.take(name, nbytes).data_ptr()
        fn = _lib.ALLOC_FN(cb)
        self._cbs.append(fn)
        return fn

    def finish(self) -> "_Lease":
        """Call after the library returns: drops the ctypes callbacks (they close over ``self`` -- a
        reference cycle that would otherwise keep the lease alive until the cyclic GC runs) and hands
        out the lease, whose lifetime alone decides when the buffers go back to the pool."""
        self._cbs.clear()
        lease, self.lease = self.lease, None
        return lease


def _f32(t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    if t is None or t.numel() == 0:
        return None
    if t.device != device:
        raise ValueError(f"tensor on {t.device}, expected {device}")
    if t.dtype == torch.float32 and t.is_contiguous():      # the usual case: no copies, one cheap op at most
        return t.detach() if t.requires_grad else t
    return t.detach().contiguous().float()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _require_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.FsgsError("fsgs_b200 runs on a CUDA (sm_100a) device only: there is no CPU fallback")


def make_settings(rs, n_coeffs: int = 0, sh_degree: Optional[int] = None) -> _lib.Settings:
    return _lib.Settings(int(rs.image_height), int(rs.image_width), float(rs.tanfovx), float(rs.tanfovy),
                         float(rs.scale_modifier), int(rs.sh_degree if sh_degree is None else sh_degree),
                         int(n_coeffs), 1 if rs.debug else 0, _FLAGS["flags"])


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _on_device:
    """``with torch.cuda.device(dev)`` only when ``dev`` is not already current (the context manager costs
    several microseconds of host time per use, and the hot path enters it three times per frame)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if torch.cuda.current_device() == device.index else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


# ------------------------------------------------------------------------------------------------
# `_C`-level functions (what the reference's pybind module exposes)
# ------------------------------------------------------------------------------------------------
def rasterize_gaussians(bg, means3D, colors_precomp, opacities, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tanfovx, tanfovy, image_height, image_width, sh, degree, campos,
                        prefiltered, debug):
    """-> (num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer)"""
    _require_cuda(means3D)
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    dev = means3D.device
    P = means3D.shape[0]
    H, W = int(image_height), int(image_width)
    sh_t = _f32(sh, dev)
    n_coeffs = 0 if sh_t is None else sh_t.shape[1]
    st = _lib.Settings(H, W, float(tanfovx), float(tanfovy), float(scale_modifier), int(degree), int(n_coeffs),
                       1 if debug else 0, _FLAGS["flags"])
    t = dict(bg=_f32(bg, dev), means3D=_f32(means3D, dev), colors=_f32(colors_precomp, dev), sh=sh_t,
             opac=_f32(opacities, dev), scales=_f32(scales, dev), rots=_f32(rotations, dev),
             cov=_f32(cov3D_precomp, dev), view=_f32(viewmatrix, dev), proj=_f32(projmatrix, dev),
             campos=_f32(campos, dev))
    color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
    depth = torch.empty(1, H, W, dtype=torch.float32, device=dev)
    radii = torch.empty(P, dtype=torch.int32, device=dev)       # every entry is written by the projection kernel
    arena = _Arena(dev)
    nr, nrect = ctypes.c_int64(0), ctypes.c_int64(0)
    with _on_device(dev):
        rc = _lib.lib().fsgs_rasterize_forward(
            ctypes.byref(st), P, _ptr(t["bg"]), _ptr(t["means3D"]), _ptr(t["colors"]), _ptr(t["sh"]), _ptr(t["opac"]),
            _ptr(t["scales"]), _ptr(t["rots"]), _ptr(t["cov"]), _ptr(t["view"]), _ptr(t["proj"]), _ptr(t["campos"]),
            arena.callback("geom"), None, arena.callback("binning"), None, arena.callback("img"), None,
            _ptr(color), _ptr(depth), _ptr(radii), ctypes.byref(nr), ctypes.byref(nrect), _stream(dev))
    _lib.check(rc)
    empty = torch.empty(0, dtype=torch.uint8, device=dev)
    # The caller gets ALIASES of the pooled buffers, and the aliases carry the lease: when the caller drops all
    # three, the lease dies (plain reference counting) and the pooled tensors go back to the workspace pool.
    # (Tagging the pooled tensors themselves would close a cycle tensor -> lease -> tensor that only the cyclic
    # GC can break -- the buffers then come back late and every frame allocates ~400 MB of fresh scratch.)
    bufs = [arena.tensors[k][:] if k in arena.tensors else empty for k in ("geom", "binning", "img")]
    _CALL_TLS.last_num_rect = int(nrect.value)        # per thread: the viewer thread rasterises too (train.py:124-152)
    lease = arena.finish()
    for b in bufs:
        b._fsgs_lease = lease
    return int(nr.value), color, depth, radii, bufs[0], bufs[1], bufs[2]


import threading as _threading

_CALL_TLS = _threading.local()


def rasterize_gaussians_backward(bg, means3D, radii, colors_precomp, scales, rotations, scale_modifier, cov3D_precomp,
                                 viewmatrix, projmatrix, tanfovx, tanfovy, grad_out_color, grad_out_depth, sh, degree,
                                 campos, geomBuffer, num_rendered, binningBuffer, imgBuffer, debug, opacities=None,
                                 image_height=None, image_width=None):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)"""
    _require_cuda(means3D)
    dev = means3D.device
    P = means3D.shape[0]
    H = int(grad_out_color.shape[1]) if image_height is None else int(image_height)
    W = int(grad_out_color.shape[2]) if image_width is None else int(image_width)
    sh_t = _f32(sh, dev)
    n_coeffs = 0 if sh_t is None else sh_t.shape[1]
    st = _lib.Settings(H, W, float(tanfovx), float(tanfovy), float(scale_modifier), int(degree), int(n_coeffs),
                       1 if debug else 0, _FLAGS["flags"])
    z = (lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)) if P > 0 else \
        (lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev))     # the library overwrites every output
    g = dict(means2D=z(P, 3), colors=z(P, 3), opacity=z(P, 1), means3D=z(P, 3), cov3D=z(P, 6),
             sh=z(P, max(n_coeffs, 0), 3), scales=z(P, 3), rots=z(P, 4))
    if P == 0:
        return tuple(g[k] for k in ("means2D", "colors", "opacity", "means3D", "cov3D", "sh", "scales", "rots"))
    t = dict(bg=_f32(bg, dev), means3D=_f32(means3D, dev), colors=_f32(colors_precomp, dev), sh=sh_t,
             opac=_f32(opacities, dev), scales=_f32(scales, dev), rots=_f32(rotations, dev),
             cov=_f32(cov3D_precomp, dev), view=_f32(viewmatrix, dev), proj=_f32(projmatrix, dev),
             campos=_f32(campos, dev), gc=_f32(grad_out_color, dev), gd=_f32(grad_out_depth, dev))
    arena = _Arena(dev)
    scratch = arena.take("grad_scratch", _lib.lib().fsgs_grad_scratch_bytes(P))
    with _on_device(dev):
        rc = _lib.lib().fsgs_rasterize_backward(
            ctypes.byref(st), P, int(num_rendered), _ptr(t["bg"]), _ptr(t["means3D"]), _ptr(t["colors"]), _ptr(t["sh"]),
            _ptr(t["opac"]), _ptr(t["scales"]), _ptr(t["rots"]), _ptr(t["cov"]), _ptr(t["view"]), _ptr(t["proj"]),
            _ptr(t["campos"]), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer), _ptr(t["gc"]), _ptr(t["gd"]),
            _ptr(scratch), _ptr(g["means2D"]), _ptr(g["colors"]), _ptr(g["opacity"]), _ptr(g["means3D"]),
            _ptr(g["cov3D"]), _ptr(g["sh"]) if n_coeffs > 0 else None, _ptr(g["scales"]), _ptr(g["rots"]), _stream(dev))
    arena.finish().release()              # kernels are enqueued; reuse is ordered on this stream
    _lib.check(rc)
    return tuple(g[k] for k in ("means2D", "colors", "opacity", "means3D", "cov3D", "sh", "scales", "rots"))


def mark_visible(means3D, viewmatrix, projmatrix):
    _require_cuda(means3D)
    dev = means3D.device
    P = means3D.shape[0]
    vis = torch.zeros(P, dtype=torch.uint8, device=dev)
    m, v, p = _f32(means3D, dev), _f32(viewmatrix, dev), _f32(projmatrix, dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().fsgs_mark_visible(P, _ptr(m), _ptr(v), _ptr(p), _ptr(vis), _stream(dev))
    _lib.check(rc)
    return vis.bool()


# ------------------------------------------------------------------------------------------------
# autograd function + module (the Python layer of the reference package)
# ------------------------------------------------------------------------------------------------
class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        rs = raster_settings
        num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = rasterize_gaussians(
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree,
            rs.campos, rs.prefiltered, rs.debug)
        ctx.raster_settings = rs
        ctx.lease = getattr(geomBuffer, "_fsgs_lease", None)
        ctx.num_rendered = num_rendered
        ctx.num_rect = getattr(_CALL_TLS, "last_num_rect", 0)
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, opacities,
                              geomBuffer, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, grad_out_depth):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, opacities, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        if grad_out_color is None:
            grad_out_color = torch.zeros(3, rs.image_height, rs.image_width, device=means3D.device)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = rasterize_gaussians_backward(
            rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix,
            rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, grad_out_depth, sh, rs.sh_degree, rs.campos,
            geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, rs.debug, opacities=opacities,
            image_height=rs.image_height, image_width=rs.image_width)
        none_if_empty = lambda g, src: g if src.numel() > 0 else None
        return (grad_means3D, grad_means2D, none_if_empty(grad_sh, sh), none_if_empty(grad_colors_precomp, colors_precomp),
                grad_opacities, none_if_empty(grad_scales, scales), none_if_empty(grad_rotations, rotations),
                none_if_empty(grad_cov3Ds_precomp, cov3Ds_precomp), None)


def rasterize_gaussians_autograd(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                 raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([]).to(means3D.device)
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        return rasterize_gaussians_autograd(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                            cov3D_precomp, rs)
