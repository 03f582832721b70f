"""ORACLE (test infrastructure, never a product path).

ctypes binding + ``torch.autograd.Function`` around ``oracle/raster_oracle.c`` (the plain-C
restatement of the rasteriser; PARITY UNPINNED, see that file's header).  Used by tests as the
scalable truth (float64 build), and by ``bench.py`` (``cpu_baseline`` / ``--impl reference``)
as the timed CPU port (float32 build, OpenMP over all host cores).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force: bool = False) -> None:
    """Compile the two oracle libraries (``make -C oracle``)."""
    out = os.path.join(_HERE, "_build")
    have = all(os.path.exists(os.path.join(out, f"liboracle_{s}.so")) for s in ("f32", "f64"))
    src_newer = have and os.path.getmtime(os.path.join(_HERE, "raster_oracle.c")) > \
        os.path.getmtime(os.path.join(out, "liboracle_f32.so"))
    if force or not have or src_newer:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)


def _lib(dtype: torch.dtype):
    sfx = "f64" if dtype == torch.float64 else "f32"
    if sfx not in _LIBS:
        path = os.path.join(_HERE, "_build", f"liboracle_{sfx}.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        real = ctypes.c_double if sfx == "f64" else ctypes.c_float
        p = ctypes.c_void_p
        fwd = getattr(lib, f"fsgs_oracle_forward_{sfx}")
        fwd.restype = ctypes.c_void_p
        fwd.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, p, p, p, p, p, real, p, p, p, p, p,
                        ctypes.c_int, ctypes.c_int, real, real, p, p, p, p, p, p, p, real]
        bwd = getattr(lib, f"fsgs_oracle_backward_{sfx}")
        bwd.restype = None
        bwd.argtypes = [p] * 11
        fr = getattr(lib, f"fsgs_oracle_free_{sfx}")
        fr.restype = None
        fr.argtypes = [p]
        nt = getattr(lib, f"fsgs_oracle_num_threads_{sfx}")
        nt.restype = ctypes.c_int
        _LIBS[sfx] = (lib, fwd, bwd, fr, nt)
    return _LIBS[sfx]


def num_threads() -> int:
    return int(_lib(torch.float32)[4]())


def set_num_threads(n: int) -> None:
    """Use ``n`` OpenMP threads in both builds (overrides an inherited OMP_NUM_THREADS, e.g. torchrun's 1)."""
    for dt, sfx in ((torch.float32, "f32"), (torch.float64, "f64")):
        lib = _lib(dt)[0]
        fn = getattr(lib, f"fsgs_oracle_set_num_threads_{sfx}")
        fn.restype, fn.argtypes = None, [ctypes.c_int]
        fn(int(n))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class _Handle:
    def __init__(self, ptr, free, keep):
        self.ptr, self._free, self.keep = ptr, free, keep

    def __del__(self):
        if self.ptr:
            self._free(self.ptr)
            self.ptr = None


def forward(means3D, opacities, st, colors_precomp=None, shs=None, scales=None, rotations=None,
            cov3D_precomp=None, want_margin=False, sort_depth=None, margin_kappa=0.0):
    """Raw forward.  ``st`` is anything with the GaussianRasterizationSettings fields.
    Returns (color[3,H,W], radii[P] int32, depth[1,H,W], handle, num_rendered)."""
    dt = means3D.dtype
    _, fwd, _, fr, _ = _lib(dt)
    c = lambda t: None if t is None else t.detach().to(dt).contiguous()
    means3D, opacities, colors_precomp, shs, scales, rotations, cov3D_precomp = map(
        c, (means3D, opacities, colors_precomp, shs, scales, rotations, cov3D_precomp))
    P = means3D.shape[0]
    H, W = int(st.image_height), int(st.image_width)
    view = torch.as_tensor(st.viewmatrix).detach().cpu().to(dt).reshape(16).contiguous()
    proj = torch.as_tensor(st.projmatrix).detach().cpu().to(dt).reshape(16).contiguous()
    campos = torch.as_tensor(st.campos).detach().cpu().to(dt).reshape(3).contiguous()
    bg = torch.as_tensor(st.bg).detach().cpu().to(dt).reshape(3).contiguous()
    color = torch.empty(3, H, W, dtype=dt)
    depth = torch.empty(1, H, W, dtype=dt)
    radii = torch.zeros(max(P, 1), dtype=torch.int32)
    nr = ctypes.c_int64(0)
    margin = torch.empty(2, H, W, dtype=dt) if want_margin else None     # [threshold margin, depth-order margin]
    n_coeffs = 0 if shs is None else shs.shape[1]
    if sort_depth is not None:
        sort_depth = sort_depth.detach().to(torch.float32).contiguous()
        assert sort_depth.numel() == P, "sort_depth must hold one float32 depth per Gaussian"
    ptr = fwd(P, int(st.sh_degree), n_coeffs, _ptr(means3D), _ptr(shs), _ptr(colors_precomp),
              _ptr(opacities), _ptr(scales), float(st.scale_modifier), _ptr(rotations),
              _ptr(cov3D_precomp), _ptr(view), _ptr(proj), _ptr(campos), W, H, float(st.tanfovx),
              float(st.tanfovy), _ptr(bg), _ptr(color), _ptr(depth), _ptr(radii), ctypes.byref(nr),
              _ptr(margin), _ptr(sort_depth), float(margin_kappa))
    keep = (means3D, opacities, colors_precomp, shs, scales, rotations, cov3D_precomp)
    h = _Handle(ptr, fr, keep)
    h.pix_margin = margin
    h.num_rendered = int(nr.value)
    return color, radii[:P], depth, h, int(nr.value)


def backward(handle: _Handle, dL_dcolor, dL_ddepth, P: int, n_coeffs: int, dt):
    _, _, bwd, _, _ = _lib(dt)
    n = max(P, 1)
    g = dict(means2D=torch.zeros(n, 3, dtype=dt), colors=torch.zeros(n, 3, dtype=dt),
             opacity=torch.zeros(n, 1, dtype=dt), means3D=torch.zeros(n, 3, dtype=dt),
             cov3D=torch.zeros(n, 6, dtype=dt), sh=torch.zeros(n, max(n_coeffs, 1), 3, dtype=dt),
             scales=torch.zeros(n, 3, dtype=dt), rots=torch.zeros(n, 4, dtype=dt))
    dc = dL_dcolor.detach().to(dt).contiguous()
    dd = None if dL_ddepth is None else dL_ddepth.detach().to(dt).contiguous()
    bwd(ctypes.c_void_p(handle.ptr), _ptr(dc), _ptr(dd), _ptr(g["means2D"]), _ptr(g["colors"]),
        _ptr(g["opacity"]), _ptr(g["means3D"]), _ptr(g["cov3D"]),
        _ptr(g["sh"]) if n_coeffs > 0 else None, _ptr(g["scales"]), _ptr(g["rots"]))
    return {k: v[:P] for k, v in g.items()}


class _RasterizeC(torch.autograd.Function):
    want_margin = False     # tests flip this to also get the per-pixel threshold margins
    sort_depth = None       # optional float32 depths the sort keys are taken from (see raster_oracle.c)
    margin_kappa = 0.0      # gradient weighting of the threshold margins (see raster_oracle.c)
    last_handle = None

    @staticmethod
    def forward(ctx, means3D, means2D, opacities, colors_precomp, shs, scales, rotations, cov3D_precomp, st):
        color, radii, depth, h, nr = forward(means3D, opacities, st, colors_precomp, shs, scales,
                                             rotations, cov3D_precomp, want_margin=_RasterizeC.want_margin,
                                             sort_depth=_RasterizeC.sort_depth, margin_kappa=_RasterizeC.margin_kappa)
        _RasterizeC.last_handle = h
        ctx.h, ctx.P, ctx.dt = h, means3D.shape[0], means3D.dtype
        ctx.n_coeffs = 0 if shs is None else shs.shape[1]
        ctx.flags = (colors_precomp is not None, shs is not None, scales is not None, cov3D_precomp is not None)
        ctx.mark_non_differentiable(radii)
        ctx.num_rendered = nr
        return color, radii, depth

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth):
        g = backward(ctx.h, g_color, g_depth, ctx.P, ctx.n_coeffs, ctx.dt)
        has_col, has_sh, has_sr, has_cov = ctx.flags
        return (g["means3D"], g["means2D"], g["opacity"], g["colors"] if has_col else None,
                g["sh"] if has_sh else None, g["scales"] if has_sr else None,
                g["rots"] if has_sr else None, g["cov3D"] if has_cov else None, None)


def rasterize(means3D, means2D, opacities, st, colors_precomp=None, shs=None, scales=None,
              rotations=None, cov3D_precomp=None, want_aux=False, sort_depth=None, margin_kappa=0.0):
    """Differentiable ``GaussianRasterizer(st)(...)`` on the CPU through the C oracle.
    Returns (color, radii, depth, aux) like ``raster_oracle.rasterize``; with ``want_aux`` the aux
    dict carries what ``raster_oracle.fragile_pixel_mask`` needs."""
    _RasterizeC.want_margin = bool(want_aux)
    _RasterizeC.sort_depth = sort_depth
    _RasterizeC.margin_kappa = float(margin_kappa)
    try:
        color, radii, depth = _RasterizeC.apply(means3D, means2D, opacities, colors_precomp, shs, scales,
                                                rotations, cov3D_precomp, st)
    finally:
        _RasterizeC.want_margin = False
        _RasterizeC.sort_depth = None
        _RasterizeC.margin_kappa = 0.0
    h = _RasterizeC.last_handle
    aux = {"num_rendered": getattr(h, "num_rendered", None)}
    if want_aux:
        from . import raster_oracle as ro
        dt = means3D.dtype
        with torch.no_grad():
            pre = ro.preprocess(means3D.detach(), torch.zeros_like(means3D), opacities.detach(),
                                None if scales is None else scales.detach(),
                                None if rotations is None else rotations.detach(),
                                None if cov3D_precomp is None else cov3D_precomp.detach(),
                                ro.RasterSettings.from_any(st, dt), colors_precomp=torch.zeros_like(means3D))
        aux.update(pix_margin=h.pix_margin[0].double(), order_margin=h.pix_margin[1].double(), pre=pre)
    return color, radii, depth, aux
