"""TEST INFRASTRUCTURE: builds tests/cpu_emul/emul.cpp (g++) and exposes it through ctypes.
The emulation compiles the kernels' shared arithmetic header for the host; it is not a product
path and the package never imports it."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libemul.so")
_SRC = os.path.join(_HERE, "emul.cpp")
_HDR = os.path.join(_HERE, "..", "..", "free-surgs_b200", "csrc", "fsgs_math.cuh")
_lib = None


def lib():
    global _lib
    if _lib is None:
        stale = (not os.path.exists(_SO)) or any(os.path.getmtime(f) > os.path.getmtime(_SO) for f in (_SRC, _HDR))
        if stale:
            os.makedirs(os.path.dirname(_SO), exist_ok=True)
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", _SO, _SRC],
                           check=True)
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _c(t):
    return None if t is None else t.detach().float().contiguous()


def render_fused(params, pose, cam, cam_center, sh_deg, dplanes=None, gs_grad=True, cam_grad=True, no_cull=False):
    P = params["_xyz"].shape[0]
    H, W = cam.image_height, cam.image_width
    t = [_c(x) for x in (cam.bg, params["_xyz"], params["_features_dc"], params["_features_rest"], params["_opacity"],
                         params["_scaling"], params["_rotation"], pose, cam_center, cam.viewmatrix, cam.projmatrix)]
    planes = torch.zeros(6, H, W)
    radii = torch.zeros(P, dtype=torch.int32)
    R, Rr = ctypes.c_int64(0), ctypes.c_int64(0)
    g = dict(xyz=torch.zeros(P, 3), f_dc=torch.zeros(P, 1, 3), f_rest=torch.zeros(P, 15, 3), opacity=torch.zeros(P, 1),
             scaling=torch.zeros(P, 3), rotation=torch.zeros(P, 4), pose=torch.zeros(4, 4), means2D=torch.zeros(P, 3),
             pose_only=torch.zeros(4, 4))
    dp = _c(dplanes)
    lib().emul_render_fused(
        P, W, H, ctypes.c_float(cam.tanfovx), ctypes.c_float(cam.tanfovy), ctypes.c_float(cam.scale_modifier), int(sh_deg),
        _p(t[0]), _p(t[1]), _p(t[2]), _p(t[3]), _p(t[4]), _p(t[5]), _p(t[6]), _p(t[7]), _p(t[8]), _p(t[9]), _p(t[10]),
        int(no_cull), _p(planes), _p(radii), ctypes.byref(R), ctypes.byref(Rr), _p(dp), int(gs_grad), int(cam_grad),
        _p(g["xyz"]), _p(g["f_dc"]), _p(g["f_rest"]), _p(g["opacity"]), _p(g["scaling"]), _p(g["rotation"]), _p(g["pose"]),
        _p(g["means2D"]), _p(g["pose_only"]))
    return planes, radii, int(R.value), int(Rr.value), g


def rasterize_api(means3D, opacities, cam, colors_precomp=None, shs=None, scales=None, rotations=None,
                  cov3D_precomp=None, dcolor=None, ddepth=None, no_cull=False):
    P = means3D.shape[0]
    H, W = cam.image_height, cam.image_width
    n_coeffs = 0 if shs is None else shs.shape[1]
    t = [_c(x) for x in (cam.bg, means3D, colors_precomp, shs, opacities, scales, rotations, cov3D_precomp,
                         cam.viewmatrix, cam.projmatrix, cam.campos)]
    color, depth = torch.zeros(3, H, W), torch.zeros(1, H, W)
    radii = torch.zeros(P, dtype=torch.int32)
    R, Rr = ctypes.c_int64(0), ctypes.c_int64(0)
    g = dict(means2D=torch.zeros(P, 3), colors=torch.zeros(P, 3), opacity=torch.zeros(P, 1), means3D=torch.zeros(P, 3),
             cov3D=torch.zeros(P, 6), sh=torch.zeros(P, max(n_coeffs, 1), 3), scales=torch.zeros(P, 3),
             rots=torch.zeros(P, 4))
    dc, dd = _c(dcolor), _c(ddepth)
    lib().emul_rasterize_api(
        P, W, H, ctypes.c_float(cam.tanfovx), ctypes.c_float(cam.tanfovy), ctypes.c_float(cam.scale_modifier),
        int(cam.sh_degree), int(n_coeffs), *[_p(x) for x in t], int(no_cull), _p(color), _p(depth), _p(radii),
        ctypes.byref(R), ctypes.byref(Rr), _p(dc), _p(dd), _p(g["means2D"]), _p(g["colors"]), _p(g["opacity"]),
        _p(g["means3D"]), _p(g["cov3D"]), _p(g["sh"]) if n_coeffs else None, _p(g["scales"]), _p(g["rots"]))
    return color, radii, depth, int(R.value), int(Rr.value), g
