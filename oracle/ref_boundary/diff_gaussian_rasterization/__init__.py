"""ORACLE-backed ``diff_gaussian_rasterization`` (test infrastructure, never a product path).

The same package surface as ``free-surgs_b200/diff_gaussian_rasterization`` (settings tuple + ``GaussianRasterizer``
module, reference call sites gaussian_renderer/__init__.py:68,69,131), but every call is answered by the plain-C
float32 CPU oracle (oracle/raster_oracle.c): tensors are moved to the host, rasterised there, moved back.  It exists
so that the UNMODIFIED Free-SurGS driver can be run twice on the same synthetic sequence -- once on the CUDA library,
once on this -- and PSNR / ATE compared (config 3, tools/run_config3.py --backend oracle).  Far too slow for
anything but reduced-size sequences.
"""
from __future__ import annotations

import os
import sys
from typing import NamedTuple

import torch
import torch.nn as nn

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import c_oracle  # noqa: E402

N_CALLS = {"forward": 0}


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


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            V = self.raster_settings.viewmatrix.reshape(4, 4).to(positions)
            z = positions @ V[:3, 2] + V[3, 2]
            return z > 0.2

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        dev = means3D.device
        host = lambda t: None if t is None else t.float().cpu()        # differentiable: gradients flow back to `dev`
        N_CALLS["forward"] += 1
        color, radii, depth, _ = c_oracle.rasterize(
            host(means3D), host(means2D), host(opacities), self.raster_settings, colors_precomp=host(colors_precomp),
            shs=host(shs), scales=host(scales), rotations=host(rotations), cov3D_precomp=host(cov3D_precomp))
        return color.to(dev), radii.to(dev), depth.to(dev)
