"""fsgs_b200 -- B200-native (sm_100a) differentiable Gaussian rasteriser for Free-SurGS.

Public surface (mirrors the reference's names):
  * ``GaussianRasterizationSettings``, ``GaussianRasterizer``   (``diff_gaussian_rasterization`` API)
  * ``render(viewpoint_camera, index, pc, gs_grad, cam_grad)``   (fused ``gaussian_renderer.render``)
  * ``render_two_pass``                                          (the reference's un-fused formulation)
  * ``GraphedStep``                                              (CUDA-graph capture of a whole render step)
All compute goes through ``libfsgs_raster.so`` (C ABI in ``include/fsgs_raster.h``); there is no
CPU or PyTorch fallback.
"""
from . import _lib
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, mark_visible, rasterize_gaussians,
                         rasterize_gaussians_backward, set_debug_flags)
from .frame_render import render, render_planes, render_two_pass
from .graphs import GraphedStep

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "render", "render_planes", "render_two_pass",
           "GraphedStep", "mark_visible", "rasterize_gaussians", "rasterize_gaussians_backward", "set_debug_flags", "_lib"]
