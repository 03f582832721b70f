"""Drop-in module name for the rasteriser package Free-SurGS imports
(``from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer``;
reference gaussian_renderer/__init__.py:15, scene/pose_optimizer.py:5, scene/gaussian_model.py:18,
vis/visualizer.py:18-19).  Put ``free-surgs_b200/`` on ``sys.path`` and the reference's
``train.py`` / ``scene.pose_optimizer`` run unchanged on the sm_100a library."""
from fsgs_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                  rasterize_gaussians_autograd as rasterize_gaussians)
from . import _C  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "_C"]
