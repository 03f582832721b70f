"""ORACLE tooling (test infrastructure): generate golden vectors from the REFERENCE's own Python.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden_ref_python.py        # writes tests/golden/ref_python_half.npz

It imports the reference's modules unmodified -- third-party imports that are absent from this
image (kornia, matplotlib, lpips, ...) and the un-vendored rasteriser are replaced by inert stub
modules, and ``.cuda()`` / ``device='cuda'`` are redirected to the CPU -- then calls, on seeded
inputs, exactly the functions that make up the Python half of the hot path:

  utils/sh_utils.py:57            eval_sh                       (deg 0..3)
  scene/pose_optimizer.py:822     LearnPose.forward (q2rot, getWorld2View2)
  scene/pose_optimizer.py:960     transform_to_frame
  scene/pose_optimizer.py:600     PoseModel.setup_camera        (settings tuple fields)
  scene/gaussian_model.py:118-138 activations / get_features
  scene/gaussian_model.py:260     get_depth_and_silhouette
  scene/gaussian_model.py:308     transformed_params2rendervar  (SH -> colors_precomp)
  utils/general_utils.py:204      build_rotation

The outputs pin ``oracle/render_oracle.py`` (tests/test_oracle_golden.py).  The rasteriser core has
no reference implementation on disk, so it cannot be pinned this way ("parity unpinned").
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
from typing import NamedTuple
from unittest import mock

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_python_half.npz")
STUB_ROOTS = {"kornia", "matplotlib", "mpl_toolkits", "lpips", "skimage", "plyfile", "imageio", "torchviz",
              "cv2", "open3d", "viser", "nerfview", "splines", "simple_knn", "wandb", "PIL"}


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return mock.MagicMock(name=f"{self.__name__}.{name}")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in STUB_ROOTS:
            try:
                if root == "PIL":
                    return None  # PIL may exist; prefer the real one
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
            except Exception:
                return None
        return None

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


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


def _install_stubs():
    sys.meta_path.insert(0, _StubFinder())
    dgr = types.ModuleType("diff_gaussian_rasterization")
    dgr.GaussianRasterizationSettings = GaussianRasterizationSettings
    dgr.GaussianRasterizer = mock.MagicMock(name="GaussianRasterizer")
    sys.modules["diff_gaussian_rasterization"] = dgr
    sys.path.insert(0, REF)


def _cpu_redirect():
    """Make the reference's hard-coded 'cuda' land on the CPU (this container has no GPU)."""
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    for name in ("zeros", "ones", "eye", "tensor", "zeros_like", "ones_like", "empty", "arange", "full", "randint"):
        orig = getattr(torch, name)

        def wrapped(*a, __orig=orig, **k):
            if "device" in k and str(k["device"]).startswith("cuda"):
                k["device"] = "cpu"
            return __orig(*a, **k)
        setattr(torch, name, wrapped)


def main():
    _install_stubs()
    _cpu_redirect()
    from utils.sh_utils import eval_sh                      # noqa: E402
    from utils.general_utils import build_rotation          # noqa: E402
    import scene.pose_optimizer as po                        # noqa: E402
    import scene.gaussian_model as gm                        # noqa: E402

    g = torch.Generator().manual_seed(20260101)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32)
    P = 257
    out = {}

    # ---- inputs ---------------------------------------------------------------------------
    xyz = rn(P, 3) * 0.5 + torch.tensor([0.0, 0.0, 1.0])
    f_dc, f_rest = 0.5 * rn(P, 1, 3), 0.1 * rn(P, 15, 3)
    opacity, scaling, rotation = 1.5 * rn(P, 1), rn(P, 3) * 0.3 - 4.0, rn(P, 4)
    r_all = torch.tensor([[1.0, 0.03, -0.02, 0.05], [0.9, -0.1, 0.2, 0.05]]).T[None].contiguous()  # [1,4,2]
    t_all = torch.tensor([[0.01, -0.005, 0.02], [-0.03, 0.02, 0.05]]).T.contiguous()               # [3,2]
    out.update(in_xyz=xyz, in_f_dc=f_dc, in_f_rest=f_rest, in_opacity=opacity, in_scaling=scaling,
               in_rotation=rotation, in_r=r_all, in_t=t_all)

    # ---- eval_sh --------------------------------------------------------------------------
    dirs = torch.nn.functional.normalize(rn(P, 3))
    sh = rn(P, 3, 16)
    out.update(in_sh_dirs=dirs, in_sh=sh)
    for deg in range(4):
        out[f"eval_sh_deg{deg}"] = eval_sh(deg, sh, dirs)

    # ---- build_rotation -------------------------------------------------------------------
    out["build_rotation"] = build_rotation(rotation)

    # ---- LearnPose ------------------------------------------------------------------------
    lp = po.LearnPose(2, 512, 640, 1.0, 1.0)
    with torch.no_grad():
        lp.r.copy_(r_all)
        lp.t.copy_(t_all)
    for k in range(2):
        Rt = lp.forward(k)
        out[f"learnpose_Rt{k}"] = Rt.detach()
        # gradient of a fixed linear functional of Rt wrt (r, t): pins the backward of q2rot
        Gm = torch.arange(16, dtype=torch.float32).reshape(4, 4) / 7.0 - 1.0
        lp.zero_grad()
        (Rt * Gm).sum().backward()
        out[f"learnpose_dr{k}"] = lp.r.grad.clone()
        out[f"learnpose_dt{k}"] = lp.t.grad.clone()
    out["learnpose_G"] = Gm

    # ---- transform_to_frame ---------------------------------------------------------------
    Rt0 = out["learnpose_Rt1"]
    out["transform_to_frame"] = po.transform_to_frame(xyz, Rt0)

    # ---- setup_camera (unbound, fake self) --------------------------------------------------
    K = np.array([[1035.0 * 640 / 1280, 0, 320.0], [0, 1035.0 * 512 / 1024 * 1.01, 250.0], [0, 0, 1]])
    fake = types.SimpleNamespace(record_data={"intrinsic": K, "image_width": 640, "image_height": 512})
    for name, w2c in (("I", np.eye(4)), ("Rt", Rt0.numpy().astype(np.float64))):
        cam = po.PoseModel.setup_camera(fake, w2c)
        out[f"cam_{name}_scalars"] = torch.tensor([cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                                                   cam.scale_modifier, cam.sh_degree], dtype=torch.float64)
        out[f"cam_{name}_viewmatrix"] = cam.viewmatrix
        out[f"cam_{name}_projmatrix"] = cam.projmatrix
        out[f"cam_{name}_campos"] = cam.campos
        out[f"cam_{name}_bg"] = cam.bg
        out[f"cam_{name}_w2c"] = torch.tensor(w2c)
    out["cam_K"] = torch.tensor(K)

    # ---- GaussianModel pieces ---------------------------------------------------------------
    gmodel = gm.GaussianModel(3, None)
    gmodel.params = {"_xyz": xyz, "_features_dc": f_dc, "_features_rest": f_rest, "_opacity": opacity,
                     "_scaling": scaling, "_rotation": rotation}
    gmodel.active_sh_degree = 3
    out["get_opacity"] = gmodel.get_opacity
    out["get_scaling"] = gmodel.get_scaling
    out["get_rotation"] = gmodel.get_rotation
    out["get_features"] = gmodel.get_features
    means_cam = out["transform_to_frame"]
    cam_I = po.PoseModel.setup_camera(fake, np.eye(4))
    out["depth_and_silhouette"] = gmodel.get_depth_and_silhouette(means_cam, cam_I.viewmatrix)
    for deg in (0, 1, 2, 3):
        gmodel.active_sh_degree = deg
        rv = gmodel.transformed_params2rendervar(means_cam, gmodel.get_scaling, gmodel.get_rotation,
                                                 gmodel.get_features, gmodel.get_opacity,
                                                 torch.zeros_like(xyz), camera_center=fake.cam_center)
        out[f"rendervar_colors_deg{deg}"] = rv["colors_precomp"]
    out["cam_center_used"] = fake.cam_center

    # ---- losses the mapping / tracking steps apply to the render (utils/loss_utils.py:40-127) -------------
    from utils import loss_utils as lu                       # noqa: E402
    gl = torch.Generator().manual_seed(77)
    img = torch.rand(3, 256, 384, generator=gl)
    gt = (img + 0.1 * torch.randn(3, 256, 384, generator=gl)).clamp(0, 1)
    msk = (torch.rand(1, 256, 384, generator=gl) > 0.2)
    dep_a = torch.rand(256, 384, generator=gl) + 0.5
    dep_b = dep_a * 1.7 + 0.05 * torch.randn(256, 384, generator=gl)
    out.update(loss_img=img, loss_gt=gt, loss_mask=msk, loss_dep_a=dep_a, loss_dep_b=dep_b)
    out["loss_l1"] = lu.l1_loss(img, gt)
    out["loss_ssim"] = lu.ssim(img, gt)
    out["loss_rgb"] = lu.rgb_loss_func(img, gt)
    out["loss_rgb_masked"] = lu.rgb_loss_func(img, gt, mask=msk)
    out["loss_pearson"] = lu.pearson_depth_loss(dep_a, dep_b)
    torch.manual_seed(1234)
    out["loss_local_pearson"] = lu.local_pearson_loss(dep_a, dep_b, 128, 0.5)
    out["loss_local_pearson_seed"] = torch.tensor(1234)
    # gradient of the mapping-style combination w.r.t. the rendered image / depth
    img_g = img.clone().requires_grad_(True)
    dep_g = dep_b.clone().requires_grad_(True)
    torch.manual_seed(1234)
    total = lu.rgb_loss_func(img_g, gt) * 5.0 + lu.pearson_depth_loss(dep_a, dep_g) * 0.05 + \
        lu.local_pearson_loss(dep_a, dep_g, 128, 0.5) * 0.15
    total.backward()
    out["loss_total"] = total.detach()
    out["loss_dimg"] = img_g.grad[:, ::8, ::8].clone()
    out["loss_ddep"] = dep_g.grad[::8, ::8].clone()

    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **arrays)
    print("wrote", os.path.normpath(OUT), f"{os.path.getsize(OUT) / 1024:.1f} KiB", len(arrays), "arrays")


if __name__ == "__main__":
    main()
