"""CPU emulation of the CUDA kernels' arithmetic (tests/cpu_emul, compiled from the same
host/device header the kernels use) against the float64 oracle.  Covers what can be checked
without a GPU: projection, exact tile culling, SH, the per-pair compositing backward, activation and
pose backward.  The parallel mechanics are covered by the -m gpu tests."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import cpu_emul  # noqa: E402
from parity import check_grad, check_image  # noqa: E402

from fsgs_b200.synth import make_camera, make_scene, pose_matrix  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import raster_oracle as ro  # noqa: E402
from oracle import render_oracle as R  # noqa: E402


@pytest.mark.parametrize("gs_grad,cam_grad,sh_deg,n_grad_planes", [(True, True, 3, 6), (False, True, 2, 3), (True, False, 0, 4)])
def test_fused_render_emulation_matches_oracle(gs_grad, cam_grad, sh_deg, n_grad_planes):
    P, W, H = 900, 168, 120
    sc = make_scene(P, W, H, size_mult=2.0, seed=4)
    dt = torch.float64
    params = {k: v.to(dt).requires_grad_(True) for k, v in sc.params.items()}
    r, t = sc.pose_q.to(dt).requires_grad_(True), sc.pose_t.to(dt).requires_grad_(True)
    out = R.render(params, r, t, sc.camera, sh_deg, sc.camera.campos, gs_grad, cam_grad, want_aux=True)
    G6 = torch.randn(6, H, W, generator=torch.Generator().manual_seed(1))
    G6[n_grad_planes:] = 0            # 3 = pose tracking (colour loss only), 4 = mapping (+ depth), 6 = everything
    ref_planes = torch.cat([out["render"], out["_depth_sil"]], 0)
    out["render_w2c"].retain_grad()
    # viewspace gradient of the reference comes from the RGB pass only
    (ref_planes * G6.to(dt)).sum().backward()
    pose32 = R.learn_pose_forward(sc.pose_q, sc.pose_t)
    results = {}
    for no_cull in (True, False):
        planes, radii, Rn, Rrect, g = cpu_emul.render_fused(sc.params, pose32, sc.camera, sc.camera.campos, sh_deg,
                                                           dplanes=G6, gs_grad=gs_grad, cam_grad=cam_grad, no_cull=no_cull)
        assert Rrect == out["_num_rendered"]
        assert (radii != out["radii"]).sum().item() == 0
        check_image("planes", planes, ref_planes, out["_aux"], scale=4.0)   # depth^2 plane reaches 4
        check_image("rgb", planes[:3], ref_planes[:3], out["_aux"])
        ref = {"xyz": params["_xyz"].grad, "f_dc": params["_features_dc"].grad, "f_rest": params["_features_rest"].grad,
               "opacity": params["_opacity"].grad, "scaling": params["_scaling"].grad, "rotation": params["_rotation"].grad,
               "means2D": out["viewspace_points"].grad}
        for k, v in ref.items():
            if sh_deg == 0 and k == "f_rest":
                assert g[k].abs().max().item() == 0 and v.abs().max().item() == 0
                continue
            check_grad(k, g[k], v)
        if cam_grad:
            check_grad("pose", g["pose"][:3], out["render_w2c"].grad[:3])
            # the pose-only body (tracking against a frozen model) sees only the mean2D / conic / view-depth
            # columns of the accumulator row and must give the same dL/dRt
            check_grad("pose_only", g["pose_only"][:3], out["render_w2c"].grad[:3])
            assert (g["pose_only"] - g["pose"]).abs().max().item() <= 1e-6 * max(1.0, g["pose"].abs().max().item())
        else:
            assert g["pose"].abs().max().item() == 0
        results[no_cull] = (planes, g, Rn)
    # exact tile culling must not change anything, only shrink the instance count
    (p0, g0, R0), (p1, g1, R1) = results[True], results[False]
    assert R1 < R0
    assert torch.equal(p0, p1)
    for k in g0:
        assert (g0[k] - g1[k]).abs().max().item() <= 1e-6 * max(1.0, g0[k].abs().max().item()), k


@pytest.mark.parametrize("mode,sh_deg", [("sh", 3), ("precomp", 0), ("cov", 0)])
def test_api_emulation_matches_oracle(mode, sh_deg):
    P, W, H = 700, 150, 100          # ragged: not multiples of 16
    dt = torch.float64
    sc = make_scene(P, W, H, size_mult=2.0, seed=3)
    cam = make_camera(W, H, pose_matrix((1, 0.05, -0.03, 0.02), (0.02, 0.01, -0.03)))
    cam.bg = torch.tensor([0.2, 0.7, 1.0])
    cam.sh_degree = sh_deg
    xyz = (sc.Rt(dt) @ torch.cat([sc.params["_xyz"].to(dt), torch.ones(P, 1, dtype=dt)], 1).T).T[:, :3]
    xyz = xyz * torch.tensor([1.6, 1.6, 1.0], dtype=dt)     # push some splats into the frustum clamp
    d = dict(means3D=xyz, means2D=torch.zeros(P, 3, dtype=dt), opacities=torch.sigmoid(sc.params["_opacity"].to(dt)))
    if mode == "cov":
        S = ro.build_cov3d(torch.exp(sc.params["_scaling"].to(dt)), torch.nn.functional.normalize(sc.params["_rotation"].to(dt)), 1.0)
        d["cov3D_precomp"] = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1)
    else:
        d["scales"] = torch.exp(sc.params["_scaling"].to(dt))
        d["rotations"] = torch.nn.functional.normalize(sc.params["_rotation"].to(dt)) * 1.1
    if mode == "sh":
        d["shs"] = torch.cat([sc.params["_features_dc"], sc.params["_features_rest"]], 1).to(dt)
    else:
        d["colors_precomp"] = torch.rand(P, 3, generator=torch.Generator().manual_seed(1)).to(dt)
    d = {k: v.detach().clone().requires_grad_(True) for k, v in d.items()}
    gen = torch.Generator().manual_seed(5)
    Gc, Gd = torch.randn(3, H, W, generator=gen), torch.randn(1, H, W, generator=gen)
    color, radii, depth, aux = ro.rasterize(st=cam, want_aux=True, **d)
    ((color * Gc.to(dt)).sum() + (depth * Gd.to(dt)).sum()).backward()
    kw = {k: v for k, v in d.items() if k not in ("means2D",)}
    c2, r2, d2, Rn, Rrect, g = cpu_emul.rasterize_api(cam=cam, dcolor=Gc, ddepth=Gd, **kw)
    assert Rrect == aux["num_rendered"] and Rn <= Rrect
    assert (r2 != radii).sum().item() == 0
    check_image("color", c2, color, aux)
    check_image("depth", d2, depth, aux, scale=2.0)
    names = {"means3D": "means3D", "means2D": "means2D", "opacities": "opacity", "scales": "scales", "rotations": "rots",
             "shs": "sh", "colors_precomp": "colors", "cov3D_precomp": "cov3D"}
    for k, v in d.items():
        check_grad(k, g[names[k]].reshape(v.shape), v.grad)


def test_pose_kernel_math_matches_reference_golden():
    """pose_forward / pose_backward (the bodies of k_pose_forward / k_pose_backward) against the
    reference's own LearnPose outputs and autograd gradients (tests/golden/ref_python_half.npz)."""
    import ctypes
    import numpy as np
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_python_half.npz"))
    L = cpu_emul.lib()
    for k in (0, 1):
        r = torch.from_numpy(G["in_r"])[0, :, k].contiguous()
        t = torch.from_numpy(G["in_t"])[:, k].contiguous()
        dRt = torch.from_numpy(G["learnpose_G"]).contiguous()
        Rt, dr, dt = torch.zeros(16), torch.zeros(4), torch.zeros(3)
        p = lambda x: ctypes.c_void_p(x.data_ptr())
        L.emul_pose(p(r), p(t), p(Rt), p(dRt), p(dr), p(dt))
        assert (Rt.view(4, 4) - torch.from_numpy(G[f"learnpose_Rt{k}"])).abs().max().item() < 2e-7
        assert (dr - torch.from_numpy(G[f"learnpose_dr{k}"])[0, :, k]).abs().max().item() < 2e-5
        assert (dt - torch.from_numpy(G[f"learnpose_dt{k}"])[:, k]).abs().max().item() < 1e-6
