"""GPU parity tests (run with ``-m gpu`` on the B200 box): the CUDA path, called through the C-ABI
library, against the oracle on the same seeded inputs.

  * small scenes   -> float64 torch.autograd oracle (independent formulation), live
  * config 1       -> plain-C float64 oracle, live (validated against the autograd oracle on CPU)
                      and the committed golden fixture tests/golden/config1_golden.npz
  * full size      -> size-independent properties (determinism, linearity of the backward,
                      sortedness of the tile lists, silhouette == 1, culling / TMA invariance)
Tolerances: <= 1e-5 abs on RGB/depth planes (pixels sitting on a flipped alpha<1/255 or T<1e-4
decision are counted and bounded, see tests/parity.py), <= 1e-4 rel on every gradient.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from parity import (check_grad, check_image, check_mask_fraction, check_radii, fragile_mask, rel_err,  # noqa: E402
                    report)

from fsgs_b200 import _lib  # noqa: E402
from fsgs_b200.synth import make_camera, make_scene, pose_matrix  # noqa: E402
from oracle import raster_oracle as ro  # noqa: E402
from oracle import render_oracle as R  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gpu_modules():
    import fsgs_b200
    from fsgs_b200 import frame_render as render
    from fsgs_b200 import model, rasterizer
    return fsgs_b200, model, rasterizer, render


def _settings_to_cuda(cam, rasterizer):
    return rasterizer.GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=cam.bg.to(DEV), scale_modifier=cam.scale_modifier, viewmatrix=cam.viewmatrix.to(DEV),
        projmatrix=cam.projmatrix.to(DEV), sh_degree=cam.sh_degree, campos=cam.campos.to(DEV), prefiltered=False,
        debug=True)


def _run_fused(sc, G6, gs_grad=True, cam_grad=True, sh_deg=3, which="fused", frozen=False):
    _, model, _, render = _gpu_modules()
    poses, pc = model.scene_to_device(sc, DEV)
    pc.active_sh_degree = sh_deg
    if frozen:
        for v in pc.params.values():
            v.requires_grad_(False)
    pc.cam = pc.cam._replace(debug=True)
    fn = render.render if which == "fused" else render.render_two_pass
    out = fn(poses, 0, pc, gs_grad=gs_grad, cam_grad=cam_grad)
    planes = torch.stack([out["render"][0], out["render"][1], out["render"][2], out["render_dep"],
                          out["render_opacity"], out["uncertainty"][0] + out["render_dep"].detach() ** 2])
    loss = (out["render"] * G6[:3].to(DEV)).sum() + (out["render_dep"] * G6[3].to(DEV)).sum() + \
           (out["render_opacity"] * G6[4].to(DEV)).sum()
    out["render_w2c"].retain_grad()
    loss.backward()
    g = {k: v.grad.detach().cpu() if v.grad is not None else None for k, v in pc.params.items()}
    g["pose"] = None if out["render_w2c"].grad is None else out["render_w2c"].grad.detach().cpu()
    g["r"] = None if poses.pose_param_net.r.grad is None else poses.pose_param_net.r.grad.detach().cpu()
    g["t"] = None if poses.pose_param_net.t.grad is None else poses.pose_param_net.t.grad.detach().cpu()
    g["means2D"] = None if out["viewspace_points"].grad is None else out["viewspace_points"].grad.detach().cpu()
    return out, planes.detach().cpu(), g


def _gpu_sort_depths(sc, sh_deg=3):
    """Float32 view depths of the CUDA path (NaN where it culled the Gaussian): handed to the oracle as its sort
    keys.  The order in which two splats of (nearly) equal depth are composited is decided by float32 rounding of
    the depth -- it differs between ANY two float32 implementations -- so the oracle is told the order the
    implementation under test used; _check_sort_depths verifies these depths against the float64 ones."""
    _, model, _, render = _gpu_modules()
    poses, pc = model.scene_to_device(sc, DEV)
    pc.active_sh_degree = sh_deg
    with torch.no_grad(), render.keep_geometry() as g:
        render.render(poses, 0, pc, gs_grad=False, cam_grad=False)
    rec = g.records[0].cpu()
    vis = rec[:, 10].view(torch.int32) > 0
    return torch.where(vis, rec[:, 9], torch.full_like(rec[:, 9], float("nan")))


def _check_sort_depths(sort_depth, z64):
    """The injected keys must be the true view depths to float32 rounding (<= 4 ulp of the 3-term dot product)."""
    vis = ~torch.isnan(sort_depth)
    rel = ((sort_depth[vis].double() - z64[vis]).abs() / z64[vis].abs()).max().item()
    assert rel <= 4 * 2.0 ** -23, f"GPU view depths are off by {rel:.3g} (relative) from the float64 depths"
    return rel


def _oracle_fused(sc, G6, gs_grad, cam_grad, sh_deg, backend, mask=None, sort_depth=None):
    dt = torch.float64
    params = {k: v.to(dt).requires_grad_(True) for k, v in sc.params.items()}
    r, t = sc.pose_q.to(dt).requires_grad_(True), sc.pose_t.to(dt).requires_grad_(True)
    out = R.render(params, r, t, sc.camera, sh_deg, sc.camera.campos, gs_grad, cam_grad, want_aux=True, backend=backend,
                   sort_depth=sort_depth)
    if sort_depth is not None:
        out["_sort_depth_rel_err"] = _check_sort_depths(sort_depth, out["_means_cam"][:, 2].detach())
    if mask is None:
        mask = fragile_mask(out["_aux"], sc.height, sc.width)
    out["_mask"] = mask
    G6 = G6 * (~mask).float()[None]
    # depth^2 plane is detached in the reference (uncertainty), so it takes no gradient here either
    loss = (out["render"] * G6[:3].to(dt)).sum() + (out["render_dep"] * G6[3].to(dt)).sum() + \
           (out["render_opacity"] * G6[4].to(dt)).sum()
    out["render_w2c"].retain_grad()
    loss.backward()
    planes = torch.cat([out["render"], out["_depth_sil"]], 0).detach()
    g = {k: v.grad for k, v in params.items()}
    g.update(pose=out["render_w2c"].grad, r=r.grad, t=t.grad, means2D=out["viewspace_points"].grad)
    return out, planes, g, G6


def _compare_fused(sc, got_out, got_planes, got_g, ref_out, ref_planes, ref_g, gs_grad, cam_grad, case=None, soft=False):
    """Image planes, radii and every gradient of one fused render against the float64 oracle.  The fragile-pixel
    mask fraction, the number of pixels that actually flipped, the radii mismatches and every gradient's relative
    error are asserted AND reported (tests/parity.py::report)."""
    aux, mask = ref_out["_aux"], ref_out.get("_mask")
    if mask is None:
        mask = fragile_mask(aux, sc.height, sc.width)
    stats = {"P": sc.P, "W": sc.width, "H": sc.height, "mask_fraction": check_mask_fraction("fragile mask", mask),
             "mask_pixels": int(mask.sum())}
    stats["radii_mismatch"], stats["radii_fragile"] = check_radii("radii", got_out["radii"], ref_out["radii"], aux)
    e0, f0, s0 = check_image("rgb", got_planes[:3], ref_planes[:3], aux, mask=mask, soft=soft)
    e1, f1, s1 = check_image("depth_sil", got_planes[3:5], ref_planes[3:5], aux, scale=2.0, mask=mask, soft=soft)
    e2, f2, s2 = check_image("depth_sq", got_planes[5:6], ref_planes[5:6], aux, scale=4.0, mask=mask, soft=soft)
    # *_pixels_above_gate: all pixels off by more than the gate (flipped decisions + float32 rounding);
    # unmasked_*: those of them outside the fragile mask (asserted: none without `soft`, <= 0.02 % within 3x with it)
    stats.update(rgb_max_err=e0, rgb_pixels_above_gate=f0, depth_sil_max_err=e1, depth_sil_pixels_above_gate=f1,
                 depth_sq_max_err=e2, depth_sq_pixels_above_gate=f2, unmasked_pixels_above_gate=[s0, s1, s2])
    if "_sort_depth_rel_err" in ref_out:
        stats["sort_depth_rel_err"] = ref_out["_sort_depth_rel_err"]
    errs = {}
    for k in ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "means2D"):
        if ref_g[k] is None or ref_g[k].abs().max() == 0:
            assert got_g[k] is None or got_g[k].abs().max().item() == 0, k
            continue
        if k == "means2D" and not gs_grad:
            continue      # the reference only retains viewspace_points.grad when gs_grad (__init__.py:57-58)
        errs[k] = check_grad(k, got_g[k].reshape(ref_g[k].shape), ref_g[k])
    if cam_grad:
        errs["dL/dRt"] = check_grad("pose", got_g["pose"][:3], ref_g["pose"][:3])
        errs["dL/dr"] = check_grad("dL/dr", got_g["r"][0, :, 0], ref_g["r"])
        errs["dL/dt"] = check_grad("dL/dt", got_g["t"][:, 0], ref_g["t"])
    stats["grad_rel_err"] = errs
    if case is not None:
        stats["tile_instances"] = [int(x) for x in got_out["num_rendered"]]
        report(case, **stats)
    return stats


# ------------------------------------------------------------------------------------------------
def test_library_loaded_and_arch():
    L = _lib.lib()
    assert L.fsgs_abi_version() == 1
    assert torch.cuda.get_device_capability(0)[0] == 10, "these kernels are sm_100a only"


@pytest.mark.parametrize("gs_grad,cam_grad,sh_deg,which", [(True, True, 3, "fused"), (False, True, 3, "fused"),
                                                          (True, False, 1, "fused"), (True, True, 3, "two_pass")])
def test_fused_render_small_vs_autograd_oracle(gs_grad, cam_grad, sh_deg, which):
    sc = make_scene(1200, 200, 152, size_mult=2.0, seed=4)
    G6 = torch.randn(6, sc.height, sc.width, generator=torch.Generator().manual_seed(1))
    *ref, G6m = _oracle_fused(sc, G6, gs_grad, cam_grad, sh_deg, "py")
    got = _run_fused(sc, G6m, gs_grad, cam_grad, sh_deg, which)
    _compare_fused(sc, *got, *ref, gs_grad, cam_grad)


@pytest.mark.parametrize("n_planes", [3, 4, 5])
def test_pose_only_backward_for_a_frozen_model(n_planes):
    """Pose tracking against a frozen Gaussian model takes the library's pose-only backward (compositor without the
    colour / opacity columns, lean per-Gaussian kernel, no per-Gaussian gradient writes).  dL/dpose, dL/dr, dL/dt
    must match the float64 oracle and the general path; n_planes = 3 / 4 / 5 walks the compositor's gradient
    levels (colour only | + depth | + silhouette)."""
    _, _, rasterizer, _ = _gpu_modules()
    sc = make_scene(1500, 200, 152, size_mult=2.0, seed=6)
    G6 = torch.randn(6, sc.height, sc.width, generator=torch.Generator().manual_seed(3))
    G6[n_planes:] = 0
    *ref, G6m = _oracle_fused(sc, G6, False, True, 3, "py")
    _lib.profile_enable(True)
    try:
        got = _run_fused(sc, G6m, False, True, 3, "fused", frozen=True)
        prof = _lib.profile_collect()
    finally:
        _lib.profile_enable(False)
    assert prof["k_preprocess_pose_bwd"][1] == 1 and prof["k_preprocess_fused_bwd"][1] == 0, "pose-only path not taken"
    assert all(got[2][k] is None for k in ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation"))
    ref_g = ref[2]
    check_grad("pose", got[2]["pose"][:3], ref_g["pose"][:3])
    check_grad("dL/dr", got[2]["r"][0, :, 0], ref_g["r"])
    check_grad("dL/dt", got[2]["t"][:, 0], ref_g["t"])
    assert torch.equal(got[2]["pose"][3], torch.zeros(4))
    # against the general path (trainable model, same upstream gradient): float summation order only
    gen = _run_fused(sc, G6m, False, True, 3, "fused", frozen=False)
    assert torch.equal(gen[1], got[1])
    assert rel_err(got[2]["pose"], gen[2]["pose"]) < 1e-5
    # FSGS_FLAG_NO_POSE_ONLY sends the frozen model down the general kernels too
    try:
        rasterizer.set_debug_flags(no_pose_only=True)
        _lib.profile_enable(True)
        forced = _run_fused(sc, G6m, False, True, 3, "fused", frozen=True)
        prof = _lib.profile_collect()
    finally:
        _lib.profile_enable(False)
        rasterizer.set_debug_flags()
    assert prof["k_preprocess_pose_bwd"][1] == 0 and prof["k_preprocess_fused_bwd"][1] == 1
    assert rel_err(forced[2]["pose"], gen[2]["pose"]) < 1e-5


@pytest.mark.parametrize("P", [1203, 257, 5])
def test_fused_render_ragged_gaussian_counts(P):
    """Gaussian counts that are not multiples of 4 / 32 / 256: the bulk-TMA staging of the SH rows takes its
    plain-load remainder path, the flat gradient buffer carves sub-tensors at odd offsets, tail warps and a
    tail CTA are partially filled."""
    sc = make_scene(P, 120, 88, size_mult=2.0, seed=11)
    G6 = torch.randn(6, sc.height, sc.width, generator=torch.Generator().manual_seed(2))
    *ref, G6m = _oracle_fused(sc, G6, True, True, 3, "py")
    got = _run_fused(sc, G6m, True, True, 3, "fused")
    _compare_fused(sc, *got, *ref, True, True)


@pytest.mark.parametrize("mode,sh_deg", [("sh", 3), ("sh", 1), ("precomp", 0), ("cov", 0)])
def test_api_rasterizer_small_vs_autograd_oracle(mode, sh_deg):
    _, _, rasterizer, _ = _gpu_modules()
    P, W, H = 900, 150, 100          # ragged image: not multiples of 16
    dt = torch.float64
    sc = make_scene(P, W, H, size_mult=2.0, seed=3)
    cam = make_camera(W, H, pose_matrix((1, 0.05, -0.03, 0.02), (0.02, 0.01, -0.03)))
    cam.bg = torch.tensor([0.2, 0.7, 1.0])
    cam.sh_degree = sh_deg
    xyz = (sc.Rt(dt) @ torch.cat([sc.params["_xyz"].to(dt), torch.ones(P, 1, dtype=dt)], 1).T).T[:, :3]
    xyz = xyz * torch.tensor([1.6, 1.6, 1.0], dtype=dt)     # some splats hit the +-1.3 tanfov clamp
    d = dict(means3D=xyz, means2D=torch.zeros(P, 3, dtype=dt), opacities=torch.sigmoid(sc.params["_opacity"].to(dt)))
    if mode == "cov":
        S = ro.build_cov3d(torch.exp(sc.params["_scaling"].to(dt)),
                           torch.nn.functional.normalize(sc.params["_rotation"].to(dt)), 1.0)
        d["cov3D_precomp"] = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1)
    else:
        d["scales"] = torch.exp(sc.params["_scaling"].to(dt))
        d["rotations"] = torch.nn.functional.normalize(sc.params["_rotation"].to(dt)) * 1.1
    if mode == "sh":
        d["shs"] = torch.cat([sc.params["_features_dc"], sc.params["_features_rest"]], 1).to(dt)
    else:
        d["colors_precomp"] = torch.rand(P, 3, generator=torch.Generator().manual_seed(1)).to(dt)
    ref_in = {k: v.detach().clone().requires_grad_(True) for k, v in d.items()}
    gen = torch.Generator().manual_seed(5)
    Gc, Gd = torch.randn(3, H, W, generator=gen), torch.randn(1, H, W, generator=gen)
    color, radii, depth, aux = ro.rasterize(st=cam, want_aux=True, **ref_in)
    keep = (~fragile_mask(aux, H, W)).float()
    Gc, Gd = Gc * keep, Gd * keep
    ((color * Gc.to(dt)).sum() + (depth * Gd.to(dt)).sum()).backward()

    gpu_in = {k: v.detach().float().to(DEV).requires_grad_(True) for k, v in d.items()}
    rs = _settings_to_cuda(cam, rasterizer)
    c2, r2, d2 = rasterizer.GaussianRasterizer(raster_settings=rs)(**gpu_in)
    ((c2 * Gc.to(DEV)).sum() + (d2 * Gd.to(DEV)).sum()).backward()
    assert r2.dtype == torch.int32 and c2.shape == (3, H, W) and d2.shape == (1, H, W)
    assert (r2.cpu() != radii).sum().item() == 0
    check_image("color", c2, color, aux)
    check_image("depth", d2, depth, aux, scale=2.0)
    for k, v in ref_in.items():
        check_grad(k, gpu_in[k].grad.reshape(v.shape), v.grad)


def test_config1_vs_c_oracle_and_golden():
    """BASELINE.json configs[0]: 10k Gaussians, 640x512, m=2, SH degree 3, fwd+bwd incl. dL/d(pose)."""
    sc = make_scene(10000, 640, 512, size_mult=2.0, seed=0)
    G6 = torch.zeros(6, 512, 640)
    G6[:3] = sc.grads_out["G_rgb"]
    G6[3] = sc.grads_out["G_dep"]
    # committed golden fixture (generated in the build container by tests/golden/make_config1_golden.py);
    # its fragile-pixel mask is the one the live oracle must reproduce
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "config1_golden.npz"))
    mask = torch.from_numpy(np.unpackbits(gold["fragile_mask_bits"])[:512 * 640].reshape(512, 640).astype(bool))
    *ref, G6m = _oracle_fused(sc, G6, True, True, 3, "c", mask=mask)
    assert torch.equal(fragile_mask(ref[0]["_aux"], 512, 640), mask)
    got = _run_fused(sc, G6m, True, True, 3, "fused")
    _compare_fused(sc, *got, *ref, True, True, case="config1 P=10000 640x512 m=2 seed0 vs float64 C oracle")
    planes = got[1]
    sub = planes[:, ::4, ::4].double()
    err = (sub - torch.from_numpy(gold["planes_sub4"])).abs()
    assert (err.amax(0) > 1e-5 * 4).double().mean().item() < 2e-3 and err.max().item() < 2e-2
    for k in ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "pose", "means2D"):
        gk = got[2][k] if k != "pose" else got[2]["pose"][:3]
        rk = torch.from_numpy(gold["g_" + k])
        check_grad("golden " + k, gk.reshape(rk.shape), rk)
    assert int(got[0]["num_rendered"][1]) == int(gold["num_rendered_rect"])


def test_tma_and_culling_do_not_change_results():
    _, _, rasterizer, _ = _gpu_modules()
    sc = make_scene(6000, 320, 256, size_mult=2.0, seed=2)
    G6 = torch.randn(6, sc.height, sc.width, generator=torch.Generator().manual_seed(1))
    res = {}
    try:
        for name, kw in (("default", {}), ("no_tma", dict(no_tma=True)), ("no_cull", dict(no_tile_cull=True)),
                         ("no_optimistic", dict(no_optimistic=True)),
                         ("sort_network", dict(sort_network=True)),
                         ("sort_window_large", dict(sort_window_large=True)),
                         ("no_bins", dict(no_bins=True))):
            rasterizer.set_debug_flags(**kw)
            out, planes, g = _run_fused(sc, G6)
            res[name] = (planes, g, tuple(out["num_rendered"]))
    finally:
        rasterizer.set_debug_flags()
    p0, g0, n0 = res["default"]
    assert torch.equal(p0, res["no_tma"][0]), "bulk-TMA staging must be a pure data-movement change"
    assert torch.equal(p0, res["no_cull"][0]), "exact tile culling must not change any pixel"
    assert int(res["no_cull"][2][0]) == int(n0[1]) and int(n0[0]) < int(n0[1])
    assert torch.equal(p0, res["no_optimistic"][0])
    assert torch.equal(p0, res["sort_network"][0]), "bucket sort and compare-exchange network give the same order"
    assert torch.equal(p0, res["sort_window_large"][0])
    assert torch.equal(p0, res["no_bins"][0]), "keys binned by the counting pass vs the scatter pass: same lists"
    for name in ("no_tma", "no_cull", "no_optimistic", "sort_network", "sort_window_large", "no_bins"):
        tol = 2e-6                                        # float summation order (atomics) only
        for k, v in g0.items():
            if v is not None:
                assert rel_err(res[name][1][k], v) < tol, (name, k)


def test_upstream_style_baseline_compositors_agree_with_the_library():
    """bench.py times the published rasteriser's kernel structure (one thread per pixel over the whole list, per-pair
    scalar atomics; FSGS_FLAG_UPSTREAM_STYLE, csrc/fsgs_kernels_refstyle.cuh) beside the library's compositors.  The
    baseline must compute the same thing: identical planes, gradients equal up to summation order."""
    _, _, rasterizer, _ = _gpu_modules()
    sc = make_scene(6000, 320, 256, size_mult=2.0, seed=4)
    G6 = torch.randn(6, sc.height, sc.width, generator=torch.Generator().manual_seed(5))
    res = {}
    try:
        for name, kw in (("library", {}), ("upstream_style", dict(upstream_style=True))):
            rasterizer.set_debug_flags(**kw)
            _lib.profile_enable(True)
            out, planes, g = _run_fused(sc, G6, which="two_pass")
            prof = _lib.profile_collect()
            _lib.profile_enable(False)
            res[name] = (planes, g, prof)
    finally:
        _lib.profile_enable(False)
        rasterizer.set_debug_flags()
    # two passes (a forward compositor may run a third time: an optimistic launch that left on its guard, then the relaunch)
    assert res["library"][2]["k_composite_fwd"][1] >= 2 and res["upstream_style"][2]["k_composite_bwd"][1] == 2
    assert torch.equal(res["library"][0], res["upstream_style"][0])
    for k, v in res["library"][1].items():
        if v is not None:
            assert rel_err(res["upstream_style"][1][k], v) < 2e-5, k


def test_optimistic_tail_relaunches_when_the_instance_count_outgrows_the_hint():
    """The forward launches scatter/sort/composite into a buffer sized from the PREVIOUS frame's instance count
    before the host knows the current one; when the count outgrows it (densification, a new scene) the kernels
    must bail out on the device-side guard and the host must relaunch them into an exact-size buffer."""
    _, _, rasterizer, _ = _gpu_modules()
    small = make_scene(1500, 320, 256, size_mult=1.0, seed=7)
    big = make_scene(30000, 320, 256, size_mult=2.0, seed=8)
    G6 = torch.randn(6, 256, 320, generator=torch.Generator().manual_seed(3))
    try:
        rasterizer.set_debug_flags(no_optimistic=True)
        out_ref, p_ref, g_ref = _run_fused(big, G6)
    finally:
        rasterizer.set_debug_flags()
    out_s, _, _ = _run_fused(small, G6)                       # leaves a small hint behind
    out_b, p_b, g_b = _run_fused(big, G6)                     # ~7x more instances than the hint
    assert int(out_b["num_rendered"][0]) > 3 * int(out_s["num_rendered"][0])     # far beyond hint * 1.125
    assert tuple(out_b["num_rendered"]) == tuple(out_ref["num_rendered"])
    assert torch.equal(p_b, p_ref)
    for k, v in g_ref.items():
        if v is not None:
            assert rel_err(g_b[k], v) < 2e-6, k
    out_s2, p_s2, _ = _run_fused(small, G6)                   # and back down: hint far too large is fine too
    assert torch.isfinite(p_s2).all() and tuple(out_s2["num_rendered"]) == tuple(out_s["num_rendered"])


def test_render_derived_outputs_match_the_elementwise_formulation():
    """uncertainty / presence_mask / nan_mask / visibility_filter / max_radii2D are written by the forward
    kernels (fsgs_render_extras); they must equal the reference's torch expressions
    (gaussian_renderer/__init__.py:70-88) bit for bit."""
    _, model, _, render = _gpu_modules()
    sc = make_scene(6000, 320, 256, size_mult=2.0, seed=4)
    poses, pc = model.scene_to_device(sc, DEV)
    pc.variables['max_radii2D'][::3] = 7.0
    mr0 = pc.variables['max_radii2D'].clone()
    xyz = pc.params['_xyz']
    with torch.no_grad():
        (_, depth, sil, dsq), radii, _, ex = render.render_planes(
            xyz, pc.params['_features_dc'], pc.params['_features_rest'], pc.params['_opacity'], pc.params['_scaling'],
            pc.params['_rotation'], poses.get_pose(0), torch.zeros_like(xyz), pc.cam, poses.cam_center,
            pc.active_sh_degree, max_radii2D=pc.variables['max_radii2D'], want_extras=True)
    unc, presence, nan_mask, visible, mr_done = ex
    want_unc = dsq.unsqueeze(0) - depth ** 2
    assert unc.shape == want_unc.shape and torch.equal(unc, want_unc)
    assert presence.dtype == torch.bool and torch.equal(presence, sil > 0.3)
    want_nan = (~torch.isnan(depth)) & (~torch.isnan(want_unc))
    assert nan_mask.shape == want_nan.shape and torch.equal(nan_mask, want_nan)
    assert torch.equal(visible, radii > 0) and int(visible.sum()) > 1000
    assert mr_done and torch.equal(pc.variables['max_radii2D'], torch.maximum(mr0, radii.float()))
    # the public render() returns the same objects, and an int max_radii2D still goes through torch
    pc.variables['max_radii2D'] = torch.zeros(sc.P, dtype=torch.int32, device=DEV)
    out = render.render(poses, 0, pc)
    assert torch.equal(out["uncertainty"], want_unc) and torch.equal(out["presence_mask"], sil > 0.3)
    assert torch.equal(out["nan_mask"], want_nan) and torch.equal(out["visibility_filter"], radii > 0)
    assert torch.equal(pc.variables['seen'], radii > 0)
    assert torch.equal(pc.variables['max_radii2D'], radii)


def test_edge_cases():
    _, _, rasterizer, _ = _gpu_modules()
    cam = make_camera(70, 50)
    rs = _settings_to_cuda(cam, rasterizer)
    z = lambda *s: torch.zeros(*s, device=DEV)
    # empty
    c, r, d = rasterizer.GaussianRasterizer(rs)(means3D=z(0, 3), means2D=z(0, 3), opacities=z(0, 1), colors_precomp=z(0, 3),
                                               scales=z(0, 3), rotations=z(0, 4))
    assert torch.allclose(c, torch.ones_like(c)) and d.abs().max().item() == 0 and r.numel() == 0
    # everything behind the camera / near plane
    m = torch.tensor([[0.0, 0.0, -1.0], [0.0, 0.0, 0.1]], device=DEV, requires_grad=True)
    c, r, d = rasterizer.GaussianRasterizer(rs)(means3D=m, means2D=z(2, 3), opacities=torch.ones(2, 1, device=DEV),
                                               colors_precomp=torch.rand(2, 3, device=DEV), scales=torch.ones(2, 3, device=DEV) * 0.01,
                                               rotations=torch.tensor([[1.0, 0, 0, 0]] * 2, device=DEV))
    assert (r == 0).all() and torch.allclose(c, torch.ones_like(c))
    c.sum().backward()
    assert m.grad.abs().max().item() == 0
    # XOR argument checks keep the reference's messages
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rasterizer.GaussianRasterizer(rs)(means3D=m, means2D=z(2, 3), opacities=z(2, 1), scales=z(2, 3), rotations=z(2, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        rasterizer.GaussianRasterizer(rs)(means3D=m, means2D=z(2, 3), opacities=z(2, 1), colors_precomp=z(2, 3))
    assert rasterizer.GaussianRasterizer(rs).markVisible(m.detach()).tolist() == [False, False]


@pytest.mark.parametrize("n_stack,ties", [(300, True), (9000, True), (3000, True), (3000, False)])
def test_long_tile_lists_and_depth_ties(n_stack, ties):
    """One tile with a very long list (multi-batch staging; > 8192 entries takes the global-memory
    sort path) and many exactly equal depths (order must fall back to the Gaussian id; the bucket sort
    declines such a list and the compare-exchange network takes over).  ties=False: 3000 distinct depths,
    the bucket path with 12 buckets per thread."""
    _, _, rasterizer, _ = _gpu_modules()
    from oracle import c_oracle as co
    W = H = 32
    cam = make_camera(W, H)
    g = torch.Generator().manual_seed(n_stack)
    P = n_stack
    means = torch.zeros(P, 3)
    means[:, 0] = (torch.rand(P, generator=g) - 0.5) * 0.01
    means[:, 1] = (torch.rand(P, generator=g) - 0.5) * 0.01
    if ties:
        means[:, 2] = 1.0 + 0.25 * torch.randint(0, 4, (P,), generator=g).float()     # only 4 distinct depths
    else:
        means[:, 2] = 1.0 + torch.rand(P, generator=g)
    d = dict(means3D=means, opacities=torch.full((P, 1), 0.005) + 0.007 * torch.rand(P, 1, generator=g),
             colors_precomp=torch.rand(P, 3, generator=g), scales=torch.full((P, 3), 0.05),
             rotations=torch.tensor([[1.0, 0, 0, 0]]).repeat(P, 1))
    rs = _settings_to_cuda(cam, rasterizer)
    gpu = {k: v.to(DEV).requires_grad_(True) for k, v in d.items()}
    c, r, dep = rasterizer.GaussianRasterizer(rs)(means2D=torch.zeros(P, 3, device=DEV), **gpu)
    Gc = torch.randn(3, H, W, generator=g)
    ref = {k: v.double().requires_grad_(True) for k, v in d.items()}
    c_ref, r_ref, dep_ref, aux = co.rasterize(means2D=torch.zeros(P, 3, dtype=torch.float64), st=cam, want_aux=True, **ref)
    Gc = Gc * (~fragile_mask(aux, H, W)).float()
    (c * Gc.to(DEV)).sum().backward()
    (c_ref * Gc.double()).sum().backward()
    assert (r.cpu() != r_ref).sum().item() == 0
    check_image("color", c, c_ref, aux)
    check_image("depth", dep, dep_ref, aux, scale=2.0)
    for k in ("means3D", "opacities", "colors_precomp", "scales"):
        check_grad(k, gpu[k].grad, ref[k].grad, tol=2e-4)


def test_sort_window_and_key_bins_do_not_change_results():
    """k_tile_sort is launched with a 32 KB shared-memory window when the PREVIOUS frame's longest tile list was short
    (4 resident CTAs instead of 3).  A frame whose lists then turn out longer -- beyond the small window's bucket
    path (2048), beyond the window itself (4096) -- must come out bit-identical to the 64 KB-window launch."""
    _, _, rasterizer, _ = _gpu_modules()
    W = H = 32
    cam = make_camera(W, H)
    rs = _settings_to_cuda(cam, rasterizer)

    def stack(P, seed):
        g = torch.Generator().manual_seed(seed)
        means = torch.zeros(P, 3)
        means[:, 0] = (torch.rand(P, generator=g) - 0.5) * 0.01
        means[:, 1] = (torch.rand(P, generator=g) - 0.5) * 0.01
        means[:, 2] = 1.0 + torch.rand(P, generator=g)
        return dict(means3D=means.to(DEV), opacities=(torch.full((P, 1), 0.005) + 0.007 * torch.rand(P, 1, generator=g)).to(DEV),
                    colors_precomp=torch.rand(P, 3, generator=g).to(DEV), scales=torch.full((P, 3), 0.05, device=DEV),
                    rotations=torch.tensor([[1.0, 0, 0, 0]]).repeat(P, 1).to(DEV), means2D=torch.zeros(P, 3, device=DEV))

    def run(d):
        c, r, dep = rasterizer.GaussianRasterizer(rs)(**d)
        return c.clone(), dep.clone()

    small, mid, long_ = stack(300, 1), stack(3000, 2), stack(9000, 3)
    res = {}
    try:
        for name, kw in (("auto", {}), ("large", dict(sort_window_large=True)), ("no_bins", dict(no_bins=True))):
            rasterizer.set_debug_flags(**kw)
            out = []
            for d in (mid, long_):
                run(small); run(small)          # leave "longest list = 300" behind (the library keeps the last two
                out.append(run(d))              # frames' values): the next launch takes the small window
            res[name] = out
    finally:
        rasterizer.set_debug_flags()
    # (the same sequence also makes a tile outgrow the per-tile key bin sized from the previous frame's longest list:
    # the optimistic kernels leave on their device guard and the host relaunches the scatter path)
    for other in ("large", "no_bins"):
        for (ca, da), (cl, dl) in zip(res["auto"], res[other]):
            assert torch.equal(ca, cl) and torch.equal(da, dl), other


def test_config4_size_runs_and_is_deterministic():
    """Config 4 size (2M Gaussians, 1280x1024): finite, deterministic forward, silhouette + T == 1, longer lists."""
    sc = make_scene(2_000_000, 1280, 1024, size_mult=2.0, seed=0)
    G6 = torch.zeros(6, 1024, 1280)
    G6[:3] = sc.grads_out["G_rgb"]
    G6[3] = sc.grads_out["G_dep"]
    out1, p1, g1 = _run_fused(sc, G6)
    out2, p2, g2 = _run_fused(sc, G6)
    assert torch.isfinite(p1).all() and torch.equal(p1, p2)
    assert (p1[4] - 1.0).abs().max().item() < 1e-4
    assert int(out1["num_rendered"][0]) > 4_000_000 and int(out1["num_rendered"][0]) < int(out1["num_rendered"][1])
    for k, v in g1.items():
        if v is not None:
            assert torch.isfinite(v).all(), k
            assert rel_err(g2[k], v) < 1e-5, k


def test_full_size_properties():
    """Config 2 size (500k Gaussians, 1280x1024): properties that need no oracle."""
    sc = make_scene(500_000, 1280, 1024, size_mult=2.0, seed=0)
    G6 = torch.zeros(6, 1024, 1280)
    G6[:3] = sc.grads_out["G_rgb"]
    G6[3] = sc.grads_out["G_dep"]
    out1, p1, g1 = _run_fused(sc, G6)
    out2, p2, g2 = _run_fused(sc, G6)
    assert torch.isfinite(p1).all()
    assert torch.equal(p1, p2), "forward must be deterministic (sorted order fixes the scatter order)"
    assert (p1[4] - 1.0).abs().max().item() < 1e-4, "silhouette + T_final == 1 (white background, quirk iv)"
    assert int(out1["num_rendered"][1]) == 3_637_628 or abs(int(out1["num_rendered"][1]) - 3_640_000) < 40_000
    for k, v in g1.items():
        if v is not None:
            assert torch.isfinite(v).all(), k
            assert rel_err(g2[k], v) < 1e-5, k               # atomics order only
    # linearity of the backward in the upstream gradient
    _, _, g3 = _run_fused(sc, 2.0 * G6)
    for k, v in g1.items():
        if v is not None:
            assert rel_err(g3[k], 2.0 * v) < 1e-5, k
    # sortedness of every tile list, read back through the documented buffer layout
    _, model, rasterizer, _ = _gpu_modules()
    poses, pc = model.scene_to_device(sc, DEV)
    with torch.no_grad():
        means_cam = model.transform_to_frame(pc.params["_xyz"], poses.get_pose(0))
        rs = pc.cam
        nr, color, depth, radii, geom, binning, img = rasterizer.rasterize_gaussians(
            rs.bg, means_cam, torch.rand(sc.P, 3, device=DEV), pc.get_opacity, pc.get_scaling, pc.get_rotation, 1.0,
            torch.empty(0, device=DEV), rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
            rs.image_width, torch.empty(0, device=DEV), 0, rs.campos, False, True)
    assert nr == int(out1["num_rendered"][0])
    T = (1280 // 16) * (1024 // 16)
    io, bo = _lib.img_offsets(1280, 1024), _lib.binning_offsets(nr)
    tile_offset = img[io["tile_offset"]:io["tile_offset"] + 4 * (T + 1)].view(torch.int32).long()
    assert bo["records"] == 0
    rec = binning[:48 * nr].view(torch.int32).view(nr, 12).long()
    keys = (rec[:, 9] << 32) | rec[:, 11]              # (depth float bits, Gaussian id) of every sorted instance
    assert int(tile_offset[-1]) == nr and (tile_offset[1:] >= tile_offset[:-1]).all()
    asc = keys[1:] >= keys[:-1]                       # depth bits are positive floats -> signed compare is fine
    boundary = torch.zeros(nr - 1, dtype=torch.bool, device=DEV)
    inner = tile_offset[1:-1]
    boundary[(inner[(inner > 0) & (inner < nr)] - 1)] = True
    assert bool((asc | boundary).all()), "every tile list must be sorted by (depth bits, Gaussian id)"
    ids = rec[:, 11]
    assert int(ids.max()) < sc.P and int((radii[ids] > 0).all())
    assert int((radii > 0).sum()) > 400_000
