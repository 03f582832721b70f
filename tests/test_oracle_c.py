"""The plain-C oracle (closed-form backward, upstream regulariser) against the float64
torch.autograd oracle (independent formulation) on small scenes."""
import pytest
import torch

from oracle import c_oracle as co
from oracle import raster_oracle as ro
from fsgs_b200.synth import make_camera, make_scene, pose_matrix


def _inputs(sc, dt, mode, spread=1.0):
    P = sc.P
    xyz = (sc.Rt(dt) @ torch.cat([sc.params["_xyz"].to(dt), torch.ones(P, 1, dtype=dt)], 1).T).T[:, :3]
    xyz = xyz * torch.tensor([spread, spread, 1.0], dtype=dt)
    d = dict(means3D=xyz.clone(), means2D=torch.zeros(P, 3, dtype=dt),
             opacities=torch.sigmoid(sc.params["_opacity"].to(dt)))
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
    return {k: v.detach().clone().requires_grad_(True) for k, v in d.items()}


@pytest.mark.parametrize("mode,W,H,spread,sh_deg", [("sh", 200, 136, 1.0, 3), ("precomp", 97, 75, 1.0, 0),
                                                    ("cov", 128, 96, 1.0, 0), ("sh", 160, 128, 2.2, 2)])
def test_c_oracle_matches_autograd_oracle(mode, W, H, spread, sh_deg):
    dt = torch.float64
    sc = make_scene(500, W, H, size_mult=2.0, seed=3)
    cam = make_camera(W, H, pose_matrix((1, 0.05, -0.03, 0.02), (0.02, 0.01, -0.03)))
    cam.bg = torch.tensor([0.2, 0.7, 1.0])
    cam.sh_degree = sh_deg
    cam.scale_modifier = 1.0
    gen = torch.Generator().manual_seed(5)
    Gc = torch.randn(3, H, W, dtype=dt, generator=gen)
    Gd = torch.randn(1, H, W, dtype=dt, generator=gen)
    res = {}
    for name, fn in (("py", lambda **kw: ro.rasterize(st=cam, **kw)[:3]), ("c", lambda **kw: co.rasterize(st=cam, **kw)[:3])):
        inp = _inputs(sc, dt, mode, spread)
        color, radii, depth = fn(**inp)
        ((color * Gc).sum() + (depth * Gd).sum()).backward()
        res[name] = (color.detach(), radii, depth.detach(), {k: v.grad for k, v in inp.items()})
    a, b = res["py"], res["c"]
    assert (a[1] != b[1]).sum().item() == 0
    assert (a[1] > 0).sum().item() > 100
    assert (a[0] - b[0]).abs().max().item() < 1e-12
    assert (a[2] - b[2]).abs().max().item() < 1e-12
    for k in a[3]:
        rel = ((a[3][k] - b[3][k]).norm() / a[3][k].norm().clamp(min=1e-30)).item()
        # 1/(det^2+1e-7) regulariser of the closed form is the only intended deviation
        assert rel < 2e-5, (k, rel)
    if spread > 2:   # the +-1.3 tanfov clamp branch must have been exercised
        pre = ro.preprocess(_inputs(sc, dt, mode, spread)["means3D"].detach(), torch.zeros(sc.P, 3, dtype=dt), None,
                            torch.ones(sc.P, 3, dtype=dt), torch.ones(sc.P, 4, dtype=dt), None,
                            ro.RasterSettings.from_any(cam, dt), colors_precomp=torch.zeros(sc.P, 3, dtype=dt))
        t = pre["p_view"]
        assert ((t[:, 0] / t[:, 2]).abs() > 1.3 * cam.tanfovx).sum() > 0


def test_c_oracle_float32_close_to_float64():
    sc = make_scene(400, 160, 128, size_mult=2.0, seed=1)
    cam = make_camera(160, 128)
    out = {}
    for dt in (torch.float32, torch.float64):
        inp = _inputs(sc, dt, "precomp")
        color, radii, depth, _ = co.rasterize(st=cam, **inp)
        out[dt] = (color.detach().double(), depth.detach().double())
    assert (out[torch.float32][0] - out[torch.float64][0]).abs().max().item() < 5e-3   # decision flips bound
    assert (out[torch.float32][0] - out[torch.float64][0]).abs().median().item() < 1e-6


def test_empty_scene():
    cam = make_camera(64, 48)
    z = lambda *s: torch.zeros(*s, dtype=torch.float64)
    color, radii, depth, h, nr = co.forward(z(0, 3), z(0, 1), cam, colors_precomp=z(0, 3), scales=z(0, 3), rotations=z(0, 4))
    assert nr == 0 and radii.numel() == 0
    assert torch.allclose(color, torch.ones_like(color)) and depth.abs().max().item() == 0
