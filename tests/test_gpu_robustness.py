"""Host-layer robustness of the fused path on the GPU: empty models, several renders inside one captured
step, recapture after the model grew, shape validation, the device watchdog flag."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from parity import rel_err  # noqa: E402

from fsgs_b200.synth import frame_pose_params, make_scene  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(P=20000, W=320, H=256, m=2.0, seed=5, n_cams=1):
    from fsgs_b200 import frame_render as render
    from fsgs_b200 import model
    sc = make_scene(P, W, H, size_mult=m, seed=seed)
    poses, pc = model.scene_to_device(sc, DEV)
    return render, model, sc, poses, pc


def test_empty_model_forward_and_backward_reach_the_pose():
    """P == 0 (an empty or fully pruned model): the frame is the background and the backward must still run --
    zero pose gradient, empty parameter gradients (round-1 advisor finding: the backward dereferenced a
    saved tensor that is None for an empty model)."""
    render, model, sc, poses, pc = _setup(P=64)
    empty = {k: v.detach()[:0].clone().requires_grad_(True) for k, v in pc.params.items()}
    pc0 = model.SplatModel(empty, cam=pc.cam)
    out = render.render(poses, 0, pc0, gs_grad=True, cam_grad=True)
    assert out["render"].shape == (3, sc.height, sc.width) and torch.allclose(out["render"], torch.ones_like(out["render"]))
    assert out["radii"].numel() == 0 and out["visibility_filter"].numel() == 0
    (out["render"].sum() + out["render_dep"].sum()).backward()
    assert poses.pose_param_net.r.grad is not None and poses.pose_param_net.r.grad.abs().max().item() == 0
    assert poses.pose_param_net.t.grad.abs().max().item() == 0
    for k, v in pc0.params.items():
        assert v.grad is not None and v.grad.shape == v.shape, k


def test_fused_render_rejects_wrong_parameter_shapes():
    render, model, sc, poses, pc = _setup(P=500, W=96, H=64)
    bad = dict(pc.params)
    bad["_features_rest"] = pc.params["_features_rest"][:, :8].detach().clone()      # max_sh_degree = 2 storage
    pc2 = model.SplatModel(bad, cam=pc.cam)
    with pytest.raises(ValueError, match="_features_rest has shape"):
        render.render(poses, 0, pc2)
    bad = dict(pc.params)
    bad["_opacity"] = pc.params["_opacity"].detach().reshape(-1).clone()
    with pytest.raises(ValueError, match="_opacity has shape"):
        render.render(poses, 0, model.SplatModel(bad, cam=pc.cam))


def test_graphed_step_with_two_frames_sizes_its_capacity_from_the_larger_one():
    """A captured step may render several frames.  The fixed binning capacity must cover the LARGEST of them, not
    the last (round-1 advisor finding: frame A 2x the instances of frame B made every replay of frame A skip its
    binning / compositing kernels and return stale planes and zero gradients)."""
    from fsgs_b200 import GraphedStep, model
    render, _, sc, _, pc = _setup(P=30000, W=320, H=256)
    cam = sc.camera
    K = [[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]]
    poses = model.FramePoses(2, K, sc.width, sc.height, device=DEV)
    pc.cam = poses.setup_camera(torch.eye(4).numpy())
    poses.set_pose(0, sc.pose_q.tolist(), sc.pose_t.tolist())
    # frame 1 looks at the scene from further back: its splats are smaller on screen -> far fewer tile instances
    poses.set_pose(1, sc.pose_q.tolist(), (sc.pose_t + torch.tensor([0.0, 0.0, 1.5])).tolist())
    G = torch.randn(3, sc.height, sc.width, generator=torch.Generator().manual_seed(1)).to(DEV)

    def step():
        pc.zero_grad()
        poses.pose_param_net.zero_grad(set_to_none=True)
        oa = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
        ob = render.render(poses, 1, pc, gs_grad=True, cam_grad=True)
        ((oa["render"] * G).sum() + (ob["render"] * G).sum()).backward()
        return (oa["render"].detach(), ob["render"].detach(), pc.params["_xyz"].grad, poses.pose_param_net.r.grad,
                torch.tensor([oa["num_rendered"][0], ob["num_rendered"][0]]))

    eager = [x.clone() for x in step()]
    na, nb = int(eager[4][0]), int(eager[4][1])
    assert na > 1.5 * nb, (na, nb)                      # the first frame is the large one
    gs = GraphedStep(step, warmup=2)
    assert gs.capacity >= na, (gs.capacity, na, nb)
    out = gs.replay()
    torch.cuda.synchronize()
    assert not gs.overflowed()
    assert sorted(gs.instance_counts()) == sorted([na, nb])
    assert torch.equal(out[0], eager[0]) and torch.equal(out[1], eager[1])
    assert rel_err(out[2], eager[2]) < 1e-5 and rel_err(out[3], eager[3]) < 1e-5
    gs.release()


def test_recapture_after_the_model_grew_and_the_pool_does_not_strand_buffers():
    """Densification changes P and the instance count between frames: the optimistic tail must relaunch, a captured
    step must report the overflow and recapture, and repeated recaptures must not leave their warm-up scratch
    cached under dead stream keys."""
    from fsgs_b200 import GraphedStep, model
    from fsgs_b200.rasterizer import _POOL
    render, _, sc, poses, pc = _setup(P=8000, W=320, H=256, m=1.0)
    G = torch.randn(3, sc.height, sc.width, generator=torch.Generator().manual_seed(2)).to(DEV)
    state = {"pc": pc}

    def step():
        m_ = state["pc"]
        m_.zero_grad()
        poses.pose_param_net.zero_grad(set_to_none=True)
        o = render.render(poses, 0, m_, gs_grad=True, cam_grad=True)
        (o["render"] * G).sum().backward()
        return o["render"].detach(), m_.params["_xyz"].grad

    gs = GraphedStep(step, warmup=2)
    gs.replay()
    assert not gs.overflowed()
    # "densify": the same Gaussians twice over with much larger splats, IN PLACE so the captured graph sees them
    with torch.no_grad():
        pc.params["_scaling"] += 1.2
    gs.replay()
    torch.cuda.synchronize()
    assert gs.overflowed(), "3x larger splats must outgrow a capacity of 1.25x the warm-up count"
    gs.recapture()
    out = gs.replay()
    torch.cuda.synchronize()
    assert not gs.overflowed()
    eager = step()
    assert torch.equal(out[0], eager[0]) and rel_err(out[1], eager[1]) < 1e-5
    # a model with a different P needs a new capture (new tensors); nothing of the old captures may linger in the pool
    keys_before = {k[:2] for k in _POOL._free}
    big = make_scene(16000, 320, 256, size_mult=1.0, seed=6)
    _, pc2 = model.scene_to_device(big, DEV)
    state["pc"] = pc2
    for _ in range(3):
        gs.recapture()
    keys_after = {k[:2] for k in _POOL._free}
    assert len(keys_after) <= len(keys_before) + 1, (keys_before, keys_after)
    out2 = gs.replay()
    torch.cuda.synchronize()
    assert not gs.overflowed() and torch.isfinite(out2[0]).all()
    gs.release()


def test_watchdog_flag_is_clear_after_normal_work():
    from fsgs_b200 import _lib
    render, _, sc, poses, pc = _setup(P=5000, W=160, H=128)
    out = render.render(poses, 0, pc)
    out["render"].sum().backward()
    assert _lib.lib().fsgs_watchdog_flag(0, 0) == 0


def test_distCUDA2_at_the_real_initialisation_size_is_exact_and_memory_bounded():
    """``simple_knn._C.distCUDA2`` stand-in (reference submodules/simple-knn/simple_knn.cu:147-219) at the size the
    reference calls it with: ~10 % of a 1280x1024 frame's pixels (gaussian_model.py:247,346) back-projected through a
    depth map.  Exact against a float64 k-d tree; peak memory stays within a few tiles of the 256 MiB budget (the
    round-1 version needed 3.2 GB here)."""
    import numpy as np
    from scipy.spatial import cKDTree
    from simple_knn._C import distCUDA2
    g = torch.Generator().manual_seed(3)
    P = 131072
    u, v = torch.rand(P, generator=g) * 1280, torch.rand(P, generator=g) * 1024
    zd = 1.0 + 0.4 * torch.sin(u / 200) * torch.cos(v / 160) + 0.01 * torch.randn(P, generator=g)
    pts = torch.stack([(u - 640) / 1035 * zd, (v - 512) / 1035 * zd, zd], dim=1)
    dev_pts = pts.to(DEV)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    d = distCUDA2(dev_pts)
    torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated() - base
    assert peak < 1.5 * (1 << 30), f"distCUDA2 peak memory {peak / 2**20:.0f} MiB"
    dd, _ = cKDTree(pts.double().numpy()).query(pts.double().numpy(), k=4)
    ref = (dd[:, 1:] ** 2).mean(axis=1)
    err = np.abs(d.cpu().numpy().astype(np.float64) - ref) / ref
    assert d.shape == (P,) and float(err.max()) < 1e-4, float(err.max())
    # the reference's caller: scales = log(sqrt(clamp_min(dist2, 1e-7))) must be finite
    assert torch.isfinite(torch.log(torch.sqrt(torch.clamp_min(d, 1e-7)))).all()
