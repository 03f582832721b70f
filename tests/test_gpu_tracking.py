"""End-to-end use of the pose gradient (the SfM-free tracking loop of train.py:154-210, reduced to
its core): starting from a perturbed pose, Adam on (r, t) with the render loss against an image
rendered at the true pose must recover the pose.  Also checks fused vs two-pass consistency at a
mid-size scene and the mapping direction (Gaussian colours recover)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from parity import rel_err  # noqa: E402

from fsgs_b200.synth import make_scene  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(P=20000, W=320, H=256):
    from fsgs_b200 import frame_render as render
    from fsgs_b200 import model
    sc = make_scene(P, W, H, size_mult=2.0, seed=5)
    poses, pc = model.scene_to_device(sc, DEV)
    return render, model, sc, poses, pc


def test_pose_tracking_recovers_a_perturbed_pose():
    render, model, sc, poses, pc = _setup()
    with torch.no_grad():
        target = render.render(poses, 0, pc, gs_grad=False, cam_grad=False)["render"].clone()
        r_true = poses.pose_param_net.r.detach().clone()
        t_true = poses.pose_param_net.t.detach().clone()
        poses.pose_param_net.r += torch.tensor([0.0, 0.004, -0.003, 0.002], device=DEV).view(1, 4, 1)
        poses.pose_param_net.t += torch.tensor([0.01, -0.008, 0.006], device=DEV).view(3, 1)
    def pose_err():
        Rt = poses.get_pose(0).detach()
        poses_true = model.LearnPose(1, device=DEV)
        with torch.no_grad():
            poses_true.r.copy_(r_true); poses_true.t.copy_(t_true)
        return (Rt - poses_true.forward(0).detach()).abs().max().item()
    e0 = pose_err()
    opt = torch.optim.Adam([poses.pose_param_net.r, poses.pose_param_net.t], lr=1e-3, eps=1e-15)
    losses = []
    for _ in range(150):
        opt.zero_grad(set_to_none=True)
        pc.zero_grad()
        out = render.render(poses, 0, pc, gs_grad=False, cam_grad=True)       # tracking mode (train.py:167-171)
        loss = (out["render"] - target).abs().mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    e1 = pose_err()
    assert losses[-1] < 0.15 * losses[0], (losses[0], losses[-1])
    assert e1 < 0.2 * e0, (e0, e1)


def test_fused_and_two_pass_agree_at_mid_size():
    render, model, sc, poses, pc = _setup(P=60000, W=640, H=512)
    G = torch.randn(4, 512, 640, generator=torch.Generator().manual_seed(3)).to(DEV)
    res = {}
    for name, fn in (("fused", render.render), ("two_pass", render.render_two_pass)):
        pc.zero_grad()
        poses.pose_param_net.zero_grad(set_to_none=True)
        out = fn(poses, 0, pc, gs_grad=True, cam_grad=True)
        ((out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()).backward()
        res[name] = (out["render"].detach().clone(), out["render_dep"].detach().clone(), out["radii"].clone(),
                     {k: v.grad.detach().clone() for k, v in pc.params.items()},
                     poses.pose_param_net.r.grad.clone(), poses.pose_param_net.t.grad.clone(),
                     out["viewspace_points"].grad.detach().clone())
    a, b = res["fused"], res["two_pass"]
    assert torch.equal(a[2], b[2])
    # the two paths differ only in float32 rounding of the pre-processing (fused SH / pose transform)
    assert ((a[0] - b[0]).abs().amax(0) > 1e-5).float().mean().item() < 2e-3
    assert ((a[1] - b[1]).abs() > 2e-5).float().mean().item() < 2e-3
    for k in a[3]:
        assert rel_err(a[3][k], b[3][k]) < 3e-4, k        # includes a few decision flips between the two roundings
    assert rel_err(a[4], b[4]) < 3e-4 and rel_err(a[5], b[5]) < 3e-4
    assert rel_err(a[6], b[6]) < 3e-4


def test_mapping_step_reduces_the_render_loss():
    render, model, sc, poses, pc = _setup()
    with torch.no_grad():
        target = render.render(poses, 0, pc)["render"].clone()
        pc.params["_features_dc"] += 0.3 * torch.randn_like(pc.params["_features_dc"])
    opt = torch.optim.Adam([pc.params["_features_dc"]], lr=2e-2, eps=1e-15)
    l0 = l1 = None
    for it in range(60):
        opt.zero_grad(set_to_none=True)
        out = render.render(poses, 0, pc, gs_grad=True, cam_grad=False)       # mapping mode (train.py:245-249)
        loss = (out["render"] - target).abs().mean()
        loss.backward()
        opt.step()
        l0 = loss.item() if l0 is None else l0
        l1 = loss.item()
    assert l1 < 0.3 * l0, (l0, l1)


def test_concurrent_threads_and_side_stream_give_identical_frames():
    """The reference calls the rasteriser from the training thread AND from the viewer thread
    (train.py:124-152); the library keeps no global mutable state besides the scratch pool (locked,
    stream-keyed).  Two Python threads rendering concurrently, and a render on a non-default stream,
    must reproduce the serial result bit for bit."""
    import threading
    render, model, sc, poses, pc = _setup(P=15000, W=256, H=192)
    with torch.no_grad():
        ref = render.render(poses, 0, pc, gs_grad=False, cam_grad=False)
        ref_planes = torch.cat([ref["render"], ref["render_dep"][None]]).clone()
    results, errors = {}, []

    def worker(tid):
        try:
            torch.cuda.set_device(0)
            outs = []
            for _ in range(6):
                with torch.no_grad():
                    o = render.render(poses, 0, pc, gs_grad=False, cam_grad=False)
                outs.append(torch.cat([o["render"], o["render_dep"][None]]).clone())
            results[tid] = outs
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    assert not errors, errors
    for outs in results.values():
        for o in outs:
            assert torch.equal(o, ref_planes)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        o = render.render(poses, 0, pc, gs_grad=False, cam_grad=False)
        side_planes = torch.cat([o["render"], o["render_dep"][None]]).clone()
    side.synchronize()
    assert torch.equal(side_planes, ref_planes)
    # training step on the side stream: gradients equal the default-stream ones up to atomics order
    G = torch.randn(3, 192, 256, generator=torch.Generator().manual_seed(9)).to(DEV)
    grads = {}
    for name, stream in (("default", torch.cuda.current_stream()), ("side", side)):
        pc.zero_grad()
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
            (out["render"] * G).sum().backward()
        stream.synchronize()
        grads[name] = {k: v.grad.detach().clone() for k, v in pc.params.items()}
    for k in grads["default"]:
        assert rel_err(grads["side"][k], grads["default"][k]) < 1e-5, k


def test_frame_parallel_compact_exchange_equals_full_gradient_sum():
    """Two 'ranks' simulated on one GPU: the compact exchange (14 floats/Gaussian reduced inside the backward, SH
    gradients expanded from the reduced masked colour gradient by fsgs_sh_grad_expand) must give the same
    gradients as summing the two frames' full gradients."""
    from fsgs_b200 import frame_render as render
    from fsgs_b200 import model
    from fsgs_b200.synth import frame_pose_params, make_scene
    sc = make_scene(20000, 320, 256, size_mult=2.0, seed=3)
    G = [torch.randn(4, sc.height, sc.width, generator=torch.Generator().manual_seed(10 + k)).cuda() for k in range(2)]

    def run(frame, sh_deg):
        poses, pc = model.scene_to_device(sc, "cuda")
        pc.active_sh_degree = sh_deg
        poses.set_pose(0, *frame_pose_params(frame))
        out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
        ((out["render"] * G[frame][:3]).sum() + (out["render_dep"] * G[frame][3]).sum()).backward()
        return {k: p.grad.detach().clone() for k, p in pc.params.items()}, poses.pose_param_net.r.grad.clone()

    for sh_deg in (3, 2):
        full = [run(f, sh_deg) for f in range(2)]
        want = {k: full[0][0][k] + full[1][0][k] for k in full[0][0]}
        other = {}
        try:
            render.set_grad_reducer(lambda flat: other.__setitem__("compact", flat.clone()))   # rank 1: just record
            run(1, sh_deg)
            render.set_grad_reducer(lambda flat: flat.add_(other["compact"]))                  # rank 0: SUM with rank 1
            got, r_grad = run(0, sh_deg)
        finally:
            render.set_grad_reducer(None)
        for k, v in want.items():
            err = (got[k] - v).norm() / v.norm()
            assert err < 1e-5, (sh_deg, k, float(err))
        if sh_deg < 3:
            assert got["_features_rest"][:, (sh_deg + 1) ** 2 - 1:, :].abs().max().item() == 0
        # pose gradients stay local (same frame rendered twice: float atomics order only)
        assert rel_err(r_grad, full[0][1]) < 1e-5


def test_scratch_returns_to_the_pool_without_the_cyclic_gc():
    """The API path hands its geometry / binning / image-state buffers to the caller (as the reference's pybind
    module does); they must return to the workspace pool by plain reference counting when the autograd graph
    dies, so steady-state frames allocate nothing new (a reference cycle here once leaked ~400 MB per frame
    until the cyclic collector ran)."""
    import gc
    from fsgs_b200 import frame_render as render
    from fsgs_b200 import model
    sc = make_scene(50_000, 640, 512, size_mult=2.0, seed=1)
    poses, pc = model.scene_to_device(sc, "cuda")
    G = torch.randn(4, sc.height, sc.width, generator=torch.Generator().manual_seed(0)).cuda()

    def step(fn):
        pc.zero_grad()
        poses.pose_param_net.zero_grad(set_to_none=True)
        out = fn(poses, 0, pc, gs_grad=True, cam_grad=True)
        ((out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()).backward()

    gc.collect()
    gc.disable()
    try:
        for fn in (render.render_two_pass, render.render):
            for _ in range(3):
                step(fn)
            torch.cuda.synchronize()
            base = torch.cuda.memory_allocated()
            for _ in range(6):
                step(fn)
            torch.cuda.synchronize()
            grown = torch.cuda.memory_allocated() - base
            assert grown < (8 << 20), (fn.__name__, grown >> 20)
    finally:
        gc.enable()


def test_cuda_graph_capture_of_a_whole_step_matches_eager():
    """fsgs_b200.GraphedStep: forward + loss + backward captured once (fixed-capacity mode, no host read-back) and
    replayed; parameters / pose / targets are updated in place between replays.  Also the overflow protocol: more
    instances than the captured capacity -> kernels skip themselves, overflowed() says so, recapture() fixes it."""
    import fsgs_b200
    from fsgs_b200 import frame_render as render
    from fsgs_b200 import model
    sc = make_scene(20000, 320, 256, size_mult=1.5, seed=11)
    poses, pc = model.scene_to_device(sc, "cuda")
    G = torch.randn(4, sc.height, sc.width, generator=torch.Generator().manual_seed(5)).cuda()
    keys = list(pc.params)

    def step():
        pc.zero_grad()
        poses.pose_param_net.zero_grad(set_to_none=True)
        out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
        loss = (out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()
        loss.backward()
        return loss.detach(), out["render"].detach(), poses.pose_param_net.r.grad, [pc.params[k].grad for k in keys]

    def eager():
        loss, img, rg, pg = step()
        return loss.clone(), img.clone(), rg.clone(), [g.clone() for g in pg]

    def check(got, want, tag):
        assert torch.equal(got[1], want[1]), tag + ": image"
        assert abs(float(got[0]) - float(want[0])) <= 1e-5 * abs(float(want[0])), tag
        assert rel_err(got[2], want[2]) < 1e-5, tag + ": pose grad"
        for k, a, b in zip(keys, got[3], want[3]):
            assert rel_err(a, b) < 1e-5, (tag, k)

    want0 = eager()
    gs = fsgs_b200.GraphedStep(step, warmup=2, headroom=1.25)
    check(gs.replay(), want0, "first replay")
    assert not gs.overflowed() and gs.instance_counts()[0] > 10000
    # new pose and new target, written in place; replay must follow
    with torch.no_grad():
        poses.pose_param_net.t[:, 0] += torch.tensor([0.004, -0.003, 0.01], device="cuda")
        G.mul_(0.5)
    got = gs.replay()
    got = (got[0].clone(), got[1].clone(), got[2].clone(), [g.clone() for g in got[3]])
    check(got, eager(), "after in-place update")
    # grow the splats: far more instances than the captured capacity
    with torch.no_grad():
        pc.params["_scaling"] += 0.9
    gs.replay()
    assert gs.overflowed()
    gs.recapture()
    got = gs.replay()
    assert not gs.overflowed()
    got = (got[0].clone(), got[1].clone(), got[2].clone(), [g.clone() for g in got[3]])
    check(got, eager(), "after recapture")
    # release(): the graph is destroyed (what a frame-parallel job does before tearing its process group down --
    # NCCL keeps a communicator alive while a captured graph references it); eager steps keep working
    gs.release()
    assert gs.graph is None and gs.outputs is None
    check(eager(), got, "eager after release")


def test_frozen_model_forward_is_bit_identical_and_follows_the_model():
    """Tracking against a frozen model renders from the pose-independent rows of fsgs_freeze_model
    (fsgs_render_forward_frozen): the frame, radii, derived maps and the pose gradient must equal the plain forward
    bit for bit; rows are re-evaluated when a parameter is written to; a captured step follows too."""
    from fsgs_b200 import GraphedStep
    render, model, sc, poses, pc = _setup(P=30000, W=640, H=512)
    for v in pc.params.values():
        v.requires_grad_(False)
    G = torch.randn(3, 512, 640, generator=torch.Generator().manual_seed(11)).to(DEV)

    def step():
        # (returns detached copies only: an autograd graph kept alive here would pin the pose parameters'
        # AccumulateGrad nodes to this stream and break the capture below)
        poses.pose_param_net.zero_grad(set_to_none=True)
        out = render.render(poses, 0, pc, gs_grad=False, cam_grad=True)
        (out["render"] * G).sum().backward()
        snap = {k: out[k].detach().clone() for k in ("render", "render_dep", "render_opacity", "uncertainty", "radii",
                                                     "visibility_filter", "presence_mask", "nan_mask")}
        return snap, poses.pose_param_net.r.grad.clone(), poses.pose_param_net.t.grad.clone()

    def snapshot(out):
        return list(out.values())

    def same(a, b):
        return all(torch.equal(x, y) for x, y in zip(a, b))

    for trial in range(2):
        render.USE_FROZEN_MODEL = False
        out, gr0, gt0 = step()
        ref = snapshot(out)
        render.USE_FROZEN_MODEL = True
        out, gr1, gt1 = step()                   # evaluates the rows (trial 0) / finds them stale (trial 1)
        assert same(ref, snapshot(out)), trial
        out, gr2, gt2 = step()                   # reuses them
        assert same(ref, snapshot(out)), trial
        # the pose gradient is a float sum over atomics: equal up to summation order
        assert rel_err(gr1, gr0) < 1e-5 and rel_err(gt1, gt0) < 1e-5 and rel_err(gr2, gr0) < 1e-5
        with torch.no_grad():                    # in-place change of the model: the rows must follow
            pc.params["_features_dc"] += 0.05
            pc.params["_scaling"] -= 0.02
            pc.params["_xyz"] += 0.001
    # a captured tracking step reads the rows' buffer: replay() refreshes it when the model was written to
    static_G = G.clone()

    def captured():
        poses.pose_param_net.r.grad = None
        poses.pose_param_net.t.grad = None
        out = render.render(poses, 0, pc, gs_grad=False, cam_grad=True)
        (out["render"] * static_G).sum().backward()
        return out["render"].detach()

    gs = GraphedStep(captured)
    assert len(gs._frozen) == 1
    img = gs.replay().clone()
    render.USE_FROZEN_MODEL = False
    out, _, _ = step()
    assert torch.equal(img, out["render"])
    with torch.no_grad():
        pc.params["_opacity"] -= 0.3
    out, _, _ = step()                            # plain forward on the changed model
    render.USE_FROZEN_MODEL = True
    img = gs.replay().clone()
    assert torch.equal(img, out["render"])
    gs.release()
