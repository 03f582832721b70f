"""Densification on the GPU path (BASELINE.json configs[3] "2M Gaussians, densification on"):

  * add_densification_stats folded into the fused backward == the reference's formulation on viewspace_points.grad;
  * the frame-parallel exchange in Gaussian ranges on a side stream == the single exchange == the plain sum;
  * a mapping loop at config-4 size with densify_and_prune every few iterations: the Gaussian count and the instance
    count change under the workspace pool, the optimistic forward tail and a captured step (overflow -> recapture).
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from parity import rel_err, report  # noqa: E402

from fsgs_b200.synth import frame_pose_params, make_scene  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"
LR = {"_xyz": 1.6e-4 * 5, "_features_dc": 0.0025, "_features_rest": 0.0025 / 20, "_opacity": 0.05, "_scaling": 0.005,
      "_rotation": 0.001}


def _setup(P, W, H, m=2.0, seed=5):
    from fsgs_b200 import frame_render as render
    from fsgs_b200 import model
    sc = make_scene(P, W, H, size_mult=m, seed=seed)
    poses, pc = model.scene_to_device(sc, DEV)
    return render, model, sc, poses, pc


@pytest.mark.parametrize("P,W,H", [(20000, 320, 256), (1203, 120, 88)])
def test_folded_densification_stats_equal_the_reference_formulation(P, W, H):
    from fsgs_b200 import densify
    render, model, sc, poses, pc = _setup(P, W, H)
    G = torch.randn(4, H, W, generator=torch.Generator().manual_seed(1)).to(DEV)

    def step(fold):
        pc.zero_grad()
        pc.variables['xyz_gradient_accum'].zero_(); pc.variables['denom'].zero_()
        pc.fold_densification_stats = fold
        outs = []
        for _ in range(2):                                  # two iterations accumulate
            out = render.render(poses, 0, pc, gs_grad=True, cam_grad=False)
            ((out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()).backward()
            if not fold:
                densify.add_densification_stats(pc.variables, out["viewspace_points"], out["visibility_filter"])
            outs.append(out)
        return pc.variables['xyz_gradient_accum'].clone(), pc.variables['denom'].clone(), outs[-1]

    a_ref, d_ref, out = step(False)
    a_fold, d_fold, _ = step(True)
    pc.fold_densification_stats = False
    assert torch.equal(d_fold, d_ref) and float(d_ref.max()) == 2.0
    assert torch.equal(d_ref[:, 0] > 0, out["visibility_filter"])
    assert rel_err(a_fold, a_ref) < 1e-5                     # float atomics order of the two backward passes
    assert float(a_ref.max()) > 0


def test_chunked_overlapped_exchange_equals_the_single_exchange_and_the_plain_sum():
    """Two 'ranks' simulated on one GPU.  chunks = 1: one exchange of the 56-byte rows after the per-Gaussian kernel;
    chunks = 3: the kernel runs in three Gaussian ranges and each range is exchanged + expanded on the side stream
    while the next one is computed.  Both must reproduce the sum of the two frames' full gradients."""
    from fsgs_b200 import frame_render as render
    from fsgs_b200 import model
    sc = make_scene(20003, 320, 256, size_mult=2.0, seed=3)          # ragged: the last range is short and unaligned
    G = [torch.randn(4, sc.height, sc.width, generator=torch.Generator().manual_seed(10 + k)).cuda() for k in range(2)]

    def run(frame):
        poses, pc = model.scene_to_device(sc, "cuda")
        poses.set_pose(0, *frame_pose_params(frame))
        out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
        ((out["render"] * G[frame][:3]).sum() + (out["render_dep"] * G[frame][3]).sum()).backward()
        torch.cuda.synchronize()
        return {k: p.grad.detach().clone() for k, p in pc.params.items()}, poses.pose_param_net.r.grad.clone()

    full = [run(f) for f in range(2)]
    want = {k: full[0][0][k] + full[1][0][k] for k in full[0][0]}
    for chunks in (1, 3):
        recorded = []
        try:
            render.set_grad_reducer(lambda flat: recorded.append(flat.clone()), chunks=chunks)      # rank 1: record
            run(1)
            assert len(recorded) == chunks and sum(x.numel() for x in recorded) == 14 * sc.P
            it = iter(recorded)
            render.set_grad_reducer(lambda flat: flat.add_(next(it)), chunks=chunks)               # rank 0: SUM
            got, r_grad = run(0)
        finally:
            render.set_grad_reducer(None)
        for k, v in want.items():
            assert rel_err(got[k], v) < 1e-5, (chunks, k)
        assert rel_err(r_grad, full[0][1]) < 1e-5                    # pose gradients stay local


def test_config4_size_mapping_loop_with_densification():
    """2 M Gaussians, 1280x1024: a mapping loop (render, L1 image loss against a target frame, backward with the
    densification statistics folded in, Adam step) with densify_and_prune every 4 iterations.  P and the instance
    count jump at every densification: the forward's optimistic tail must relaunch, the workspace pool must serve
    the new sizes, and a captured step must notice the overflow and recapture."""
    from fsgs_b200 import GraphedStep, densify
    render, model, sc, poses, pc = _setup(2_000_000, 1280, 1024, m=2.0, seed=0)
    with torch.no_grad():
        target = render.render(poses, 0, pc, gs_grad=False, cam_grad=False)["render"].clone()
        pc.params["_features_dc"].add_(0.2 * torch.randn_like(pc.params["_features_dc"]))
        pc.params["_xyz"].add_(2e-3 * torch.randn_like(pc.params["_xyz"]))
    densify.training_setup(pc, LR)
    pc.variables['scene_radius'] = torch.tensor(0.75, device=DEV)
    pc.fold_densification_stats = True
    history = []

    def iteration():
        pc.optimizer.zero_grad(set_to_none=True)
        out = render.render(poses, 0, pc, gs_grad=True, cam_grad=False)
        loss = (out["render"] - target).abs().mean()
        loss.backward()
        pc.optimizer.step()
        return loss.detach(), out["num_rendered"]

    for it in range(1, 13):
        loss, nr = iteration()
        history.append((it, pc.params["_xyz"].shape[0], int(nr[0]), float(loss)))
        if it % 4 == 0:
            before = pc.params["_xyz"].shape[0]
            assert float(pc.variables['denom'].max()) == 4.0
            densify.densify_and_prune(pc, 2e-6, 0.005, None)
            after = pc.params["_xyz"].shape[0]
            assert after != before and pc.variables['denom'].shape[0] == after
            assert float(pc.variables['denom'].max()) == 0.0
    counts = [h[1] for h in history]
    assert len(set(counts)) == 3, counts                    # P changed at both densifications inside the loop
    assert all(torch.isfinite(torch.tensor(h[3])) for h in history)
    assert history[-1][3] < history[0][3]
    # a captured step across a densification: overflow is detected, recapture repairs it
    state = {"n": 0}

    def graph_step():
        pc.optimizer.zero_grad(set_to_none=True)
        out = render.render(poses, 0, pc, gs_grad=True, cam_grad=False)
        loss = (out["render"] - target).abs().mean()
        loss.backward()
        return loss.detach()

    gs = GraphedStep(graph_step, warmup=2, headroom=1.05)
    l0 = float(gs.replay())
    assert not gs.overflowed()
    with torch.no_grad():
        pc.params["_scaling"].add_(0.35)                     # every splat 1.4x larger: ~2x the instances, same tensors
    gs.replay()
    torch.cuda.synchronize()
    assert gs.overflowed()
    gs.recapture()
    l1 = float(gs.replay())
    assert not gs.overflowed() and l1 == l1 and l1 != l0
    assert abs(l1 - float(graph_step())) <= 1e-5 * abs(l1)
    gs.release()
    report("config4 size mapping loop with densification", history=[list(h) for h in history],
           note="(iteration, Gaussians, tile instances, L1 loss); densify_and_prune after iterations 4, 8, 12")
