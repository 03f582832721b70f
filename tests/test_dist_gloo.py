"""Host logic of the frame data-parallel path, world_size 2 over gloo on CPU (no GPU needed):
frame sharding, the sum all-reduce of the six Gaussian gradient tensors (bucketed and not),
ranks without gradients, and the densification-statistics reductions."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fsgs_b200 import dist as fd

SHAPES = {"_xyz": (7, 3), "_features_dc": (7, 1, 3), "_features_rest": (7, 15, 3), "_opacity": (7, 1),
          "_scaling": (7, 3), "_rotation": (7, 4)}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _grad(rank, key):
    g = torch.Generator().manual_seed(1000 * rank + list(SHAPES).index(key))
    return torch.randn(*SHAPES[key], generator=g)


def _worker(rank, world, port, bucket, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = {k: torch.zeros(*s, requires_grad=True) for k, s in SHAPES.items()}
        for k, p in params.items():
            if not (rank == 1 and k == "_opacity"):          # rank 1 has no grad for one tensor
                p.grad = _grad(rank, k)
        nbytes = fd.allreduce_gaussian_grads(params, bucket=bucket)
        assert nbytes == 7 * 59 * 4
        for k, p in params.items():
            want = sum(_grad(r, k) for r in range(world) if not (r == 1 and k == "_opacity"))
            assert torch.allclose(p.grad, want, atol=1e-6), k
        # gradients handed out as views of one flat buffer (what the fused backward does): detected and
        # reduced with a single collective
        flat = torch.zeros(7 * 59)
        views, off = {}, 0
        for k in ("_rotation", "_features_rest", "_xyz", "_features_dc", "_scaling", "_opacity"):
            n = int(torch.tensor(SHAPES[k]).prod())
            views[k] = flat[off:off + n].view(*SHAPES[k])
            off += n
        p2 = {k: torch.zeros(*s, requires_grad=True) for k, s in SHAPES.items()}
        for k, p in p2.items():
            views[k].copy_(_grad(rank, k))
            p.grad = views[k]
        assert fd._shared_flat([p2[k].grad for k in fd.PARAM_KEYS]) is not None
        fd.allreduce_gaussian_grads(p2, bucket=bucket)
        for k, p in p2.items():
            assert torch.allclose(p.grad, sum(_grad(r, k) for r in range(world)), atol=1e-6), k
        var = {"xyz_gradient_accum": torch.full((7, 1), float(rank + 1)), "denom": torch.ones(7, 1),
               "max_radii2D": torch.arange(7.0) * (rank + 1)}
        fd.allreduce_densification_stats(var)
        assert torch.equal(var["xyz_gradient_accum"], torch.full((7, 1), 3.0))
        assert torch.equal(var["denom"], torch.full((7, 1), 2.0))
        assert torch.equal(var["max_radii2D"], torch.arange(7.0) * 2)
        # frame-DP step with a stand-in renderer (the CUDA renderer needs a GPU): loss is a function of
        # the shared parameter and of the frame id, so the reduced gradient must equal the serial sum
        pc = type("PC", (), {})()
        pc.params = {k: torch.ones(*s, requires_grad=True) for k, s in SHAPES.items()}
        frames = fd.shard_frames(list(range(5)), world, rank)
        render = lambda poses, f, pc, gs_grad, cam_grad: {"x": sum((p * (f + 1)).sum() for p in pc.params.values())}
        loss, pkgs = fd.dp_render_step(render, None, pc, frames, lambda f, pkg: pkg["x"])
        assert len(pkgs) == len(frames)
        for k, p in pc.params.items():
            assert torch.allclose(p.grad, torch.full(SHAPES[k], float(sum(f + 1 for f in range(5))))), k
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket", [True, False])
def test_frame_dp_host_logic_world2_gloo(bucket):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bucket, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def _compact_worker(rank, world, port, sh_deg, q):
    """The frame-parallel exchange protocol (fsgs_b200.dist.enable_frame_parallel) on the float64 oracle:
    rank r renders frame r; all-reducing the 14-float compact gradient and expanding the SH gradients from the
    reduced masked colour gradient must equal all-reducing the full 59-float gradient."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fsgs_b200.synth import frame_pose_params, make_scene
        from oracle import raster_oracle as ro
        from oracle import render_oracle as R
        dt = torch.float64
        sc = make_scene(300, 96, 64, size_mult=2.0, seed=5)
        params = {k: v.to(dt).requires_grad_(True) for k, v in sc.params.items()}
        qf, tf = frame_pose_params(rank)
        r = torch.tensor(qf, dtype=dt, requires_grad=True)
        t = torch.tensor(tf, dtype=dt, requires_grad=True)
        out = R.render(params, r, t, sc.camera, sh_deg, sc.camera.campos, True, True, backend="c")
        gen = torch.Generator().manual_seed(77 + rank)          # every frame has its own target
        G = torch.randn(4, sc.height, sc.width, generator=gen, dtype=dt)
        ((out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()).backward()
        g = {k: p.grad.clone() for k, p in params.items()}
        assert g["_features_dc"].abs().max() > 0
        gc = g["_features_dc"][:, 0, :] / ro.SH_C0                  # dL/df_dc = basis_0 * gc, basis_0 = SH_C0
        compact = torch.cat([g[k].reshape(sc.P, w) if k != "gc" else gc for k, w in fd.COMPACT_LAYOUT], dim=1)
        assert compact.shape == (sc.P, 14)
        dist.all_reduce(compact)
        for k in g:
            dist.all_reduce(g[k])
        basis = lambda deg, dirs: ro.eval_sh(deg, torch.eye(16, dtype=dt).expand(dirs.shape[0], 16, 16), dirs)
        f_dc, f_rest = fd.expand_sh_grads_reference(compact[:, 11:14], params["_xyz"].detach(), sc.camera.campos.to(dt),
                                                    sh_deg, basis)
        off = 0
        for k, w in fd.COMPACT_LAYOUT:
            if k != "gc":
                assert torch.equal(compact[:, off:off + w].reshape(g[k].shape), g[k]), k
            off += w
        scale = g["_features_rest"].abs().max()
        assert (f_dc - g["_features_dc"]).abs().max() <= 1e-12 * scale
        assert (f_rest - g["_features_rest"]).abs().max() <= 1e-12 * scale, (f_rest - g["_features_rest"]).abs().max()
        if sh_deg < 3:
            assert g["_features_rest"][:, (sh_deg + 1) ** 2 - 1:, :].abs().max() == 0
        # dp_render_step refuses uneven frame counts once the exchange is folded into the backward
        from fsgs_b200 import frame_render
        frame_render.set_grad_reducer(lambda flat: dist.all_reduce(flat))
        try:
            pc = type("PC", (), {})()
            pc.params = {"_xyz": torch.ones(3, 3, requires_grad=True)}
            render = lambda poses, f, pc, gs_grad, cam_grad: {"x": pc.params["_xyz"].sum()}
            try:
                fd.dp_render_step(render, None, pc, fd.shard_frames([0, 1, 2], world, rank), lambda f, pkg: pkg["x"])
                raise AssertionError("uneven frame counts must be rejected")
            except ValueError:
                pass
            fd.dp_render_step(render, None, pc, fd.shard_frames([0, 1], world, rank), lambda f, pkg: pkg["x"])
            assert torch.equal(pc.params["_xyz"].grad, torch.ones(3, 3))      # no second reduction on top
        finally:
            frame_render.set_grad_reducer(None)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()[-600:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sh_deg", [3, 1])
def test_compact_gradient_exchange_world2_gloo(sh_deg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_compact_worker, args=(r, 2, port, sh_deg, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_shard_frames():
    assert fd.shard_frames(list(range(8)), 8, 3) == [3]
    assert fd.shard_frames(list(range(5)), 2, 0) == [0, 2, 4] and fd.shard_frames(list(range(5)), 2, 1) == [1, 3]
    assert fd.shard_frames([], 4, 1) == []
    with pytest.raises(ValueError):
        fd.shard_frames([0, 1], 2, 2)


def test_single_process_is_a_noop():
    params = {k: torch.zeros(*s, requires_grad=True) for k, s in SHAPES.items()}
    for k, p in params.items():
        p.grad = torch.ones_like(p)
    assert fd.allreduce_gaussian_grads(params) == 7 * 59 * 4
    assert all(torch.equal(p.grad, torch.ones_like(p)) for p in params.values())
