#!/usr/bin/env python
"""Config 3 (BASELINE.json configs[2]): run the UNMODIFIED Free-SurGS driver -- the reference's own
``train.py`` (progressive tracking + mapping, then global refinement), ``gaussian_renderer.render``, ``GaussianModel``
and ``PoseModel`` -- on top of this library and report PSNR + ATE as the reference computes them
(train.py:446-515: utils.general_utils.rgb_evaluation, utils.geometry_utils.align_pose).

    python tools/run_config3.py --ref <Free-SurGS checkout> --backend fsgs|fsgs-fused|oracle [--frames 8 --size 320x256]

  --backend fsgs        the reference's render() as is: PyTorch pre-processing + two GaussianRasterizer calls, the
                        package name ``diff_gaussian_rasterization`` resolving to free-surgs_b200/ (integration level 1)
  --backend fsgs-fused  additionally ``gaussian_renderer.render`` is rebound to ``fsgs_b200.render`` before train.py
                        imports it (integration level 2: the one-import change of INTEGRATION.md, done from outside)
  --backend oracle      ``diff_gaussian_rasterization`` resolves to oracle/ref_boundary/ (plain-C float32 CPU oracle):
                        the comparison run of SURVEY.md 8d config 3

The reference checkout is NOT part of this repository: ``--ref`` defaults to baseline/_ref/Free-SurGS (staged from
/root/reference by __graft_entry__.build(), git-ignored) and then /root/reference.  Nothing of the reference is
modified: train.py is executed with runpy under its own ``__main__`` guard; absent third-party imports are answered
by tools/ref_shims; the dataset is the synthetic SCARED-format sequence of tools/make_scared_synth.py.
Prints ONE JSON line (also appended to --report if given).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import runpy
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference(explicit=None):
    for cand in (explicit, os.environ.get("FSGS_REFERENCE_DIR"), os.path.join(ROOT, "baseline", "_ref", "Free-SurGS"),
                 "/root/reference"):
        if cand and os.path.exists(os.path.join(cand, "train.py")) and os.path.isdir(os.path.join(cand, "gaussian_renderer")):
            return os.path.abspath(cand)
    return None


def _cpu_redirect():
    """--cpu (plumbing tests in a container without a GPU; oracle backend only): the reference hard-codes 'cuda'
    (``.cuda()``, ``device="cuda"``, CUDA events); send all of it to the host."""
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    for name in ("zeros", "ones", "eye", "tensor", "zeros_like", "ones_like", "empty", "arange", "full", "randint",
                 "rand", "randn", "linspace"):
        orig = getattr(torch, name)

        def wrapped(*a, __orig=orig, **k):
            if "device" in k and str(k["device"]).startswith("cuda"):
                k["device"] = "cpu"
            return __orig(*a, **k)
        setattr(torch, name, wrapped)
    orig_to = torch.nn.Module.to

    def module_to(self, *a, **k):
        a = tuple("cpu" if isinstance(x, str) and x.startswith("cuda") else x for x in a)
        if "device" in k and str(k["device"]).startswith("cuda"):
            k["device"] = "cpu"
        return orig_to(self, *a, **k)
    torch.nn.Module.to = module_to

    class _Event:
        def __init__(self, *a, **k):
            pass

        def record(self, *a, **k):
            pass
    torch.cuda.Event = _Event
    torch.cuda.empty_cache = lambda: None


def run(ref, backend, data_root, model_path, iterations, quiet=True):
    """Execute the reference's train.py in this process.  Returns (globals of train.py, wall seconds)."""
    import torch
    pkg = os.path.join(ROOT, "free-surgs_b200")
    boundary = os.path.join(ROOT, "oracle", "ref_boundary") if backend == "oracle" else pkg
    # import order: the reference's own packages first, then the boundary that answers `diff_gaussian_rasterization`,
    # then free-surgs_b200 (simple_knn stand-in, fsgs_b200), then the shims for absent third-party modules
    for p in (pkg, boundary, ref):
        while p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_shims
    shimmed = ref_shims.install()
    sys.path.remove(os.path.join(ROOT, "tools"))
    import diff_gaussian_rasterization as dgr
    assert os.path.abspath(dgr.__file__).startswith(boundary), (dgr.__file__, boundary)
    if backend == "fsgs-fused":
        import gaussian_renderer            # the reference's package
        import fsgs_b200
        gaussian_renderer.render = fsgs_b200.render
    argv = ["train.py", "-s", data_root, "--model_path", model_path, "--iterations", str(iterations), "--quiet"]
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = argv
    os.chdir(model_path)
    t0 = time.time()
    try:
        sink = io.StringIO()
        with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
            g = runpy.run_path(os.path.join(ref, "train.py"), run_name="__main__")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return g, time.time() - t0, shimmed


def evaluate(g):
    """PSNR / SSIM on the held-out frames and RPE / ATE of the estimated trajectory, with the reference's own
    functions and the frame selection of FreeSurGS.validation / eval_pose (train.py:446-515)."""
    import numpy as np
    import torch
    slam = g["slam"]
    render = g["render"]
    from utils.general_utils import rgb_evaluation
    from utils.geometry_utils import align_pose
    pred, gt = [], []
    with torch.no_grad():
        for index in slam.poses.i_test:
            pkg = render(slam.poses, index, slam.gaussians, gs_grad=False, cam_grad=False)
            pred.append(torch.clamp(pkg["render"].detach().cpu(), 0.0, 1.0))
            gt.append(torch.clamp(slam.poses.record_data['colors'][index], 0.0, 1.0))
    with contextlib.redirect_stdout(io.StringIO()):
        psnr, ssim, _lp = rgb_evaluation(np.stack(gt, 0), np.stack(pred, 0))
        train_pred, train_gt = [], []
        with torch.no_grad():
            for index in slam.poses.i_train:
                pkg = render(slam.poses, index, slam.gaussians, gs_grad=False, cam_grad=False)
                train_pred.append(torch.clamp(pkg["render"].detach().cpu(), 0.0, 1.0))
                train_gt.append(torch.clamp(slam.poses.record_data['colors'][index], 0.0, 1.0))
        psnr_tr, ssim_tr, _ = rgb_evaluation(np.stack(train_gt, 0), np.stack(train_pred, 0))
        data_ind = slam.poses.record_data['data_ind']
        metrics = np.zeros(3)
        for i, (_key, gt_poses) in enumerate(slam.poses.record_data['gt_poses'].items()):
            pred_w2c = torch.from_numpy(slam.poses.record_data['pred_w2c'])[data_ind[i]:data_ind[i + 1]]
            _, m = align_pose(pred_w2c, gt_poses)
            metrics += np.array(m) * slam.poses.record_data['weights'][i]
    viewer = viewer_check(g)
    return {"viewer": viewer,
            "psnr_test": float(psnr), "ssim_test": float(ssim), "psnr_train": float(psnr_tr), "ssim_train": float(ssim_tr),
            "rpe_trans": float(metrics[0]), "rpe_rot_deg": float(metrics[1]), "ate": float(metrics[2]),
            "n_gaussians": int(slam.gaussians.params['_xyz'].shape[0]), "n_frames": int(slam.poses.num_cams),
            "iterations_run": int(slam.iteration)}


def viewer_check(g, n_frames=3):
    """The reference's viewer path (train.py:124-152): ``FreeSurGS.render_fn`` builds a 2048x1200 camera with
    ``setup_camera(I, visualize_data)`` and calls ``render_custom`` (gaussian_renderer/__init__.py:112-135) from the
    nerfview THREAD while the training thread keeps rendering.  Here: a second Python thread calls the unmodified
    ``render_fn`` ``n_frames`` times while this thread runs mapping iterations on the same model."""
    import threading
    import types
    import numpy as np
    import torch
    slam = g["slam"]
    res = {"frames": 0, "errors": []}

    def worker():
        try:
            for k in range(n_frames):
                c2w = np.eye(4)
                c2w[:3, 3] = [0.01 * k, 0.0, 0.0]
                state = types.SimpleNamespace(fov=np.float64(0.9), c2w=c2w)
                img = slam.render_fn(state, (2048, 1200))
                assert img.shape == (1200, 2048, 3) and img.dtype == np.uint8, (img.shape, img.dtype)
                res["frames"] += 1
                res["mean_intensity"] = float(img.mean())
        except Exception as exc:  # noqa: BLE001
            res["errors"].append(repr(exc)[:300])

    th = threading.Thread(target=worker)
    th.start()
    with contextlib.redirect_stdout(io.StringIO()):
        slam.mapping(int(slam.poses.i_train[-1]), mapping_iter=3, progressive=False)
    th.join(timeout=300)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    res["ok"] = res["frames"] == n_frames and not res["errors"]
    res["size"] = "2048x1200"
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=None)
    ap.add_argument("--backend", default="fsgs", choices=["fsgs", "fsgs-fused", "oracle"])
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--size", default="320x256")
    ap.add_argument("--P", type=int, default=20000, help="Gaussians of the ground-truth scene the frames are rendered from")
    ap.add_argument("--iterations", type=int, default=100, help="global_run iterations (train.py --iterations)")
    ap.add_argument("--data", default=None, help="reuse / write the synthetic sequence here (default: a temp dir)")
    ap.add_argument("--renderer", default="fsgs", choices=["fsgs", "oracle"], help="what renders the ground-truth frames")
    ap.add_argument("--report", default=None)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--cpu", action="store_true", help="redirect the reference's hard-coded 'cuda' to the host "
                    "(oracle backend + oracle renderer only: plumbing test without a GPU)")
    a = ap.parse_args()
    if a.cpu:
        if a.backend != "oracle" or a.renderer != "oracle":
            raise SystemExit("--cpu needs --backend oracle --renderer oracle (the CUDA library has no CPU fallback)")
        _cpu_redirect()
    ref = find_reference(a.ref)
    if ref is None:
        print(json.dumps({"config3": "skipped", "why": "no Free-SurGS checkout (baseline/_ref/Free-SurGS or /root/reference)"}))
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_scared_synth import write_sequence
    W, H = (int(x) for x in a.size.split("x"))
    tmp = tempfile.mkdtemp(prefix="fsgs_config3_")
    data_root = a.data or os.path.join(tmp, "data")
    if not os.path.isdir(os.path.join(data_root, "input")):
        write_sequence(data_root, a.frames, W, H, a.P, renderer=a.renderer)
    model_path = os.path.join(tmp, "model")
    os.makedirs(model_path, exist_ok=True)
    g, wall, shimmed = run(ref, a.backend, data_root, model_path, a.iterations, quiet=not a.verbose)
    res = evaluate(g)
    res.update({"config3": "ok", "backend": a.backend, "size": a.size, "global_iterations": a.iterations,
                "wall_s": round(wall, 1), "reference": ref, "shimmed_modules": shimmed,
                "checkpoints": sorted(f for f in os.listdir(model_path) if f.endswith(".pth"))})
    line = json.dumps(res)
    print(line)
    if a.report:
        with open(a.report, "a") as f:
            f.write(line + "\n")
    shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
