"""BASELINE.json configs[0] through the Python-free driver ("a CPU-side driver ... without Python"):
tools/c_driver/config1_driver.cpp links libfsgs_raster.so directly and runs LearnPose forward -> fused render forward
-> backward -> LearnPose backward with cudaMalloc'd buffers and its own allocation callbacks.  This test only writes
its input file (scene + the float64 oracle's expected planes / gradients) and reads its verdict."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from parity import fragile_mask, report  # noqa: E402

from fsgs_b200.synth import make_scene  # noqa: E402
from oracle import render_oracle as R  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "free-surgs_b200", "fsgs_b200", "fsgs_config1_driver")


def test_driver_binary_is_built_and_refuses_to_run_without_a_case():
    assert os.path.exists(DRIVER), "run __graft_entry__.build() (make -C free-surgs_b200/csrc)"
    p = subprocess.run([DRIVER], capture_output=True, text=True)
    assert p.returncode == 64 and "usage" in p.stderr


@pytest.mark.gpu
def test_config1_through_the_c_driver(tmp_path):
    P, W, H = 10000, 640, 512
    sc = make_scene(P, W, H, size_mult=2.0, seed=0)
    dt = torch.float64
    params = {k: v.detach().clone().to(dt).requires_grad_(True) for k, v in sc.params.items()}
    r, t = sc.pose_q.clone().to(dt).requires_grad_(True), sc.pose_t.clone().to(dt).requires_grad_(True)
    out = R.render(params, r, t, sc.camera, 3, sc.camera.campos, True, True, want_aux=True, backend="c")
    mask = fragile_mask(out["_aux"], H, W)
    G = torch.cat([sc.grads_out["G_rgb"], sc.grads_out["G_dep"][None]]) * (~mask).float()[None]
    ((out["render"] * G[:3].to(dt)).sum() + (out["render_dep"] * G[3].to(dt)).sum()).backward()
    planes = torch.cat([out["render"], out["_depth_sil"]], 0).detach()
    cam = sc.camera
    f32 = lambda x: np.ascontiguousarray(torch.as_tensor(x).detach().float().numpy().reshape(-1))
    path = tmp_path / "config1.bin"
    with open(path, "wb") as f:
        f.write(b"FSGSC1\0\0")
        f.write(np.array([P, W, H, 3], dtype=np.int32).tobytes())
        f.write(np.array([cam.tanfovx, cam.tanfovy], dtype=np.float32).tobytes())
        for a in (cam.bg, sc.pose_q, sc.pose_t, cam.campos, cam.viewmatrix, cam.projmatrix, sc.params["_xyz"],
                  sc.params["_features_dc"], sc.params["_features_rest"], sc.params["_opacity"], sc.params["_scaling"],
                  sc.params["_rotation"], G, planes, mask.float(), params["_xyz"].grad, params["_features_dc"].grad,
                  params["_features_rest"].grad, params["_opacity"].grad, params["_scaling"].grad,
                  params["_rotation"].grad, r.grad, t.grad):
            f.write(f32(a).tobytes())
    p = subprocess.run([DRIVER, str(path)], capture_output=True, text=True, timeout=300)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert lines, p.stdout + p.stderr
    res = json.loads(lines[-1])
    report("config1 through the Python-free C driver vs float64 C oracle", **res)
    assert p.returncode == 0 and res["ok"], (p.returncode, res, p.stderr[-500:])
    assert res["tile_instances_rect"] == int(out["_num_rendered"]) and res["pixels_above_gate"] == 0
    assert max(res["grad_rel_err"].values()) <= 1e-4
