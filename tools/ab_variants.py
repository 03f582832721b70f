"""A/B harness for kernel build variants (run on the GPU box).
   python tools/ab_variants.py libA.so libB.so ...   (paths relative to the repo root; '' = the default build)
For each library: runs the bench workload (P=500k, m=2; AB_P / AB_M override; AB_MODE=track = the tracking step
against a frozen model) in a fresh process and prints step ms and per-kernel ms."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SNIPPET = r'''
import os, sys, json
ROOT = %r
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
import torch
from fsgs_b200 import _lib, model, frame_render as render
from fsgs_b200.synth import make_scene
P, m = int(os.environ.get("AB_P", "500000")), float(os.environ.get("AB_M", "2"))
sc = make_scene(P, 1280, 1024, size_mult=m, seed=0)
poses, pc = model.scene_to_device(sc, "cuda")
G = torch.cat([sc.grads_out["G_rgb"], sc.grads_out["G_dep"][None]]).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
TRACK = os.environ.get("AB_MODE", "full") == "track"      # tracking against a frozen model (pose-only backward)
if TRACK:
    for v in pc.params.values():
        v.requires_grad_(False)
def step():
    pc.zero_grad(); poses.pose_param_net.zero_grad(set_to_none=True)
    if TRACK:
        out = render.render(poses, 0, pc, gs_grad=False, cam_grad=True)
        (out["render"] * G[:3]).sum().backward()
        return
    out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
    ((out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()).backward()
for _ in range(5): step()
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, b in ev:
    flush.zero_(); a.record(); step(); b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
_lib.profile_enable(True)
for _ in range(10):
    flush.zero_(); step()
prof = _lib.profile_collect(); _lib.profile_enable(False)
print(json.dumps({"ms_median": ms[len(ms) // 2], "kernel_ms": {k: round(v[0] / v[1], 4) for k, v in prof.items() if v[1]}}))
''' % ROOT

for lib in sys.argv[1:] or [""]:
    env = dict(os.environ)
    if lib:
        env["FSGS_RASTER_LIB"] = os.path.join(ROOT, lib)
    r = subprocess.run([sys.executable, "-c", SNIPPET], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(f"variant [{lib or 'default'}] {line}", flush=True)
