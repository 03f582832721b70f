"""GPU timeline of one fused render step: every kernel/memcpy with its start, duration and the idle gap before it.
Run on the GPU box:  python tools/gap_profile.py [P] [m]"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from fsgs_b200 import frame_render as render  # noqa: E402
from fsgs_b200 import model  # noqa: E402
from fsgs_b200.synth import make_scene  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
m = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
sc = make_scene(P, 1280, 1024, size_mult=m, seed=0)
poses, pc = model.scene_to_device(sc, "cuda")
G = torch.cat([sc.grads_out["G_rgb"], sc.grads_out["G_dep"][None]]).cuda()


def step():
    pc.zero_grad()
    poses.pose_param_net.zero_grad(set_to_none=True)
    out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
    loss = (out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# the third step: between the 2nd and 3rd occurrence of k_pose_forward
starts = [i for i, e in enumerate(ev) if "k_pose_forward" in e["name"]]
a, b = starts[2], starts[3]
t0 = ev[a]["ts"]
prev_end = None
busy = 0.0
for e in ev[a:b]:
    gap = 0.0 if prev_end is None else e["ts"] - prev_end
    print(f"{e['ts'] - t0:9.1f} us  dur {e['dur']:8.1f}  gap {gap:7.1f}  {e['name'][:70]}")
    prev_end = max(prev_end or 0, e["ts"] + e["dur"])
    busy += e["dur"]
span = ev[b]["ts"] - t0
print(f"step span {span:.1f} us, busy {busy:.1f} us, idle {span - busy:.1f} us")
