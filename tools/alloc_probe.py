"""Why is torch.empty sometimes slow inside the render step?  (GPU box diagnostic.)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
import torch
from fsgs_b200 import frame_render as render, model
from fsgs_b200.synth import make_scene

sc = make_scene(500_000, 1280, 1024, size_mult=2.0, seed=0)
poses, pc = model.scene_to_device(sc, "cuda")
G = torch.cat([sc.grads_out["G_rgb"], sc.grads_out["G_dep"][None]]).cuda()
log = []
_empty = torch.empty
def timed_empty(*a, **k):
    t0 = time.perf_counter()
    r = _empty(*a, **k)
    log.append((time.perf_counter() - t0, r.numel() * r.element_size()))
    return r
torch.empty = timed_empty

def step():
    pc.zero_grad(); poses.pose_param_net.zero_grad(set_to_none=True)
    out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
    loss = (out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()
    loss.backward()

for _ in range(5): step()
torch.cuda.synchronize()
s0 = torch.cuda.memory_stats()
log.clear()
times = []
for _ in range(20):
    t0 = time.perf_counter(); step(); times.append(time.perf_counter() - t0)
torch.cuda.synchronize()
s1 = torch.cuda.memory_stats()
print("step host ms:", [round(t * 1e3, 2) for t in times])
for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "num_sync_all_streams"):
    print(k, s1.get(k, 0) - s0.get(k, 0))
print("reserved GB", s1["reserved_bytes.all.current"] / 1e9, "allocated GB", s1["allocated_bytes.all.current"] / 1e9)
slow = sorted(log, reverse=True)[:12]
print("slowest empties (ms, MB):", [(round(a * 1e3, 3), round(b / 1e6, 2)) for a, b in slow])
print("total empty ms/step", sum(a for a, _ in log) / 20 * 1e3)
