"""Where does the end-to-end overhead of bench.py's e2e loop come from?  (run on the GPU box)
Times the same loop with the L2 flush / the H2D prefetch / the per-step D2H read-back switched off."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
import torch  # noqa: E402

from fsgs_b200 import frame_render as render  # noqa: E402
from fsgs_b200 import model  # noqa: E402
from fsgs_b200.synth import make_scene  # noqa: E402

sc = make_scene(500_000, 1280, 1024, size_mult=2.0, seed=0)
poses, pc = model.scene_to_device(sc, "cuda")
dev = torch.device("cuda", 0)
G_host = torch.empty(4, 1024, 1280).pin_memory()
G_host[:3] = sc.grads_out["G_rgb"]; G_host[3] = sc.grads_out["G_dep"]
G_dev = G_host.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
copy_stream = torch.cuda.Stream(device=dev)
G_in = [torch.empty_like(G_dev) for _ in range(2)]
copied = [torch.cuda.Event() for _ in range(2)]
consumed = [torch.cuda.Event() for _ in range(2)]
result_host = torch.empty(8).pin_memory()
result_ready = torch.cuda.Event()
main = torch.cuda.current_stream(dev)


def step(G):
    pc.zero_grad()
    poses.pose_param_net.zero_grad(set_to_none=True)
    out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
    loss = (out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()
    loss.backward()
    return loss


def run(n, do_flush=True, do_h2d=True, do_d2h=True):
    for e in consumed:
        e.record(main)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record(main)
    if do_h2d:
        copy_stream.wait_event(t0)
        with torch.cuda.stream(copy_stream):
            G_in[0].copy_(G_host, non_blocking=True); copied[0].record(copy_stream)
    for k in range(n):
        b = k & 1
        if do_flush:
            flush.zero_()
        if do_h2d:
            main.wait_event(copied[b])
            if k + 1 < n:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[b ^ 1])
                    G_in[b ^ 1].copy_(G_host, non_blocking=True); copied[b ^ 1].record(copy_stream)
        loss = step(G_in[b] if do_h2d else G_dev)
        consumed[b].record(main)
        if do_d2h:
            packed = torch.cat([loss.detach().reshape(1), poses.pose_param_net.r.grad.reshape(-1),
                                poses.pose_param_net.t.grad.reshape(-1)])
            result_host.copy_(packed, non_blocking=True)
            result_ready.record(main)
            result_ready.synchronize()
            _ = float(result_host[0])
    t1.record(main)
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n, (time.perf_counter() - w0) / n * 1e3


for _ in range(5):
    step(G_dev)
run(5)
for name, kw in (("full e2e", {}), ("no flush", dict(do_flush=False)), ("no H2D", dict(do_h2d=False)),
                 ("no D2H sync", dict(do_d2h=False)), ("no flush/H2D/D2H", dict(do_flush=False, do_h2d=False, do_d2h=False))):
    ev, wall = run(30, **kw)
    print(f"{name:20s} {ev:.3f} ms/step (events)  {wall:.3f} ms/step (wall)")
