"""Worker of tests/test_gpu_multigpu.py (one process per GPU under torchrun): frame-parallel fused backward with the
NVLink exchange (fsgs_exchange_rows on a symmetric buffer) against ncclAllReduce of the same rows and against the sum
of the frames' full gradients computed locally."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fsgs_b200 import dist as fd  # noqa: E402
from fsgs_b200 import frame_render as render  # noqa: E402
from fsgs_b200 import model  # noqa: E402
from fsgs_b200.synth import frame_pose_params, make_scene  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 20003
sc = make_scene(P, 320, 256, size_mult=2.0, seed=3)
G = [torch.randn(4, sc.height, sc.width, generator=torch.Generator().manual_seed(10 + k)).to(dev) for k in range(world)]


def run(frame):
    poses, pc = model.scene_to_device(sc, dev)
    poses.set_pose(0, *frame_pose_params(frame))
    out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
    ((out["render"] * G[frame][:3]).sum() + (out["render_dep"] * G[frame][3]).sum()).backward()
    torch.cuda.synchronize()
    return {k: p.grad.detach().clone() for k, p in pc.params.items()}, poses.pose_param_net.r.grad.clone()


rel = lambda a, b: float((a - b).norm() / b.norm().clamp(min=1e-30))
res = {"rank": rank, "world": world}
local_full = [run(f) for f in range(world)]                       # no exchange: every frame rendered here
want = {k: sum(local_full[f][0][k] for f in range(world)) for k in local_full[0][0]}
got = {}
# nvlink = the library's default choice for this rank count; the two forms of its exchange are also forced in turn:
# two-shot (fsgs_exchange_rows + expansion), one-shot (rank sum folded into the expansion kernel) and pull-gather
# (reduce-scatter, the all-gather half riding on the expansion kernel)
for transport, form in (("nccl", None), ("nvlink", None), ("nvlink_two_shot", "two_shot"), ("nvlink_one_shot", "one_shot"),
                        ("nvlink_pull_gather", "pull_gather")):
    if form is None:
        os.environ.pop("FSGS_EXCHANGE_FORM", None)
    else:
        os.environ["FSGS_EXCHANGE_FORM"] = form
    fd.enable_frame_parallel(check_cam_center=torch.zeros(3, device=dev), exchange=transport.split("_")[0])
    got[transport] = run(rank)
    fd.disable_frame_parallel()
    res[transport + "_vs_local_sum"] = max(rel(got[transport][0][k], want[k]) for k in want)
    res[transport + "_pose_local"] = rel(got[transport][1], local_full[rank][1])
os.environ.pop("FSGS_EXCHANGE_FORM", None)
res["nvlink_vs_nccl"] = max(rel(got["nvlink"][0][k], got["nccl"][0][k]) for k in want)
# every rank must hold bit-identical summed gradients (replicas must not drift apart)
res["bit_identical_across_ranks"] = True
for transport in ("nvlink", "nvlink_two_shot", "nvlink_one_shot", "nvlink_pull_gather"):
    flat = torch.cat([got[transport][0][k].reshape(-1) for k in sorted(want)])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    res["bit_identical_across_ranks"] = res["bit_identical_across_ranks"] and bool(torch.equal(ref, flat))
print("MGPU " + json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
