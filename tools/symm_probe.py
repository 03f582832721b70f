"""2-GPU development probe of the NVLink exchange (run under torchrun on the GPU box):
   torchrun --nproc-per-node 2 tools/symm_probe.py
Checks fsgs_exchange_rows (through fsgs_b200.dist._NvlinkExchange) against ncclAllReduce on random rows, prints whether
the fabric offered a multicast address, and times both."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fsgs_b200 import dist as fd  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x = fd._NvlinkExchange()
for P in (1003, 500_000):
    n = P * 14
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    rows = torch.randn(n, generator=g).to(dev)
    want = rows.clone()
    dist.all_reduce(want)
    buf = x.alloc(n, dev)
    buf.copy_(rows)
    x.reduce(buf)
    torch.cuda.synchronize()
    err = float((buf - want).abs().max())
    same = buf.clone()
    dist.broadcast(same, src=0)
    print(f"rank {rank} P {P}: multicast path {'yes' if x.multicast else 'no (peer loads/stores)'} max|nvlink - nccl| {err:.3e} "
          f"bit-identical across ranks: {bool(torch.equal(same, buf))}", flush=True)
    assert err < 1e-5 * world
    for name, fn in (("nvlink", lambda: x.reduce(buf)), ("nccl", lambda: dist.all_reduce(want))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            fn()
        b.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"   {name}: {a.elapsed_time(b) / 50 * 1e3:.1f} us per all-reduce of {n * 4 / 1e6:.1f} MB", flush=True)
# CUDA-graph capture of the exchange (bench.py replays the whole step from a graph)
buf = x.alloc(500_000 * 14, dev)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    x.reduce(buf)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(gr):
        x.reduce(buf)
    buf.fill_(float(rank + 1))
    gr.replay()
    torch.cuda.synchronize()
    print(f"rank {rank}: graph replay ok, value {float(buf[0])} (want {world * (world + 1) / 2})", flush=True)
except Exception as exc:  # noqa: BLE001
    print(f"rank {rank}: graph capture of the exchange failed: {exc!r}", flush=True)
dist.barrier()
dist.destroy_process_group()
