"""PCIe / host-memory probe: pinned host -> device bandwidth, one process per GPU under torchrun (N = 1, 2, 4, 8), all ranks
copying at the same time — the ceiling of the e2e leg of bench.py on this box.  Rank 0 prints one JSON line:
per-rank and aggregate GB/s for H2D alone, and H2D with a concurrent D2H stream.

  python scripts/h2d_probe.py                                                   # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/h2d_probe.py
"""
import json, os, time
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
h = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
d = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
for t in h:
    t.fill_(1)
st = [torch.cuda.Stream() for _ in range(2)]


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(reps, with_d2h):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(st[0]):
            d[0].copy_(h[0], non_blocking=True)
        if with_d2h:
            with torch.cuda.stream(st[1]):
                h[1].copy_(d[1], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n / dt / 1e9


def gather(v):
    if world == 1:
        return [v]
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


run(2, False)
a = gather(run(12, False))
b = gather(run(12, True))
if rank == 0:
    print(json.dumps({"n_gpus": world, "host_cores": len(os.sched_getaffinity(0)), "h2d_gbs_per_rank": [round(x, 2) for x in a],
                      "h2d_gbs_aggregate": round(sum(a), 1), "h2d_with_d2h_gbs_per_rank": [round(x, 2) for x in b],
                      "h2d_with_d2h_gbs_aggregate": round(sum(b), 1),
                      "note": "every rank copies 256 MB pinned buffers at the same time; per-rank rate = bytes / that rank's wall time"}))
if world > 1:
    dist.destroy_process_group()
