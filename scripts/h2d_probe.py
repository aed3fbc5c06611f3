"""PCIe probe: pinned host -> device bandwidth on this box (one and two streams), for the e2e roofline in DESIGN.md."""
import json, sys, time
import torch
n = 512 << 20
h = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
d = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
st = [torch.cuda.Stream() for _ in range(2)]
def run(k, reps=6, d2h=False):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(k):
            with torch.cuda.stream(st[i]):
                (h[i].copy_(d[i], non_blocking=True) if d2h else d[i].copy_(h[i], non_blocking=True))
    torch.cuda.synchronize()
    return k * reps * n / (time.perf_counter() - t0) / 1e9
run(1, 2)
out = {"h2d_1stream_gbs": run(1), "h2d_2stream_gbs": run(2), "d2h_1stream_gbs": run(1, d2h=True)}
# both directions at once
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(6):
    with torch.cuda.stream(st[0]): d[0].copy_(h[0], non_blocking=True)
    with torch.cuda.stream(st[1]): h[1].copy_(d[1], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["bidir_each_gbs"] = 6 * n / dt / 1e9
print(json.dumps(out))
