"""host enqueue time vs device time of fd_nms_device (profiling helper)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rs_face_detection_b200 import Context
from rs_face_detection_b200.utils import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ctx = Context(0)
ext = torch.cuda.ExternalStream(ctx.stream(), device=0)
dets = synth.make_crowd_boxes(N, seed=42, n_faces=max(1, N // 20))
d = ctx.to_device(dets)
keep, num = ctx.alloc(4 * N), ctx.alloc(16)
for _ in range(5):
    ctx.nms_device(d, N, 0.4, keep, num)
ctx.synchronize()
host, dev = [], []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(ext)
    t0 = time.perf_counter()
    ctx.nms_device(d, N, 0.4, keep, num)
    t1 = time.perf_counter()
    b.record(ext)
    ctx.synchronize()
    host.append((t1 - t0) * 1e6); dev.append(a.elapsed_time(b) * 1e3)
print("N=%d host enqueue median %.0f us, device span median %.0f us" % (N, np.median(host), np.median(dev)))
# back-to-back: 10 calls enqueued without sync -> device time per call when the queue is full
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(ext)
for _ in range(10):
    ctx.nms_device(d, N, 0.4, keep, num)
b.record(ext)
ctx.synchronize()
print("back-to-back per call %.0f us" % (a.elapsed_time(b) * 1e3 / 10))
