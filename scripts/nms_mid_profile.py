"""fd_nms_device on one problem of a few thousand boxes: median host time (enqueue + synchronize) of 50 calls and the ABI's
per-launch CUDA-event times, for the single-CTA path (FD_NMS_CROSS large), the mid path and the spatial path (FD_NMS_MID_CAP=0).
Run once per configuration: the switches are read at first use."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rs_face_detection_b200 import Context
from rs_face_detection_b200.utils import synth
ctx = Context(0)
out = {}
for N in [int(x) for x in os.environ.get("SIZES", "1500,2000,2560,3000,4096,6000,8192").split(",")]:
    dets = synth.make_crowd_boxes(N, seed=42, n_faces=max(1, N // 20))
    d = ctx.to_device(dets)
    keep, num = ctx.alloc(4 * N), ctx.alloc(16)
    for _ in range(5):
        ctx.nms_device(d, N, 0.4, keep, num)
    ctx.synchronize()
    ts = []
    for _ in range(50):
        t0 = time.perf_counter()
        ctx.nms_device(d, N, 0.4, keep, num)
        ctx.synchronize()
        ts.append(time.perf_counter() - t0)
    ctx.profile(True)
    for _ in range(10):
        ctx.nms_device(d, N, 0.4, keep, num)
    ctx.synchronize()
    prof = ctx.profile_fetch()
    ctx.profile(False)
    out[N] = {"us": round(1e6 * float(np.median(ts)), 1), "kept": int(num.download((2,), np.int32)[0]),
              "kernels_us": {k: round(v[1] / v[0], 1) for k, v in prof.items()}}
    print(N, out[N], flush=True)
print(json.dumps({"cross": os.environ.get("FD_NMS_CROSS"), "mid_cap": os.environ.get("FD_NMS_MID_CAP"), "by_size": out}))
