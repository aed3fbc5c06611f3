"""Phase timeline of nms_big_kernel (FD_NMS_DBG=1 prints block 0's globaltimer stamps on stderr) at 100 000 boxes, and the
median host time of fd_nms_device next to the multi-kernel launch sequence (FD_NMS_MULTI_KERNEL=1 in a second process)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rs_face_detection_b200 import Context
from rs_face_detection_b200.utils import synth
ctx = Context(0)
out = {}
SIZES = [int(x) for x in sys.argv[1:]] or [16800, 100000, 340000]
for N in SIZES:
    dets = synth.make_crowd_boxes(N, seed=42, n_faces=max(1, N // 20))
    d = ctx.to_device(dets)
    keep, num = ctx.alloc(4 * N), ctx.alloc(16)
    for _ in range(5):
        ctx.nms_device(d, N, 0.4, keep, num)
    ctx.synchronize()
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        ctx.nms_device(d, N, 0.4, keep, num)
        ctx.synchronize()
        ts.append(time.perf_counter() - t0)
    out[str(N)] = {"us": round(1e6 * float(np.median(ts)), 1), "kept": int(num.download((2,), np.int32)[0])}
print(json.dumps({"multi_kernel": os.environ.get("FD_NMS_MULTI_KERNEL") == "1", "nms_device_us": out}))
