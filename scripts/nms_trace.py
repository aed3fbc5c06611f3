import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rs_face_detection_b200 import Context
from rs_face_detection_b200.utils import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
os.environ.pop("FD_TRACE", None)
warm = Context(0)
dets = synth.make_crowd_boxes(N, seed=42, n_faces=max(1, N // 20))
os.environ["FD_TRACE"] = "1"
ctx = Context(0)
d = ctx.to_device(dets)
keep, num = ctx.alloc(4 * N), ctx.alloc(16)
for _ in range(3):
    ctx.nms_device(d, N, 0.4, keep, num)
    ctx.lib.fd_ctx_synchronize(ctx.handle) if _ < 2 else None
import ctypes
ctx.synchronize()
