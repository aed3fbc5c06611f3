"""One fd_nms_device problem of N crowd boxes, a few calls (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rs_face_detection_b200 import Context
from rs_face_detection_b200.utils import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = Context(0)
dets = synth.make_crowd_boxes(N, seed=42, n_faces=max(1, N // 20))
d = ctx.to_device(dets)
keep, num = ctx.alloc(4 * N), ctx.alloc(16)
for _ in range(4):
    ctx.nms_device(d, N, 0.4, keep, num)
ctx.synchronize()
print("kept", int(num.download((2,), np.int32)[0]))
