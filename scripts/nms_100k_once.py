"""One 100 000-box fd_nms_device problem (BASELINE config 3), repeated a few times: the target of the ncu captures."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rs_face_detection_b200 import Context
from rs_face_detection_b200.utils import synth
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = Context(0)
dets = synth.make_crowd_boxes(100000, seed=42)
d = ctx.to_device(dets)
keep, num = ctx.alloc(4 * len(dets)), ctx.alloc(16)
for _ in range(n_rep):
    ctx.nms_device(d, len(dets), 0.4, keep, num)
    ctx.synchronize()
print("kept", int(num.download((2,), np.int32)[0]), "stats", ctx.nms_last_stats())
