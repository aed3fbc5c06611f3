"""FD_NMS_DBG=1 python scripts/nms_dbg.py : per-CTA stage times of the batched small-path NMS on the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rs_face_detection_b200 import Context
from rs_face_detection_b200.utils import synth
ctx = Context(0)
heads, _ = synth.make_heads(64, seed=3000, n_faces=20, content_hw=(360, 640))
devs = [ctx.to_device(h) for h in heads]
for _ in range(3):
    ctx.detect_batch(devs, 64, np.full(64, 1 / 3, np.float32), 0.7, 0.4)
    ctx.synchronize()
    sys.stderr.write("----\n")
