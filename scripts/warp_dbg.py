import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rs_face_detection_b200 import Context
import bench
ctx = Context(0)
wk = bench.Workload(ctx, "c2", 0, 0)
print("--- the bench step (preprocess -> detect -> align_detections), 3 times", flush=True)
for _ in range(3):
    wk.step()
    ctx.synchronize()
counts, det, lmk = ctx.detect_fetch(wk.B)
total = int(counts.sum())
fidx = np.repeat(np.arange(wk.B, dtype=np.int32), counts)
for F in [125, 998, 3992]:
    reps = -(-F // total)
    l = np.tile(lmk[:total], (reps, 1))[:F].astype(np.float32)
    fi = np.tile(fidx, reps)[:F].astype(np.int32)
    ld, fd_ = ctx.to_device(l), ctx.to_device(fi)
    crops = ctx.alloc(F * 112 * 112 * 3)
    print("--- align_batch F=%d after a preprocess (cold L2), twice; then warm" % F, flush=True)
    for _ in range(2):
        ctx.preprocess_batch(wk.frames_l, wk.tensor_t)
        ctx.align_batch(wk.frames_l, ld, fd_, F, crops)
    ctx.align_batch(wk.frames_l, ld, fd_, F, crops)
    ctx.synchronize()
    crops.free()
