"""warp_fixed_kernel time vs number of faces (fixed cost vs per-face cost); faces = the bench's C2 detections, replicated."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rs_face_detection_b200 import Context
import bench

ctx = Context(0)
wk = bench.Workload(ctx, "c2", 0, 0)
wk.step()
counts, det, lmk = ctx.detect_fetch(wk.B)
total = int(counts.sum())
fidx = np.repeat(np.arange(wk.B, dtype=np.int32), counts)
out = {}
for F in [125, 250, 500, 998, 1996, 3992, 7984]:
    reps = -(-F // total)
    l = np.tile(lmk[:total], (reps, 1))[:F].astype(np.float32)
    fi = np.tile(fidx, reps)[:F].astype(np.int32)
    ld, fd_ = ctx.to_device(l), ctx.to_device(fi)
    crops = ctx.alloc(F * 112 * 112 * 3)
    for _ in range(3):
        ctx.align_batch(wk.frames_l, ld, fd_, F, crops)
    ctx.profile(True)
    for _ in range(20):
        ctx.preprocess_batch(wk.frames_l, wk.tensor_t)      # evicts the frames' warp rows from L2 like the real step
        ctx.align_batch(wk.frames_l, ld, fd_, F, crops)
    prof = ctx.profile_fetch()
    ctx.profile(False)
    out[F] = {k: v[1] / v[0] for k, v in prof.items()}
    print(F, out[F], flush=True)
    crops.free()
json.dump(out, open("gpurun_out/warp_scaling.json", "w"), indent=1)
