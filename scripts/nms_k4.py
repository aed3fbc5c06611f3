import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from rs_face_detection_b200 import Context
ctx = Context(0)
dets = np.array([[10,10,50,50,0.9],[20,20,60,60,0.8],[15,15,55,55,0.95],[10,10,50,50,0.5]], np.float32)
print("K=4", ctx.nms(dets, 0.4).tolist(), flush=True)
rng = np.random.default_rng(0)
for K in (40, 400, 1000):
    xy = rng.uniform(0, 600, (K, 2)); wh = rng.uniform(10, 80, (K, 2))
    d = np.concatenate([xy, xy + wh, rng.uniform(0, 1, (K, 1))], 1).astype(np.float32)
    print(K, len(ctx.nms(d, 0.4)), flush=True)
