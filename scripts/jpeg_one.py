import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from rs_face_detection_b200 import Context
from rs_face_detection_b200.ffi import pinned_like
from rs_face_detection_b200.utils import synth
rst = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nimg = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = Context(0)
frames = [synth.make_frame(1080, 1920, 2000 + i) for i in range(nimg)]
pj = [pinned_like(np.asarray(cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 90] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else []))[1], np.uint8).ravel()) for f in frames]
for _ in range(2):
    ctx.decode_jpeg_batch([p.array for p in pj], n_threads=4)
    ctx.synchronize()
print("done")
