"""Per-kernel times of fd_decode_jpeg_batch on the bench's 64 x 1080p frames (q90 4:2:0), restart interval sweep.
Kernel times come from CUDA events recorded right before and after each kernel (torch events on the ctx stream are not
available inside the call, so the huffman kernel is isolated by timing a second identical call with FD profile marks and
subtracting the copy-only time measured separately)."""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from rs_face_detection_b200 import Context
from rs_face_detection_b200.ffi import pinned_like
from rs_face_detection_b200.utils import synth
ctx = Context(0)
frames = [synth.make_frame(1080, 1920, 2000 + i) for i in range(64)]
out = {}
for rst in [int(x) for x in os.environ.get("RSTS", "0,4,16,120,480").split(",")]:
    pj = [pinned_like(np.asarray(cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 90] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else []))[1], np.uint8).ravel()) for f in frames]
    st = [p.array for p in pj]
    for _ in range(2):
        ctx.decode_jpeg_batch(st, n_threads=16); ctx.synchronize()
    ctx.profile(True)
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        ctx.decode_jpeg_batch(st, n_threads=16)
    ctx.synchronize()
    wall = (time.perf_counter() - t0) / n
    prof = ctx.profile_fetch()
    ctx.profile(False)
    out[rst] = dict(wall_ms=wall * 1e3, mb=sum(s.size for s in st) / 1e6, kernels={k: v[1] / v[0] for k, v in prof.items()}, stats=ctx.jpeg_last_stats())
    print(rst, out[rst], flush=True)
json.dump(out, open(os.environ.get("OUT", "gpurun_out/jpeg_profile.json"), "w"), indent=1)
