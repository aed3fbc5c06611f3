"""Generates tests/golden/cv2_golden.npz with THIS container's cv2 (4.13.0) — the stand-in for the OpenCV the
reference links through the `opencv` crate 0.92.0 (Cargo.lock:1175).  Inputs and cv2 outputs for the three library
calls on the hot path: resize (face_detection.rs:156), estimateAffinePartial2D (face_alignment.rs:50-59),
warpAffine (face_alignment.rs:119-126).  Run:  python tests/golden/make_golden.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TEMPLATE = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366], [41.5493, 92.3655], [70.7299, 92.2041]], np.float32)


def main():
    cv2.setNumThreads(1)
    rng = np.random.default_rng(20261018)
    out = {"cv2_version": np.array(cv2.__version__)}
    # resize: (src h,w) -> (dst w,h); includes exact 3x and 6x down-scales (1080p/4K -> 640x360 geometry), up-scales, odd sizes
    rs = [((54, 96), (32, 18)), ((108, 192), (32, 18)), ((31, 57), (64, 50)), ((90, 120), (40, 30)), ((45, 33), (66, 90)),
          ((64, 64), (64, 64)), ((7, 9), (40, 31))]
    for i, ((sh, sw), (dw, dh)) in enumerate(rs):
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        out["resize_in_%d" % i] = img
        out["resize_dsize_%d" % i] = np.array([dw, dh], np.int32)
        out["resize_out_%d" % i] = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    out["resize_n"] = np.array(len(rs))
    # warpAffine -> 112x112, crops hanging off the borders included
    nw = 6
    for i in range(nw):
        sh, sw = [(96, 128), (140, 100), (64, 64)][i % 3]
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        s = rng.uniform(0.5, 2.5)
        th = rng.uniform(-0.6, 0.6)
        a, b = s * np.cos(th), s * np.sin(th)
        cx, cy = rng.uniform(-10, sw + 10), rng.uniform(-10, sh + 10)
        M = np.array([[a, -b, 56 - (a * cx - b * cy)], [b, a, 56 - (b * cx + a * cy)]], np.float64)
        out["warp_in_%d" % i] = img
        out["warp_M_%d" % i] = M
        out["warp_out_%d" % i] = cv2.warpAffine(img, M, (112, 112), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    out["warp_n"] = np.array(nw)
    # estimateAffinePartial2D LMEDS on noisy similarity-transformed templates (every 5th set has one gross outlier)
    n = 300
    pts_all = np.empty((n, 5, 2), np.float32)
    M_all = np.zeros((n, 2, 3), np.float64)
    inl_all = np.zeros((n, 5), np.uint8)
    ok_all = np.zeros(n, np.uint8)
    for t in range(n):
        s = rng.uniform(0.4, 3.0)
        th = np.deg2rad(rng.uniform(-35, 35))
        R = s * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        pts = (TEMPLATE - 56) @ R.T + rng.uniform(0, 1000, 2)
        pts = pts + rng.normal(0, 0.01 * 112 * s, (5, 2))
        if t % 5 == 0:
            pts[rng.integers(5)] += rng.normal(0, 0.2 * 112 * s, 2)
        pts = pts.astype(np.float32)
        M, inl = cv2.estimateAffinePartial2D(pts, TEMPLATE, method=cv2.LMEDS, ransacReprojThreshold=3.0, maxIters=2000,
                                             confidence=0.99, refineIters=10)
        pts_all[t] = pts
        if M is not None:
            M_all[t], inl_all[t], ok_all[t] = M, inl.ravel(), 1
    out.update(est_pts=pts_all, est_M=M_all, est_inliers=inl_all, est_ok=ok_all)
    np.savez_compressed(os.path.join(HERE, "cv2_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "cv2_golden.npz"))


if __name__ == "__main__":
    main()
