"""Generates tests/golden/jpeg_golden.npz: JPEG streams encoded by cv2 (libjpeg-turbo) and the BGR arrays cv2.imdecode(...,
IMREAD_UNCHANGED) returns for them — the reference's byte_data_to_opencv (utils.rs:8-52) output.  Run in a container that has
cv2 (this one: 4.13.0, libjpeg-turbo 3.1.2); the fixtures are committed so the GPU box does not need it."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SS = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}


def photo(h, w, seed):
    r = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    img = np.stack([127 + 90 * np.sin(xx / (7 + 3 * c) + c) * np.cos(yy / (11 + 2 * c) - c) + r.normal(0, 12, (h, w)) for c in range(3)], -1)
    img[h // 4:h // 2, w // 3:w // 2] += 60
    return np.clip(img, 0, 255).astype(np.uint8)


CASES = [  # (h, w, sampling, quality, restart interval, optimize)
    (97, 131, "420", 85, 0, 0), (64, 48, "444", 95, 0, 0), (50, 75, "422", 60, 2, 0), (16, 16, "420", 30, 0, 0), (2, 3, "422", 90, 0, 0),
    (1, 1, "420", 75, 0, 0), (33, 4, "420", 100, 0, 0), (270, 480, "420", 90, 0, 1), (135, 241, "422", 75, 5, 0), (120, 120, "444", 10, 0, 0),
]

if __name__ == "__main__":
    out = {"n": np.int32(len(CASES))}
    for i, (h, w, ss, q, rst, opt) in enumerate(CASES):
        params = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SS[ss]]
        if rst:
            params += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
        if opt:
            params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
        ok, buf = cv2.imencode(".jpg", photo(h, w, 100 + i), params)
        assert ok
        out["jpeg_%d" % i] = np.asarray(buf, np.uint8).ravel()
        out["bgr_%d" % i] = cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)
    ok, buf = cv2.imencode(".jpg", photo(64, 64, 7), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    out["unsupported_progressive"] = np.asarray(buf, np.uint8).ravel()
    ok, buf = cv2.imencode(".jpg", photo(64, 64, 7)[:, :, 0])
    out["unsupported_gray"] = np.asarray(buf, np.uint8).ravel()
    np.savez_compressed(os.path.join(HERE, "jpeg_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "jpeg_golden.npz"), "cv2", cv2.__version__)
